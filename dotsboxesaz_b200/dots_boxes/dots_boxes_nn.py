"""SimpleNN (reference: dots_boxes/dots_boxes_nn.py:61-105), same module names and shapes so the
reference's checkpoints load; generalised from the hard-wired 3x3 board to any L x C."""
import logging

import torch
import torch.nn.functional as F
from torch import nn

from ..nn import _load_parameters

logger = logging.getLogger(__name__)
N_CH = 256


class SimpleNN(nn.Module):
    """5 x [conv3x3(256) -> ReLU -> BN] (the last conv unpadded) -> FC 512 -> FC 256 -> {tanh value, log-softmax policy}.
    board=(3, 3) gives fc0 in = 1024 and 32 policy logits exactly as the reference."""

    def __init__(self, params=None, board=None):
        super().__init__()
        self.params = params
        if board is None:
            board = (3, 3)
            try:
                board = tuple(params.game.clazz.BOARD_DIM)
            except Exception:
                pass
        rows, cols = board[0] + 1, board[1] + 1
        self.conv0 = nn.Conv2d(3, N_CH, 3, padding=1)
        self.bn0 = nn.BatchNorm2d(N_CH)
        for i in (1, 2, 3):
            setattr(self, f"conv{i}", nn.Conv2d(N_CH, N_CH, 3, padding=1))
            setattr(self, f"bn{i}", nn.BatchNorm2d(N_CH))
        self.conv4 = nn.Conv2d(N_CH, N_CH, 3, padding=0)
        self.bn4 = nn.BatchNorm2d(N_CH)
        self.fc0 = nn.Linear(N_CH * (rows - 2) * (cols - 2), 512)
        self.bn_fc0 = nn.BatchNorm1d(512)
        self.fc1 = nn.Linear(512, 256)
        self.bn_fc1 = nn.BatchNorm1d(256)
        self.value_fc = nn.Linear(256, 1)
        self.policy_fc = nn.Linear(256, 2 * rows * cols)

    def forward(self, x):
        for i in range(5):
            x = getattr(self, f"bn{i}")(F.relu(getattr(self, f"conv{i}")(x)))
        x = x.reshape(x.size(0), -1)  # NCHW flatten order, also for channels_last inputs
        x = self.bn_fc0(F.relu(self.fc0(x)))
        x = self.bn_fc1(F.relu(self.fc1(x)))
        return F.log_softmax(self.policy_fc(x), dim=1), torch.tanh(self.value_fc(x))

    def load_parameters(self, generation, to_device=None):
        _load_parameters(self, generation, to_device)
