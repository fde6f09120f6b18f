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


def _edge_permutations(L, C):
    """For each of the 8 board symmetries (4 for non-square boards) the gather index over the A action slots:
    out[:, a] = in[:, perm[a]].  An edge is a pair of dots; a symmetry maps dots, hence edges; padding slots map
    to padding slots."""
    rows, cols = L + 1, C + 1
    plane = rows * cols

    def slot(p, l, c):
        return p * plane + l * cols + c

    def edge_dots(p, l, c):
        return ((l, c), (l, c + 1)) if p == 0 else ((l, c), (l + 1, c))

    def dots_edge(d0, d1):
        (y0, x0), (y1, x1) = sorted((d0, d1))
        return (0, y0, x0) if y0 == y1 else (1, y0, x0)

    def build(fn, transposed):
        perm = list(range(2 * plane))  # padding slots stay where they are unless mapped below
        Lo, Co = (C, L) if transposed else (L, C)
        assert (Lo, Co) == (L, C), "transposing symmetries need a square board"
        for p in range(2):
            for l in range(rows):
                for c in range(cols):
                    real = (c < C) if p == 0 else (l < L)
                    if not real:
                        continue
                    d0, d1 = edge_dots(p, l, c)
                    q, ql, qc = dots_edge(fn(*d0), fn(*d1))
                    perm[slot(q, ql, qc)] = slot(p, l, c)  # destination slot takes the value of the source edge
        return perm

    flips = [lambda y, x: (y, x), lambda y, x: (L - y, x), lambda y, x: (y, C - x), lambda y, x: (L - y, C - x)]
    perms = [build(f, False) for f in flips]
    if L == C:
        perms += [build(lambda y, x, f=f: tuple(reversed(f(y, x))), True) for f in flips]
    return torch.tensor(perms, dtype=torch.long)


class SymmetriesGenerator(nn.Module):
    """Random dihedral symmetry of a batch of (boards, policies) (reference: dots_boxes_nn.py:11-58, applied once
    per training batch).  Implemented as one gather with a precomputed edge permutation instead of flip / cat chains;
    index i matches the reference's IDXS[i] (flip dims (1,), (2,), (1,2); +4 = followed by the transpose)."""

    def __init__(self):
        super().__init__()
        self._cache = {}

    def permutations(self, rows, cols, device):
        key = (rows, cols, str(device))
        if key not in self._cache:
            self._cache[key] = _edge_permutations(rows - 1, cols - 1).to(device)
        return self._cache[key]

    @torch.no_grad()
    def forward(self, boards, policies, index=None):
        import random
        n, ch, rows, cols = boards.shape
        perms = self.permutations(rows, cols, boards.device)
        i = random.randint(0, 7) if index is None else index
        if i >= perms.shape[0]:
            i %= perms.shape[0]
        if i == 0:
            return boards, policies
        perm = perms[i]
        flat = boards[:, :2].reshape(n, -1)[:, perm].reshape(n, 2, rows, cols)
        out = torch.cat((flat, boards[:, 2:]), 1) if ch > 2 else flat
        return out, policies[:, perm]
