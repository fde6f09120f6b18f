#!/usr/bin/env python
"""ResNetZero evaluator per batch size: the tcgen05 tower plan against the library (cuDNN) plan, same weights, same
engine buffers; CUDA events over graph replays of 4 evaluations.
    python tools/tower_vs_cudnn.py [--board 5x5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedResNetZero, ResNetZero, resnet_zero_parameters
from dotsboxesaz_b200.utils.utils import DotDict

ap = argparse.ArgumentParser()
ap.add_argument("--board", default="5x5")
ap.add_argument("--games", type=int, default=16384)
args = ap.parse_args()
L, C = (int(v) for v in args.board.split("x"))
torch.manual_seed(0)
model = ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters((L, C))}}))
eng = engine.Engine((L, C), n_games=args.games, max_nodes=16)
eng.leaf_states.copy_(eng.new_states(args.games))
plans = {"tower": FusedResNetZero(model, eng), "cudnn": FusedResNetZero(model, eng, use_tower=False)}
rows_list = sorted(set(eng._ladder(plans["tower"])) | {64, 256, 1024, 4096, args.games})


def timed(ev, rows):
    eng._batch_rows = None if rows == args.games else rows
    for _ in range(3):
        ev(eng)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            ev(eng)
    g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    eng._batch_rows = None
    return a.elapsed_time(b) / 12 * 1e3


print("| leaves | tower plan us | cuDNN plan us | speed-up |")
print("|---|---|---|---|")
for rows in rows_list:
    t, c = timed(plans["tower"], rows), timed(plans["cudnn"], rows)
    print("| %d | %.0f | %.0f | %.2fx |" % (rows, t, c, c / t), flush=True)
eng.close()
