#!/usr/bin/env python
"""Evaluator time per batch size (FusedSimpleNN bf16, CUDA events over graph replays) and the batch schedule one
adaptive bench search runs through."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedSimpleNN
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
import bench

torch.manual_seed(0)
model = SimpleNN(board=(3, 3))
eng = engine.Engine((3, 3), n_games=4096, max_nodes=808, eval_cache=24)
eng.LADDER_STEPS = 16
ev = FusedSimpleNN(model, eng)
curve = {}
for rows in eng._ladder():
    eng._batch_rows = rows
    for _ in range(3):
        ev(eng)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(8):
            ev(eng)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    curve[rows] = a.elapsed_time(b) / 40 * 1e3
    print("net batch %5d: %6.1f us  (%.1f evals/us)" % (rows, curve[rows], rows / curve[rows]), flush=True)
eng._batch_rows = None
roots = bench.synthetic_roots(eng, torch, 1234)
valid = eng.valid_moves(roots).cpu().numpy()
noise = torch.from_numpy(bench.host_noise(np.random.RandomState(99), valid, 0.8)).cuda()
for rep in range(2):
    eng.reset_roots(roots)
    eng.clear_eval_cache()
    eng.run_search(800, ev, noise=noise, coeff=0.25, graph_waves=8, adaptive=True)
torch.cuda.synchronize()
sch = eng.last_schedule
tot = sum(curve[r] * 8 for r, _ in sch)
print("replays", len(sch), "net time by curve %.1f ms" % (tot / 1e3))
print("engine's own curve:", {r: round(us, 1) for (_, r), us in sorted(eng._eval_us.items(), key=lambda kv: -kv[0][1])})
print(" ".join("%d/%d" % (r, b) for r, b in sch))
st = eng.status()
print("sims", st["sims"], "hits", st["cache_hits"], "term", st["terminal_leaves"], "evals", st["sims"] - st["cache_hits"] - st["terminal_leaves"])
