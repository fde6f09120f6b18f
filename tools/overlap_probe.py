#!/usr/bin/env python
"""Experiment: do the tree kernel of one half of the games and the net of the other half overlap when the two halves
run as independent wave loops on two streams?  Prints ms per UCT_search(800) of 4096 games for 1 x 4096 and 2 x 2048."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedSimpleNN
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

SIMS, GW = 800, 16
torch.manual_seed(0)
model = SimpleNN(board=(3, 3))


def make(n, seed, inline):
    eng = engine.Engine((3, 3), n_games=n, max_nodes=SIMS + 8)
    eng.set_mode(False, inline)
    ev = FusedSimpleNN(model, eng)
    roots = bench.synthetic_roots(eng, torch, seed)
    return eng, ev, roots


def run(parts, reps=3):
    streams = [torch.cuda.Stream() for _ in parts]
    graphs = []
    for (eng, ev, roots), s in zip(parts, streams):
        with torch.cuda.stream(s):
            graphs.append(eng._capture(ev, GW, None, 0.0, 1))
    torch.cuda.synchronize()
    n_rep = (SIMS + 2 + GW - 1) // GW
    best = 1e9
    for _ in range(reps):
        for (eng, ev, roots), s in zip(parts, streams):
            with torch.cuda.stream(s):
                eng.reset_roots(roots)
                eng.begin(SIMS)
        torch.cuda.synchronize()
        t0 = time.time()
        for r in range(n_rep):
            for g, s in zip(graphs, streams):
                with torch.cuda.stream(s):
                    g.replay()
        torch.cuda.synchronize()
        best = min(best, time.time() - t0)
    sims = sum(p[0].status()["sims"] for p in parts)
    return best * 1e3, sims


for inline in (1, 4):
    one = [make(4096, 1, inline)]
    ms, sims = run(one)
    print("inline %d: 1 x 4096: %.1f ms  (%.2f M sims/s)" % (inline, ms, sims / ms / 1e3), flush=True)
    del one
    two = [make(2048, 1, inline), make(2048, 2, inline)]
    ms, sims = run(two)
    print("inline %d: 2 x 2048 on two streams: %.1f ms  (%.2f M sims/s)" % (inline, ms, sims / ms / 1e3), flush=True)
    del two
    four = [make(1024, i, inline) for i in range(4)]
    ms, sims = run(four)
    print("inline %d: 4 x 1024 on four streams: %.1f ms  (%.2f M sims/s)" % (inline, ms, sims / ms / 1e3), flush=True)
    del four
