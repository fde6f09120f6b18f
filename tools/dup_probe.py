#!/usr/bin/env python
"""How many of the leaves the net evaluates in one bench step are duplicates (same cache key evaluated more than once:
same-wave races between trees, evictions of the direct-mapped table)?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dotsboxesaz_b200 import engine
import bench

eng = engine.Engine((3, 3), n_games=4096, max_nodes=808, eval_cache=24)
eng.set_mode(True, 4)
ev = engine.FakeNetEvaluator(0)
roots = bench.synthetic_roots(eng, torch, 1234)
valid = eng.valid_moves(roots).cpu().numpy()
noise = torch.from_numpy(bench.host_noise(np.random.RandomState(99), valid, 0.8)).cuda()
eng.reset_roots(roots)
eng.clear_eval_cache()
eng.begin(800, noise, 0.25, 1)
keys, per_wave_dups, waves = [], [], 0
while True:
    eng.step()
    rows, busy = eng.wave_counts()
    if rows:
        st = eng.leaf_states[:rows].clone()
        tp = (st[:, 2] >> 32) & 0xff
        btc = torch.where(tp == 1, (st[:, 2] >> 16) & 0xffff, st[:, 2] & 0xffff)
        k = (st[:, 0] & 0xffffffff) | (btc << 40)
        keys.append(k)
        per_wave_dups.append(rows - torch.unique(k).numel())
    ev(eng)
    waves += 1
    if busy == 0:
        break
allk = torch.cat(keys)
print("waves", waves, "evaluated", allk.numel(), "distinct", torch.unique(allk).numel(),
      "same-wave duplicates", int(sum(per_wave_dups)), "first waves:", per_wave_dups[:12])
