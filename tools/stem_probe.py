#!/usr/bin/env python
"""Time the two stem kernels (tensor-core implicit GEMM vs CUDA-core FMA) in isolation, CUDA-graph replays."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedSimpleNN, FusedResNetZero, ResNetZero, resnet_zero_parameters
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
from dotsboxesaz_b200.utils.utils import DotDict
import bench


def timeit(fn):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); g.replay(); b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 40 * 1e3


for board, n, net in (((3, 3), 4096, "simple"), ((3, 3), 32768, "simple"), ((5, 5), 16384, "resnet")):
    eng = engine.Engine(board, n_games=n, max_nodes=16)
    torch.manual_seed(0)
    if net == "simple":
        model = SimpleNN(board=board)
        pm, pf = FusedSimpleNN(model, eng), FusedSimpleNN(model, eng, use_stem="fma")
        fma = lambda: eng.nn_stem(eng.leaf_states, *pf.stem, pf.stem_out[:n], mode=0)
    else:
        model = ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters(board)}}))
        pm, pf = FusedResNetZero(model, eng), FusedResNetZero(model, eng, use_stem="fma")
        fma = lambda: eng.nn_stem(eng.leaf_states, *pf.fused_stem, pf.stem_out[:n], mode=1)
    eng.leaf_states.copy_(bench.synthetic_roots(eng, torch, 1))
    mma = lambda: eng.nn_stem_mma(eng.leaf_states, pm.stem_mma, pm.stem_out[:n])
    out_mb = pm.stem_out[:n].numel() * 2 / 1e6
    t_m, t_f = timeit(mma), timeit(fma)
    print("%s %s n=%d: mma %.1f us (%.0f GB/s written), fma %.1f us; output %.1f MB" % (board, net, n, t_m, out_mb / t_m * 1e3, t_f, out_mb))
    eng.close()
