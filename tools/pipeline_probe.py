"""Two half-size engines on two streams against one full-size engine (configs[3]: 5x5, ResNetZero, 800 sims/move).

The step kernel of one half can run in the shadow of the other half's evaluator (the tower kernel leaves 2.4 KB of shared
memory and 24 k registers per SM: one k_search_step CTA fits beside it).  Prints sims/s for both arrangements.

  python tools/pipeline_probe.py --board 5x5 --games 16384 --sims 800
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedResNetZero, FusedSimpleNN, ResNetZero, resnet_zero_parameters
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
from dotsboxesaz_b200.utils.utils import DotDict


def make(board, games, sims, net, dev, seed, cache):
    eng = engine.Engine(board, n_games=games, max_nodes=sims + 8, device=dev, eval_cache=cache)
    eng.set_mode(False, 4)
    torch.manual_seed(0)
    if net == "simple":
        ev = FusedSimpleNN(SimpleNN(board=board), eng, dtype=torch.bfloat16)
    else:
        model = ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters(board)}}))
        ev = FusedResNetZero(model, eng, dtype=torch.bfloat16)
    ns = argparse.Namespace(games=games)
    roots = bench.synthetic_roots(eng, torch, seed=seed)
    valid = eng.valid_moves(roots).cpu().numpy()
    noise = torch.from_numpy(bench.host_noise(np.random.RandomState(seed), valid, bench.NOISE[0])).to(dev)
    return eng, ev, roots, noise


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--board", default="5x5")
    ap.add_argument("--games", type=int, default=16384)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--net", default="resnet")
    ap.add_argument("--cache", type=int, default=24)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--parts", type=int, default=2)
    args = ap.parse_args()
    board = tuple(int(x) for x in args.board.split("x"))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)

    def run(parts):
        n = args.games // parts
        units = [make(board, n, args.sims, args.net, dev, 1234 + i, args.cache) for i in range(parts)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(parts)]

        def step():
            for (eng, ev, roots, noise), s in zip(units, streams):
                with torch.cuda.stream(s):
                    eng.reset_roots(roots)
                    eng.clear_eval_cache()
                    eng.run_search(args.sims, ev, noise=noise, coeff=bench.NOISE[1], graph_waves=8, pending=1, adaptive=True)
        for _ in range(3):
            step()
            torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        sec = (time.time() - t0) / args.steps
        for eng, *_ in units:
            eng.status()
            eng.close()
        del units
        torch.cuda.empty_cache()
        return n * parts * args.sims / sec, sec

    for parts in (1, args.parts):
        v, sec = run(parts)
        print("%d engine(s) x %d games: %.2f M sims/s (%.1f ms per step)" % (parts, args.games // parts, v / 1e6, sec * 1e3), flush=True)


if __name__ == "__main__":
    main()
