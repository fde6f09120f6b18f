"""Where a self-play check goes: CUDA-event time of every graph replay of play_games_async (rung graphs = 16 waves of
[step -> evaluator], finish graph = sample / draw the move / re-root / begin), grouped by graph.

  python tools/finish_probe.py --games 32768
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dotsboxesaz_b200 import engine, self_play
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
from dotsboxesaz_b200.nn import FusedSimpleNN
from dotsboxesaz_b200.utils.utils import DotDict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=32768)
    ap.add_argument("--sims", type=int, default=800)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = engine.Engine((3, 3), n_games=args.games, max_nodes=4096, device=dev, eval_cache=24)
    eng.set_mode(False, 4)
    torch.manual_seed(0)
    ev = FusedSimpleNN(SimpleNN(board=(3, 3)), eng, dtype=torch.bfloat16)
    params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": bench.NOISE,
                                    "mcts": {"mcts_num_read": args.sims, "mcts_cpuct": (1.25, 19652),
                                             "temperature": {0: 1.0, 12: 0.02}, "max_async_searches": 1}}})

    def play(seed):
        sp = self_play.BatchedSelfPlay(eng, ev, params, graph_waves=16, adaptive=True)
        eng.clear_eval_cache()
        return sp.play_games_async(range(args.games), seed=seed)

    play(1)
    torch.cuda.synchronize()
    rec = []
    orig = torch.cuda.CUDAGraph.replay

    def replay(self):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(self)
        b.record()
        rec.append((id(self), a, b))
    torch.cuda.CUDAGraph.replay = replay
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    play(2)
    b.record()
    torch.cuda.synchronize()
    torch.cuda.CUDAGraph.replay = orig
    total = a.elapsed_time(b)
    by = collections.defaultdict(list)
    for i, x, y in rec:
        by[i].append(x.elapsed_time(y))
    print("whole batch: %.1f ms, %d replays" % (total, len(rec)))
    for i, v in sorted(by.items(), key=lambda kv: -sum(kv[1])):
        print("graph %x: %5d replays, mean %.3f ms, sum %.1f ms = %.1f %%" % (i & 0xffffff, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / total))


if __name__ == "__main__":
    main()
