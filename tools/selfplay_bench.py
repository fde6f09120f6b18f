#!/usr/bin/env python
"""Full self-play games (not single searches): games/hour and sims/s, BASELINE configs[1] and configs[3].
    python tools/selfplay_bench.py --board 3x3 --games 4096 --net simple
    python tools/selfplay_bench.py --board 5x5 --games 16384 --net resnet --max-nodes 4096
Prints one JSON line per run."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--board", default="3x3")
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--net", default="simple", choices=["simple", "resnet", "fake"])
    ap.add_argument("--max-nodes", type=int, default=8192)
    ap.add_argument("--mode", default="device", choices=["device", "async", "host"])
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--pending", type=int, default=1, help="max_async_searches (simulations in flight per tree)")
    ap.add_argument("--no-warmup", action="store_true")
    ap.add_argument("--eval-cache", type=int, default=0, help="log2(entries) of the device eval cache; 0 = off")
    ap.add_argument("--max-inline", type=int, default=4)
    ap.add_argument("--adaptive", type=int, default=-1, help="1/0 adaptive wave loop; -1 = on iff the cache is on")
    ap.add_argument("--graph-waves", type=int, default=8)
    ap.add_argument("--ladder-steps", type=int, default=16)
    ap.add_argument("--keep-cache", action="store_true", help="do not empty the cache between repeats (weights are fixed)")
    args = ap.parse_args()
    import torch
    from dotsboxesaz_b200 import engine, self_play
    from dotsboxesaz_b200.nn import FusedResNetZero, FusedSimpleNN, ResNetZero, resnet_zero_parameters
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    from dotsboxesaz_b200.utils.utils import DotDict
    L, C = (int(x) for x in args.board.split("x"))
    eng = engine.Engine((L, C), n_games=args.games, max_nodes=args.max_nodes, max_pending=args.pending,
                        eval_cache=args.eval_cache if args.pending == 1 else 0)
    eng.set_mode(False, args.max_inline)
    eng.LADDER_STEPS = args.ladder_steps
    torch.manual_seed(0)
    if args.net == "fake":
        ev = engine.FakeNetEvaluator(0)
    elif args.net == "simple":
        ev = FusedSimpleNN(SimpleNN(board=(L, C)), eng)
    else:
        ev = FusedResNetZero(ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters((L, C))}})), eng)
    params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": (0.8, 0.25),
                                    "mcts": {"mcts_num_read": args.sims, "mcts_cpuct": (1.25, 19652),
                                             "temperature": {0: 1.0, 12: 0.02}, "max_async_searches": args.pending}}})
    for rep in range(args.repeats + 1):  # first pass warms up (graphs, cuDNN)
        sp = self_play.BatchedSelfPlay(eng, ev, params, graph_waves=args.graph_waves, adaptive=None if args.adaptive < 0 else bool(args.adaptive))
        if not args.keep_cache:
            eng.clear_eval_cache()
        w0 = eng.n_waves
        torch.cuda.synchronize()
        t0 = time.time()
        if args.mode in ("device", "async"):
            info = (sp.play_games_device if args.mode == "device" else sp.play_games_async)(range(args.games), seed=rep)
            rows = sp.device_samples()[0].shape[0]
        else:
            sp.play_games(range(args.games), seeds=range(rep * args.games, (rep + 1) * args.games))
            info = {"sims": sp.total_sims, "max_nodes_used": None, "path_nodes": 0, "terminal_leaves": 0}
            rows = len(sp.rows)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if rep == 0 and not args.no_warmup:
            continue
        print(json.dumps({"board": args.board, "games": args.games, "sims_per_move": args.sims, "net": args.net, "mode": args.mode, "max_pending_evals": args.pending,
                          "eval_cache_log2": args.eval_cache, "max_inline": args.max_inline, "adaptive": sp.adaptive, "waves": eng.n_waves - w0,
                          "cache_hit_frac": info.get("cache_hits", 0) / max(1, info["sims"]),
                          "seconds": dt, "games_per_hour": args.games / dt * 3600, "sims_per_sec": info["sims"] / dt,
                          "total_sims": info["sims"], "sample_rows": rows, "max_nodes_used": info["max_nodes_used"],
                          "mean_path_nodes": info["path_nodes"] / max(1, info["sims"]),
                          "terminal_leaf_frac": info["terminal_leaves"] / max(1, info["sims"])}), flush=True)


if __name__ == "__main__":
    main()
