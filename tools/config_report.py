#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[3] on one B200 (configs[1] is bench.py's line, configs[4] is tools/coach_bench.py):
  configs[2]  game-engine-only random rollouts, 5x5 boxes, 1M concurrent games: plies/s, games/s, GB/s
  configs[3]  5x5 boxes, 16384 concurrent games, 800 sims/move, ResNetZero bf16: sims/s of one UCT_search, k_search_step time
Prints one JSON object per config; --cpu also times the C oracle's rollouts on one host core."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rollout-games", type=int, default=1 << 20)
    ap.add_argument("--games", type=int, default=16384)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--eval-cache", type=int, default=23)
    ap.add_argument("--max-inline", type=int, default=4)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--skip-search", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from dotsboxesaz_b200 import engine
    from dotsboxesaz_b200.nn import FusedResNetZero, ResNetZero, resnet_zero_parameters
    from dotsboxesaz_b200.utils.utils import DotDict
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))

    # ---- configs[2]: rollouts
    n = args.rollout_games
    eng = engine.Engine((5, 5), n_games=1, max_nodes=4)
    best = None
    for rep in range(4):
        st = eng.new_states(n)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        plies = eng.random_rollout(st, seed=rep)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if rep and (best is None or ms < best[0]):
            best = (ms, int(plies.sum().item()))
    ms, total_plies = best
    res = eng.result(st)
    out = {"config": "configs[2]: 5x5 random rollouts to terminal, %d concurrent games, one kernel (k_game_rollout, thread per game, Philox4x32-10)" % n,
           "ms": ms, "games_per_sec": n / ms * 1e3, "plies_per_sec": total_plies / ms * 1e3, "plies_per_game": total_plies / n,
           "all_terminal": bool((res != 2).all().item()),
           # the rollout keeps the state in registers: HBM traffic is one 32-byte state in + out per GAME
           "hbm_bytes": 64 * n, "hbm_gbs": 64 * n / ms / 1e6,
           "note": "state lives in registers for the whole game; the kernel is bound by the per-ply integer work "
                   "(Philox + legal-move select + box test), not by HBM"}
    if args.cpu:
        from oracle import oracle
        g = oracle.OracleGame(5, 5)
        t0 = time.time()
        cnt = pl = 0
        while time.time() - t0 < 3.0:
            gg = oracle.OracleGame(5, 5)
            pl += len(gg.random_rollout(0, cnt))
            cnt += 1
        dt = time.time() - t0
        out["cpu_c_oracle_1core"] = {"games_per_sec": cnt / dt, "plies_per_sec": pl / dt, "note": "incl. ctypes call overhead per game"}
    print(json.dumps(out), flush=True)
    eng.close()
    if args.skip_search:
        return

    # ---- configs[3]: 5x5 search
    board = (5, 5)
    eng = engine.Engine(board, n_games=args.games, max_nodes=args.sims + 8, eval_cache=args.eval_cache)
    eng.set_mode(False, args.max_inline)
    eng.LADDER_STEPS = 16
    torch.manual_seed(0)
    model = ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters(board)}}))
    ev = FusedResNetZero(model, eng)
    roots = bench.synthetic_roots(eng, torch, seed=1234)
    valid = eng.valid_moves(roots).cpu().numpy()
    noise = torch.from_numpy(bench.host_noise(np.random.RandomState(99), valid, 0.8)).to(eng.device)

    def step():
        eng.reset_roots(roots)
        eng.clear_eval_cache()
        eng.run_search(args.sims, ev, noise=noise, coeff=0.25, graph_waves=8, adaptive=True)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    w0 = eng.n_waves
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 2
    for _ in range(reps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    st = eng.status()
    sims = args.games * args.sims
    P = st["path_nodes"] / st["sims"]
    f_term, f_hit = st["terminal_leaves"] / st["sims"], st["cache_hits"] / st["sims"]
    f_miss = 1 - f_term - f_hit
    A, F = eng.A, eng.F
    bps = (P - 1) * (13 * A + 24) + 36 + 24 * P + (1 - f_term) * (13 * A + 4) + f_miss * (F * 2 + 4 * A) + (1 - f_term) * 16 * A + f_miss * 16 * A
    # evaluator alone at full width
    for _ in range(3):
        ev(eng)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        ev(eng)
    b.record()
    torch.cuda.synchronize()
    net_us = a.elapsed_time(b) / 10 * 1e3
    flops = 106.5e6  # per position, BASELINE.md
    sched = getattr(eng, "last_schedule", [])
    print("schedule (rows/busy per replay):", " ".join("%d/%d" % rb for rb in sched[::4]), file=sys.stderr)
    print("evaluator us per rung:", {r: round(u) for (_, r), u in sorted(eng._eval_us.items(), key=lambda kv: -kv[0][1])}, file=sys.stderr)
    print(json.dumps({"config": "configs[3]: 5x5 boxes, %d concurrent games, %d sims/move, ResNetZero(64ch x 20 blocks) bf16, one UCT_search from "
                                "synthetic roots, Dirichlet(0.8, 0.25), eval cache 2^%d emptied per step" % (args.games, args.sims, args.eval_cache),
                      "ms_per_search": ms, "sims_per_sec": sims / ms * 1e3, "waves": (eng.n_waves - w0) / reps,
                      "mean_path_nodes": P, "terminal_leaf_frac": f_term, "cache_hit_frac": f_hit, "algorithmic_bytes_per_sim": bps,
                      "net_us_per_%d_leaves" % args.games: net_us, "net_evals_per_sec": args.games / net_us * 1e6,
                      "net_tflops": args.games * flops / net_us / 1e6, "net_frac_of_sustained_bf16_peak":
                          args.games * flops / net_us / 1e6 / float(peaks.get("bf16_tflops_sustained", 1400.0)),
                      "hbm_peak_gbs": hbm}), flush=True)


if __name__ == "__main__":
    main()
