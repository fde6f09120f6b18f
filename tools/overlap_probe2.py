"""Can k_search_step CTAs run UNDER the evaluator's kernels?  Engine 1 runs the SimpleNN (or ResNetZero tower) evaluator
in a loop on one stream; engine 2 runs tree waves with the fake net on another.  Times: each alone, then both together.
    python tools/overlap_probe2.py [--net simple|resnet] [--board 3x3]
"""
import argparse
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dotsboxesaz_b200 import engine  # noqa: E402
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN  # noqa: E402
from dotsboxesaz_b200.nn import FusedResNetZero, FusedSimpleNN, ResNetZero, resnet_zero_parameters  # noqa: E402
from dotsboxesaz_b200.utils.utils import DotDict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="simple")
    ap.add_argument("--board", default="3x3")
    ap.add_argument("--games", type=int, default=4096)
    args = ap.parse_args()
    L, C = (int(v) for v in args.board.split("x"))
    dev = torch.device("cuda:0")
    e1 = engine.Engine((L, C), n_games=args.games, max_nodes=64, device=dev)
    e2 = engine.Engine((L, C), n_games=args.games, max_nodes=1024, device=dev)
    torch.manual_seed(0)
    model = SimpleNN(board=(L, C)) if args.net == "simple" else ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters((L, C))}}))
    ev = (FusedSimpleNN if args.net == "simple" else FusedResNetZero)(model, e1)
    e1.leaf_states.copy_(e1.new_states(args.games))
    fake = engine.FakeNetEvaluator(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    n_eval, n_waves = 20, 100

    def run_eval():
        with torch.cuda.stream(s1):
            for _ in range(n_eval):
                ev(e1)

    def run_tree():
        with torch.cuda.stream(s2):
            e2.reset_roots()
            e2.begin(n_waves, None, 0.0, 1)
            for _ in range(n_waves):
                e2.step()
                fake(e2)

    def timed(fns):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for f in fns:
            f()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    for _ in range(2):
        timed([run_eval]); timed([run_tree])
    te = min(timed([run_eval]) for _ in range(3))
    tt = min(timed([run_tree]) for _ in range(3))
    tb = min(timed([run_eval, run_tree]) for _ in range(3))
    tb2 = min(timed([run_tree, run_eval]) for _ in range(3))
    print(f"{args.net} {args.board} x {args.games}: evaluator x{n_eval} alone {te:.2f} ms, tree waves x{n_waves} alone {tt:.2f} ms, "
          f"together {tb:.2f} / {tb2:.2f} ms (serial would be {te + tt:.2f}; perfect overlap {max(te, tt):.2f})")
    e1.close(); e2.close()


if __name__ == "__main__":
    main()
