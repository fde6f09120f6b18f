"""In-wave duplicates of the net's rows in self-play: trees that miss the eval cache on the SAME position in the same wave
are all evaluated (the table only learns a position when its evaluation comes back).  Plays a lock-step batch without CUDA
graphs behind an evaluator wrapper that counts unique (edges, boxes_to_close, to_play) keys among the rows of every wave.

  python tools/wave_dup_probe.py 8192        # round 2: 15.5 % of 11.2 M rows are duplicates within their wave
"""
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
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    eng = engine.Engine((3, 3), n_games=n, max_nodes=4096, device=dev, eval_cache=24)
    eng.set_mode(False, 4)
    torch.manual_seed(0)
    inner = FusedSimpleNN(SimpleNN(board=(3, 3)), eng, dtype=torch.bfloat16)
    tot = {"rows": 0, "uniq": 0, "waves": 0}

    class Counting:
        engine_launches = getattr(inner, "engine_launches", 0)

        def __call__(self, e):
            st = e.leaf_states[:e.n_games]
            sel = st[e._leaf_kind[:e.n_games] == 1]
            if sel.shape[0]:
                v = sel.view(torch.uint8).reshape(-1, 32)
                key = torch.cat([v[:, :16], v[:, 16:21]], 1)
                tot["rows"] += int(sel.shape[0])
                tot["uniq"] += int(torch.unique(key, dim=0).shape[0])
                tot["waves"] += 1
            inner(e)

    params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": bench.NOISE,
                                    "mcts": {"mcts_num_read": 800, "mcts_cpuct": (1.25, 19652), "temperature": {0: 1.0, 12: 0.02},
                                             "max_async_searches": 1}}})
    sp = self_play.BatchedSelfPlay(eng, Counting(), params, graph_waves=0, adaptive=False)
    info = sp.play_games_device(range(n), seed=3)
    print("games %d: %d simulations, %d rows for the net in %d waves, %d unique within their wave -> %.1f %% duplicates; %d cache hits"
          % (n, info["sims"], tot["rows"], tot["waves"], tot["uniq"], 100.0 * (1 - tot["uniq"] / max(1, tot["rows"])), info["cache_hits"]))
    eng.close()


if __name__ == "__main__":
    main()
