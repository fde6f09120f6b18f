"""Short, deterministic workloads for ncu captures (profiles/): one target per invocation.
    python tools/ncu_targets.py tower      # k_resnet_tower: 15984 boards (6 full waves of 148 tiles), 40 stages + head
    python tools/ncu_targets.py rollout    # k_game_rollout: 2^20 random 5x5 playouts
    python tools/ncu_targets.py selfplay   # the asynchronous self-play loop, 3x3, 4096 games (launch list: k_advance_roots etc.)
    python tools/ncu_targets.py stem       # k_nn_stem_mma at 4096 leaves x 256 channels and 16384 x 64 (tiles)
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dotsboxesaz_b200 import engine  # noqa: E402


def tower():
    from dotsboxesaz_b200.nn import tower_pack
    dev = torch.device("cuda:0")
    eng = engine.Engine((5, 5), n_games=64, max_nodes=64, device=dev)
    g = eng.tower_geometry()
    n, S, hc = g["nb"] * 148 * 6, 40, 32
    torch.manual_seed(0)
    w3 = torch.randn(S, 64, 64, 3, 3, device=dev) * 0.045
    b3 = torch.randn(S, 64, device=dev) * 0.1
    packed, bias = tower_pack(w3, b3, torch.randn(hc, 64, device=dev) * 0.2, torch.randn(hc, device=dev) * 0.1)
    x = torch.rand(n, 6, 6, 64, device=dev).to(torch.bfloat16)
    tiles = eng.tower_tiles(n)
    eng.tower_planarize(x, tiles)
    out = torch.empty((n, 6, 6, hc), dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        eng.tower(tiles, packed, bias, S, hc, out)
    torch.cuda.synchronize()
    print("tower ok", float(out.float().abs().mean()))
    eng.close()


def rollout():
    dev = torch.device("cuda:0")
    eng = engine.Engine((5, 5), n_games=8, max_nodes=16, device=dev)
    st = eng.new_states(1 << 20)
    for _ in range(2):
        s2 = st.clone()
        plies = eng.random_rollout(s2, seed=0)
    torch.cuda.synchronize()
    print("rollout ok", int(plies.sum()))
    eng.close()


def selfplay(n=4096):
    from dotsboxesaz_b200 import self_play
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    from dotsboxesaz_b200.nn import FusedSimpleNN
    from dotsboxesaz_b200.utils.utils import DotDict
    dev = torch.device("cuda:0")
    eng = engine.Engine((3, 3), n_games=n, max_nodes=4096, device=dev, eval_cache=20 if n <= 4096 else 24)
    torch.manual_seed(0)
    ev = FusedSimpleNN(SimpleNN(board=(3, 3)), eng)
    params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": (0.8, 0.25),
                                    "mcts": {"mcts_num_read": 800, "mcts_cpuct": (1.25, 19652), "temperature": {0: 1.0, 12: 0.02}, "max_async_searches": 1}}})
    sp = self_play.BatchedSelfPlay(eng, ev, params, graph_waves=16, adaptive=True)
    info = sp.play_games_async(range(n), seed=1)
    planes = sp.device_samples()[0]
    torch.cuda.synchronize()
    print("selfplay ok", info["sims"], planes.shape[0])
    eng.close()


def stem():
    from dotsboxesaz_b200.nn import _stem_mma_table
    dev = torch.device("cuda:0")
    for board, cout, n in (((3, 3), 256, 4096), ((5, 5), 64, 16384)):
        eng = engine.Engine(board, n_games=8, max_nodes=16, device=dev)
        st = eng.new_states(n)
        conv = torch.nn.Conv2d(3, cout, 3, padding=1).to(dev)
        packed = eng.nn_stem_mma_pack(_stem_mma_table(conv).to(torch.bfloat16))
        out = torch.empty((n, eng.rows, eng.cols, cout), dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            eng.nn_stem_mma(st, packed, out)
        if cout == 64:
            tiles = eng.tower_tiles(n)
            for _ in range(3):
                eng.nn_stem_mma_tiles(st, packed, tiles)
        torch.cuda.synchronize()
        eng.close()
    print("stem ok")


if __name__ == "__main__":
    {"tower": tower, "rollout": rollout, "selfplay": selfplay, "selfplay32k": lambda: selfplay(32768), "stem": stem}[sys.argv[1]]()
