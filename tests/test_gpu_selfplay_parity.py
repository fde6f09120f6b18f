"""The paths that produce BASELINE's games/hour -- BatchedSelfPlay.play_games_device / play_games_async (device RNG,
production schedule: eval cache, in-kernel chains, compact rows, adaptive batch ladder, CUDA graphs) -- and full
config-1 games with host RNG, replayed move by move on the CPU oracle with the same noise and the same sampled moves:
visit counts and tree statistics of EVERY search of EVERY game must be bit-equal.
Reference: self_play.py:27-74 (get_next_move / play_game), mcts.py:183-244 (UCT_search), mcts.py:163-180 (init_mcts_tree).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TEMPERATURE = {0: 1.0, 12: 0.02}
NOISE = (0.8, 0.25)
CPUCT = (1.25, 19652)


def _params(num_read, noise=NOISE, pending=1):
    from dotsboxesaz_b200.utils.utils import DotDict
    return DotDict({"self_play": {"reuse_mcts_tree": True, "noise": noise,
                                  "mcts": {"mcts_num_read": num_read, "mcts_cpuct": CPUCT, "temperature": dict(TEMPERATURE),
                                           "max_async_searches": pending}}})


def _n_searches(k, num_read):
    return min(4 * math.factorial(k), num_read) if k < 8 else num_read   # self_play.py:64-65


def _replay_device_history(bsp, board, num_read, games):
    """Every game of a device-resident batch again on oracle.OracleTree: same roots, same pre-drawn noise, same moves."""
    from oracle import oracle
    h = bsp._device_hist
    L, C = board
    moves = torch.stack(h["moves"]).cpu().numpy()       # [n_moves, n]
    active = torch.stack(h["active"]).cpu().numpy()
    visits = torch.stack(h["visits"]).cpu().numpy()
    stats = torch.stack(h["stats"]).cpu().numpy()       # root_N, max_deepness, tree_size, terminal_count, ...
    q = torch.stack(h["q"]).cpu().numpy()
    noise = h["noise"].cpu().numpy() if h["noise"] is not None else None
    res = h["result"].cpu().numpy()
    searches = sims = 0
    for g in games:
        tree = oracle.OracleTree(L, C)
        og = oracle.OracleGame(L, C)
        for m in range(moves.shape[0]):
            if og.result() is not None:
                assert not active[m, g] and moves[m, g] == -1
                continue
            assert active[m, g], (g, m)
            valid = og.valid_moves()
            reads = _n_searches(int(valid.sum()), num_read)
            ov = tree.search(reads, cpuct=CPUCT, noise=None if noise is None else noise[m, g] * valid, coeff=NOISE[1] if noise is not None else 0.0)
            assert np.array_equal(visits[m, g], ov), ("visit counts", g, m, visits[m, g], ov)
            r = tree.root()
            assert int(stats[m, g, 0]) == r["root_N"]
            assert [int(stats[m, g, 1]), int(stats[m, g, 2]), int(stats[m, g, 3])] == r["stats"][:3], ("tree stats", g, m)
            assert np.float32(q[m, g]) == np.float32(r["stats"][3]), ("root q", g, m)
            searches += 1
            sims += reads
            mv = int(moves[m, g])
            assert valid[mv]
            tree.reroot(mv, reuse=True)
            og.play_(mv)
        assert og.result() == int(res[g])
    return searches, sims


@pytest.mark.parametrize("mode", ["device", "async"])
@pytest.mark.parametrize("board,n,num_read,check", [((3, 3), 256, 800, 96), ((5, 5), 64, 200, 24)])
def test_device_selfplay_replays_bit_exact_on_the_oracle(mode, board, n, num_read, check):
    """The games/hour path: whole batches of self-play games (3x3 @ 800 sims/move, 5x5 @ 200) through the production
    schedule with the deterministic fake net, then `check` games replayed on the oracle with the recorded device noise
    (noise_all[move, game]) and the sampled moves."""
    from dotsboxesaz_b200 import engine, self_play
    eng = engine.Engine(board, n_games=n, max_nodes=4096 if board == (3, 3) else 6144, eval_cache=16)
    try:
        bsp = self_play.BatchedSelfPlay(eng, engine.FakeNetEvaluator(0), _params(num_read), graph_waves=8 if mode == "device" else 4)
        assert bsp.adaptive and bsp.pending == 1
        info = (bsp.play_games_device if mode == "device" else bsp.play_games_async)(range(n), seed=11)
        assert info["errors"] == 0 and info["cache_hits"] > 0
        games = list(range(0, n, max(1, n // check)))[:check]
        searches, sims = _replay_device_history(bsp, board, num_read, games)
        assert searches >= 10 * len(games)
    finally:
        eng.close()


def test_config1_full_games_host_rng_vs_oracle():
    """64 fully-checked GAMES on 3x3 at 800 sims/move, Dirichlet (0.8, 0.25), temperature
    {0: 1.0, 12: 0.02}, tree reuse, one legacy-MT19937 stream per game (seed_g = base + g; per move: the Dirichlet draw,
    then the uniform of np.random.choice -- the reference's order), production schedule on.  The oracle plays the same
    games from the same seeds on its own: moves, visit counts of every search and results must be identical."""
    from dotsboxesaz_b200 import engine, self_play
    from oracle import oracle
    n, num_read, base = 64, 800, 4000
    eng = engine.Engine((3, 3), n_games=n, max_nodes=4096, eval_cache=16)
    try:
        bsp = self_play.BatchedSelfPlay(eng, engine.FakeNetEvaluator(0), _params(num_read), graph_waves=8)
        assert bsp.adaptive
        played = bsp.play_games(range(n), seeds=[base + g for g in range(n)])
        assert len(played) == n
        total = 0
        for g, (idx, moves, visits, z) in enumerate(played):
            rs = np.random.RandomState(base + g)
            tree, og = oracle.OracleTree(3, 3), oracle.OracleGame(3, 3)
            temperature, m = None, 0
            while og.result() is None:
                if m in TEMPERATURE:
                    temperature = TEMPERATURE[m]
                valid = og.valid_moves()
                noise = rs.dirichlet(np.ones(og.A) * NOISE[0], 1).ravel() * valid
                ov = tree.search(_n_searches(int(valid.sum()), num_read), cpuct=CPUCT, noise=noise, coeff=NOISE[1])
                assert np.array_equal(visits[m], ov), ("visit counts", g, m)
                probs = self_play._apply_temperature(ov, temperature)
                mv = int(rs.choice(og.A, 1, p=probs)[0])
                assert mv == moves[m], ("move", g, m)
                tree.reroot(mv, reuse=True)
                og.play_(mv)
                m += 1
                total += 1
            assert m == len(moves) and og.result() == z
        assert total > 15 * n
    finally:
        eng.close()


def test_reference_selfplay_fixture_at_800_sims():
    """tests/golden/selfplay.json.gz entry 3: a SelfPlay.play_game trajectory recorded from the real reference at 800
    sims/move (fake net kind 1, Dirichlet noise, temperature schedule); the batched engine path must reproduce its moves,
    visit counts, result and dataset rows from the same seed -- through the production schedule."""
    from golden_io import load
    from dotsboxesaz_b200 import engine, self_play
    from dotsboxesaz_b200.utils.utils import DotDict
    fixtures = [G for G in load("selfplay") if G["num_read"] == 800]
    assert fixtures, "no 800-sims fixture"
    for G in fixtures:
        eng = engine.Engine((G["L"], G["C"]), n_games=2, max_nodes=4096, eval_cache=14)  # one idle slot on purpose
        try:
            params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": tuple(G["noise"]),
                                            "mcts": {"mcts_num_read": G["num_read"], "mcts_cpuct": CPUCT,
                                                     "temperature": {int(k): v for k, v in G["temperature"].items()},
                                                     "max_async_searches": 1}}})
            bsp = self_play.BatchedSelfPlay(eng, engine.FakeNetEvaluator(G["kind"]), params, graph_waves=8)
            played = bsp.play_games([G["seed"]], seeds=[G["seed"]])
            idx, moves, visits, z = played[0]
            assert moves == G["moves"] and [v.tolist() for v in visits] == G["visits"] and z == G["z"]
            df = bsp.get_datasets(3, True).reset_index()
            assert list(df.columns) == G["columns"]
            assert np.array_equal(df.to_numpy(dtype=np.float64), np.array(G["rows"], dtype=np.float64))
        finally:
            eng.close()
