"""GPU: the reference-facing Python surface (BoxesState / mcts / SelfPlay / BatchedSelfPlay) against
fixtures recorded from the real reference.  The tests read like the reference's own usage."""
import asyncio
import warnings

import numpy as np
import pytest

from golden_io import load, unhex

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GAMES = load("games")
MCTS = load("mcts")
MCTS_PENDING = load("mcts_pending")
SELFPLAY = load("selfplay")


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dotsboxesaz_b200 import mcts, self_play, engine
    from dotsboxesaz_b200.dots_boxes.dots_boxes_game import BoxesState, nn_batch_builder
    from dotsboxesaz_b200.utils.utils import DotDict
    return dict(mcts=mcts, self_play=self_play, engine=engine, BoxesState=BoxesState, nn_batch_builder=nn_batch_builder,
                DotDict=DotDict)


def fake_nn_eval(state, kind):
    A = state.get_actions_size()
    h = int(state.get_hash()[0]) & 0xFFFFFFFF
    if kind == 0:
        raw = np.array([float((h * 2654435761 + i * 40503) % 1024) + 1 for i in range(A)], dtype=np.float32)
        return raw / raw.sum(), np.array([((h % 2001) - 1000) / 1000], dtype=np.float32)
    return (np.full(A, np.float32(1.0) / np.float32(A), dtype=np.float32),
            np.array([(((h * 31) % 5) - 2) / 2], dtype=np.float32))


def make_nn(kind, yielding=False):
    async def nn(state):
        if yielding:
            await asyncio.sleep(0)
        return fake_nn_eval(state, kind)
    return nn


def test_boxes_state_surface(api):
    BoxesState = api["BoxesState"]
    for G in [GAMES[0], GAMES[25], GAMES[34], GAMES[40]]:
        BoxesState.init_static_fields(((G["L"], G["C"]),))
        assert BoxesState.NB_ACTIONS == 2 * (G["L"] + 1) * (G["C"] + 1) and BoxesState.NB_BOXES == G["L"] * G["C"]
        s = BoxesState()
        assert s.hash == (0, 0) and s.just_played is None and s.to_play == 0 and s.get_result() is None
        assert bytes(s.board.ravel().tolist()).hex() == G["init"]["board"]
        states = [s]
        for P in G["plies"]:
            t = s.play(P["move"])                      # play() clones
            assert s.hash != t.hash or P["move"] is None
            closed = s.play_(P["move"])                # play_() mutates
            assert s == t and hash(s) == hash(t)
            assert [list(x) for x in closed] == P["closed"]
            assert bytes(s.board.ravel().tolist()).hex() == P["board"]
            assert s.to_play == P["to_play"] and (-1 if s.just_played is None else s.just_played) == P["just_played"]
            assert [int(round(2 * x)) for x in s.boxes_to_close] == P["btc2"]
            assert (2 if s.get_result() is None else s.get_result()) == P["result"]
            assert str(s.get_hash()[0]) == P["hash0"] and int(round(2 * s.get_hash()[1])) == P["hash1_x2"]
            assert np.array_equal(s.get_valid_moves(), unhex(P["valid"], np.uint8).astype(bool))
            assert s.get_valid_moves(as_indices=True) == np.flatnonzero(unhex(P["valid"], np.uint8)).tolist()
            f = s.get_features()
            assert f.dtype == np.int16 and f.shape == BoxesState.FEATURES_SHAPE
            assert np.array_equal(f.ravel().astype(np.int8), unhex(P["features"], np.int8))
            states.append(t)
        for a, _ in G["illegal"]:
            with pytest.raises(ValueError):
                s.play(a)
        batch = api["nn_batch_builder"](*[(x,) for x in states[1:4]])
        assert batch.shape == (len(states[1:4]), 3, G["L"] + 1, G["C"] + 1)
        assert np.array_equal(batch[0].ravel().astype(np.int8), unhex(G["plies"][0]["features"], np.int8))
        assert "To play" in repr(s)
    BoxesState.init_static_fields(((3, 3),))


@pytest.mark.parametrize("si", [0, 3, 10, 11, 12, 13, 14, 20, 33])
def test_uct_search_dropin(api, si):
    """create_root_uct_node / UCT_search(max_pending_evals=1) / init_mcts_tree with a Python async nn."""
    warnings.filterwarnings("ignore")
    m, BoxesState = api["mcts"], api["BoxesState"]
    S = MCTS[si]
    BoxesState.init_static_fields(((S["L"], S["C"]),))
    s = BoxesState()
    for mv in S["pre_moves"]:
        s.play_(mv)
    root = m.create_root_uct_node(s)
    if S["seed"] is not None:
        np.random.seed(S["seed"])
    nn = make_nn(S["kind"])
    steps = S["steps"][:12]
    for i, st in enumerate(steps):
        if st["op"] == "search":
            vis = asyncio.run(m.UCT_search(root, st["num_reads"], nn, cpuct=tuple(S["cpuct"]), max_pending_evals=1,
                                           dirichlet=(st["alpha"], st["coeff"])))
            assert vis.dtype == np.int32 and vis.tolist() == st["root"]["visits"], (si, i)
        else:
            prev = root
            root = m.init_mcts_tree(root, st["move"], reuse_tree=st["reuse"])
            assert prev.child_number_visits.tolist() == steps[i - 1]["root"]["visits"]  # frozen for sample extraction
        ref = st["root"]
        assert np.array_equal(root.child_total_value, unhex(ref["W"], np.float32))
        assert np.array_equal(root.child_priors, unhex(ref["priors"], np.float64))
        assert np.array_equal(root.children_ucb_score(), unhex(ref["ucb"], np.float64))
        assert root.number_visits == ref["root_N"] and np.float32(root.total_value) == np.float32(ref["root_W"])
        assert root.is_terminal == ref["is_terminal"] and root.is_expanded == ref["is_expanded"]
        ts = root.get_tree_stats()
        assert [int(ts.max_deepness), int(ts.tree_size), int(ts.terminal_count)] == ref["stats"][:3]
        assert np.float32(ts.q_value) == np.float32(ref["stats"][3])
        assert root.parent.get_tree_stats() == ts
    BoxesState.init_static_fields(((3, 3),))


@pytest.mark.parametrize("si", [0, 2, 4, 6])
def test_uct_search_dropin_with_pending_evals(api, si):
    """UCT_search(max_pending_evals=K) with a net that suspends once per call, as the reference's batching proxy does:
    K leaves are awaited concurrently per wave and the visit counts equal the reference's."""
    warnings.filterwarnings("ignore")
    m, BoxesState = api["mcts"], api["BoxesState"]
    S = MCTS_PENDING[si]
    BoxesState.init_static_fields(((S["L"], S["C"]),))
    s = BoxesState()
    for mv in S["pre_moves"]:
        s.play_(mv)
    root = m.create_root_uct_node(s)
    if S["seed"] is not None:
        np.random.seed(S["seed"])
    in_flight = {"now": 0, "max": 0}
    base = make_nn(S["kind"], yielding=True)

    async def nn(state):
        in_flight["now"] += 1
        in_flight["max"] = max(in_flight["max"], in_flight["now"])
        try:
            return await base(state)
        finally:
            in_flight["now"] -= 1
    for i, st in enumerate(S["steps"][:8]):
        if st["op"] == "search":
            vis = asyncio.run(m.UCT_search(root, st["num_reads"], nn, cpuct=tuple(S["cpuct"]), max_pending_evals=S["max_pending"],
                                           dirichlet=(st["alpha"], st["coeff"])))
            assert vis.tolist() == st["root"]["visits"], (si, i)
        else:
            root = m.init_mcts_tree(root, st["move"], reuse_tree=st["reuse"])
        assert np.array_equal(root.child_total_value, unhex(st["root"]["W"], np.float32))
    assert in_flight["max"] > 1  # the caller's net really saw concurrent requests
    BoxesState.init_static_fields(((3, 3),))


def _params(api, G):
    return api["DotDict"]({"self_play": {"reuse_mcts_tree": True, "noise": tuple(G["noise"]),
                                         "mcts": {"mcts_num_read": G["num_read"], "mcts_cpuct": (1.25, 19652),
                                                  "temperature": {int(k): v for k, v in G["temperature"].items()},
                                                  "max_async_searches": 1}}})


@pytest.mark.parametrize("gi", [0, 4, 6])
def test_selfplay_dropin(api, gi):
    """SelfPlay.play_game + get_datasets, seeded like the reference run that produced the fixture."""
    warnings.filterwarnings("ignore")
    G = SELFPLAY[gi]
    BoxesState = api["BoxesState"]
    BoxesState.init_static_fields(((G["L"], G["C"]),))
    sp = api["self_play"].SelfPlay(make_nn(G["kind"]), _params(api, G))
    np.random.seed(G["seed"])
    asyncio.run(sp.play_game(BoxesState(), G["seed"]))
    idx, seq, z = sp.played_games[0]
    assert [int(n.move) for n in seq[1:]] == G["moves"]
    assert [n.child_number_visits.tolist() for n in seq[:-1]] == G["visits"]
    assert z == G["z"]
    df = sp.get_datasets(3, with_features=True).reset_index()
    assert list(df.columns) == G["columns"]
    got = df.to_numpy(dtype=np.float64)
    ref = np.array(G["rows"], dtype=np.float64)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    BoxesState.init_static_fields(((3, 3),))


def test_batched_selfplay_matches_reference_games(api):
    """All recorded 3x3 / 100-sim reference games at once, in lock-step on one engine, each with its own
    legacy RNG stream: same moves, visit counts and dataset rows as the reference played sequentially."""
    games = [G for G in SELFPLAY if (G["L"], G["C"], G["num_read"], G["kind"]) == (3, 3, 100, 0)]
    assert len(games) >= 3
    eng = api["engine"].Engine((3, 3), n_games=len(games) + 1, max_nodes=2048)  # one idle slot on purpose
    cached = api["engine"].Engine((3, 3), n_games=len(games) + 1, max_nodes=2048, eval_cache=14)
    # plain wave loop eager / in CUDA graphs, then the production schedule (eval cache shared by the games and kept across
    # moves, in-kernel chains, compact rows, adaptive batches)
    for e, graph_waves in ((eng, 0), (eng, 8), (cached, 4)):
        bsp = api["self_play"].BatchedSelfPlay(e, api["engine"].FakeNetEvaluator(0), _params(api, games[0]), graph_waves=graph_waves)
        assert bsp.adaptive == (e is cached)
        played = bsp.play_games([G["seed"] for G in games], seeds=[G["seed"] for G in games])
        for (idx, moves, visits, z), G in zip(played, games):
            assert moves == G["moves"] and [v.tolist() for v in visits] == G["visits"] and z == G["z"]
        df = bsp.get_datasets(3, True).reset_index()
        ref_rows = np.concatenate([np.array(G["rows"], dtype=np.float64) for G in games])
        assert list(df.columns) == games[0]["columns"]
        assert np.array_equal(df.to_numpy(dtype=np.float64), ref_rows)
    assert cached.status()["cache_hits"] > 0
    eng.close()
    cached.close()


@pytest.mark.parametrize("mode", ["device", "async"])
def test_device_resident_selfplay_is_valid_play(api, mode):
    """play_games_device() (no host sync between moves, device RNG) and play_games_async() (every game at its own pace).
    Not seed-identical to the reference, so check what is invariant: every recorded move is legal, games end exactly when
    the oracle says they do, the result and z signs agree with the oracle, visit totals equal the simulation budget plus
    the reused subtree, and the recorded features are those of the position searched."""
    from oracle import oracle
    eng = api["engine"].Engine((3, 3), n_games=48, max_nodes=2048, eval_cache=14)  # the production schedule
    G = SELFPLAY[0]
    bsp = api["self_play"].BatchedSelfPlay(eng, api["engine"].FakeNetEvaluator(0), _params(api, G), graph_waves=8 if mode == "device" else 2)
    assert bsp.adaptive
    info = (bsp.play_games_device if mode == "device" else bsp.play_games_async)(range(48), seed=5)
    assert info["errors"] == 0
    h = bsp._device_hist
    moves = torch.stack(h["moves"]).cpu().numpy()       # [n_moves, n]
    active = torch.stack(h["active"]).cpu().numpy()
    visits = torch.stack(h["visits"]).cpu().numpy()
    res = h["result"].cpu().numpy()
    planes, pi, z, slot, mi = (x.cpu().numpy() for x in bsp.device_samples())
    assert np.allclose(pi.sum(1), 1.0) and set(np.unique(z)).issubset({-1.0, 0.0, 1.0})
    assert planes.shape[0] == active.sum()
    for g in range(48):
        og = oracle.OracleGame(3, 3)
        carried = 0
        for m in range(moves.shape[0]):
            if og.result() is not None:
                assert not active[m, g] and moves[m, g] == -1
                continue
            assert active[m, g]
            k = int(og.valid_moves().sum())
            import math
            n_reads = min(4 * math.factorial(k), G["num_read"])
            assert visits[m, g].sum() == n_reads + carried
            assert og.valid_moves()[moves[m, g]]
            rows = np.flatnonzero((slot == g) & (mi == m))
            assert len(rows) == 1 and np.array_equal(planes[rows[0]].ravel(), og.features().ravel())
            carried = max(int(visits[m, g][moves[m, g]]) - 1, 0) if visits[m, g][moves[m, g]] > 0 else 0
            og.play_(int(moves[m, g]))
        assert og.result() == int(res[g])
    # the vectorised DataFrame builder against the row-by-row one (same columns, dtypes, index and order)
    import pandas as pd
    for gen in (3, [2, 5]):
        fast = bsp.get_datasets(gen, True)
        if not bsp.rows:
            bsp._rows_from_device(True)
        slow = api["self_play"]._dataset(bsp.rows, gen, True)
        bsp.rows = []
        pd.testing.assert_frame_equal(fast, slow, check_exact=True)
    eng.close()


def test_batching_proxy_in_front_of_a_torch_net(api):
    """The reference's stack: UCT_search(max_pending_evals=K) -> AsyncBatchedProxy -> NeuralNetWrapper(model).  The
    proxy must see real batches because the drop-in awaits its K leaves concurrently."""
    warnings.filterwarnings("ignore")
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    from dotsboxesaz_b200.nn import NeuralNetWrapper
    from dotsboxesaz_b200.utils.proxies import AsyncBatchedProxy
    m, BoxesState, DotDict = api["mcts"], api["BoxesState"], api["DotDict"]
    BoxesState.init_static_fields(((3, 3),))
    torch.manual_seed(0)
    wrapper = NeuralNetWrapper(SimpleNN(board=(3, 3)), DotDict({"nn": {"pytorch_device": "cuda:0"}}))

    async def main():
        proxy = AsyncBatchedProxy(wrapper, batch_size=16, timeout=0.01, batch_builder=api["nn_batch_builder"], cache_size=1000)
        task = asyncio.ensure_future(proxy.run())
        root = m.create_root_uct_node(BoxesState())
        vis = await m.UCT_search(root, 96, proxy, max_pending_evals=16, dirichlet=(0.0, 0.0))
        task.cancel()
        return vis, proxy
    vis, proxy = asyncio.run(main())
    assert int(vis.sum()) == 96
    assert proxy.n_batches < proxy.n_evals  # batches of more than one leaf were dispatched


def test_elo_arena_identical_nets_split_wins(api):
    """compute_elo with the same weights on both sides: every game is decided, colours alternate, and both ratings
    move by equal and opposite amounts."""
    from dotsboxesaz_b200 import configuration
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    api["BoxesState"].init_static_fields(((3, 3),))
    params = configuration.simple
    elo = api["DotDict"]({"n_games": 64, "seed": 3, "self_play_override": {"reuse_mcts_tree": False, "noise": (0.0, 0.0),
                                                                          "mcts": {"mcts_num_read": 40}}})
    torch.manual_seed(1)
    net = SimpleNN(board=(3, 3))
    e0, e1, share = api["self_play"].compute_elo(elo, [params, params], [1, 2], (1200, 1200), models=[net, net])
    assert abs((e0 - 1200) + (e1 - 1200)) < 1e-9
    assert 0.0 <= share <= 1.0


def test_time_limited_player(api):
    """players.py:55-69 on the engine: `AZPlayer.serve_once` searches until the time limit and answers with the most
    visited legal move; the deterministic fake net makes the answer checkable against a fixed-budget search."""
    warnings.filterwarnings("ignore")
    from dotsboxesaz_b200 import configuration, players
    m, BoxesState = api["mcts"], api["BoxesState"]
    BoxesState.init_static_fields(((3, 3),))
    params = configuration.simple
    player = players.AZPlayer(params, 0.3, None, None)
    assert params.self_play.mcts.temperature == {0: 1e-5}
    state = BoxesState()
    for mv in (0, 4, 16):
        state.play_(mv)
    sims = [0]

    async def nn(gs):
        sims[0] += 1
        h = gs.get_hash()[0] & 0xFFFFFFFF
        raw = np.array([float((h * 2654435761 + i * 40503) % 1024) + 1 for i in range(32)], dtype=np.float32)
        return raw / raw.sum(), np.array([((h % 2001) - 1000) / 1000], dtype=np.float32)
    import time
    t0 = time.time()
    move = asyncio.run(player.serve_once(nn, state, 0.3))
    dt = time.time() - t0
    assert 0.25 < dt < 3.0 and sims[0] > 50
    assert move is not None and state.get_valid_moves()[int(move)]
    # a terminal position is answered with None (nothing is visited below a terminal root)
    end = BoxesState()
    while end.get_result() is None:
        end.play_(int(end.get_valid_moves(as_indices=True)[0]))
    assert asyncio.run(player.serve_once(nn, end, 0.05)) is None


def _cmp_node(view, rec, where):
    """A node view of the drop-in against the reference's node (tests/golden/treewalk.json.gz): per-child arrays, own
    N / W, UCB scores, flags, state -- bit for bit."""
    from golden_io import unhex
    assert view.child_number_visits.tolist() == rec["visits"], where
    assert np.array_equal(np.asarray(view.child_total_value, dtype=np.float32), unhex(rec["W"], np.float32)), where
    if rec["is_expanded"]:
        assert np.array_equal(np.asarray(view.child_priors, dtype=np.float64), unhex(rec["priors"], np.float64)), where
        assert np.array_equal(np.asarray(view.children_ucb_score(), dtype=np.float64), unhex(rec["ucb"], np.float64)), where
        assert np.asarray(view.child_player_changed).tolist() == rec["sign"], where
    assert int(view.number_visits) == rec["N"] and np.float32(view.total_value) == np.float32(rec["own_W"]), where
    assert bool(view.is_terminal) == rec["is_terminal"] and bool(view.is_expanded) == rec["is_expanded"], where
    gs, st = view.game_state, rec["state"]
    assert bytes(gs.board.ravel().tolist()).hex() == st["board"] and gs.to_play == st["to_play"], where


@pytest.mark.parametrize("ti", [0, 1, 2, 3])
def test_tree_walk_from_python(api, ti, capsys):
    """UCTNode.children / print_mcts_tree (mcts.py:50-60,247-272): the tree below the root as the reference's Python
    objects expose it -- root, children and grandchildren after a search and after a re-root with reuse -- read from the
    engine's node pool (dbaz_search_node)."""
    warnings.filterwarnings("ignore")
    from golden_io import load
    T = load("treewalk")[ti]
    m, BoxesState = api["mcts"], api["BoxesState"]
    BoxesState.init_static_fields(((T["L"], T["C"]),))
    try:
        s = BoxesState()
        for mv in T["pre_moves"]:
            s.play_(int(mv))
        root = m.create_root_uct_node(s)
        nn = make_nn(T["kind"])
        for phase in ("first", "second"):
            asyncio.run(m.UCT_search(root, T["num_reads"], nn, max_pending_evals=1, dirichlet=(0.0, 0.0)))
            rec = T[phase]
            _cmp_node(root, rec, (ti, phase, "root"))
            kids = root.children
            assert sorted(kids) == sorted(int(a) for a in rec["children"])
            for a, child in kids.items():
                crec = rec["children"][str(a)]
                assert child.move == a and child.parent is root
                _cmp_node(child, crec, (ti, phase, a))
                gk = child.children
                assert sorted(gk) == sorted(int(b) for b in crec["children"])
                for b, g in gk.items():
                    _cmp_node(g, crec["children"][str(b)], (ti, phase, a, b))
            if phase == "first":
                m.print_mcts_tree(root, max_level=1)
                out = capsys.readouterr().out
                assert out.count("child visits") == 1 + len(kids)
                stale = next(iter(kids.values()))
                root = m.init_mcts_tree(root, T["reroot_move"], reuse_tree=True)
                with pytest.raises(RuntimeError):
                    _ = stale.number_visits  # node indices changed with the re-root
    finally:
        BoxesState.init_static_fields(((3, 3),))
