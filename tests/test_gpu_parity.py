"""GPU parity: the CUDA engine (through the C ABI) against fixtures recorded from the real
reference and against the pinned CPU oracle on seeded inputs.  Bit-exact everywhere."""
import numpy as np
import pytest

from golden_io import load, unhex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dotsboxesaz_b200 import engine
    from oracle import oracle
    return engine, oracle


def _state_rec(eng, st_np):
    """packed state -> the record layout of the golden fixtures"""
    A = eng.A
    e = (int(st_np["edges"][0]) | (int(st_np["edges"][1]) << 64))
    plane = eng.rows * eng.cols
    board = np.zeros(A, np.uint8)
    for a in range(A):
        p, rem = divmod(a, plane)
        l, c = divmod(rem, eng.cols)
        pad = (c == eng.C) if p == 0 else (l == eng.L)
        board[a] = 255 if (e >> a) & 1 else (1 if pad else 0)
    return {"board": bytes(board.tolist()).hex(), "to_play": int(st_np["to_play"]), "just_played": int(st_np["just_played"]),
            "btc2": [int(st_np["btc2"][0]), int(st_np["btc2"][1])], "hash0": str(e)}


def test_game_rules_vs_reference_fixtures(mods):
    engine, _ = mods
    GAMES = load("games")
    by_board = {}
    for G in GAMES:
        by_board.setdefault((G["L"], G["C"]), []).append(G)
    for (L, C), games in by_board.items():
        eng = engine.Engine((L, C), n_games=1, max_nodes=4)
        n = len(games)
        st = eng.new_states(n)
        h = eng.states_to_numpy(st)
        for gi, G in enumerate(games):
            rec = _state_rec(eng, h[gi])
            for k in ("board", "to_play", "just_played", "btc2", "hash0"):
                assert rec[k] == G["init"][k]
        assert np.array_equal(eng.valid_moves(st).cpu().numpy()[0], unhex(games[0]["init_valid"], np.uint8).astype(bool))
        assert np.array_equal(eng.features(st).cpu().numpy()[0].ravel(), unhex(games[0]["init_features"], np.int8))
        max_p = max(len(G["plies"]) for G in games)
        for pi in range(max_p):
            mv = np.array([G["plies"][pi]["move"] if pi < len(G["plies"]) else -1 for G in games], np.int32)
            ncl, lc = eng.play(st, mv)
            ncl, lc = ncl.cpu().numpy(), lc.cpu().numpy()
            h = eng.states_to_numpy(st)
            res = eng.result(st).cpu().numpy()
            valid = eng.valid_moves(st).cpu().numpy()
            feats = {dt: eng.features(st, dt, cl).cpu().float().numpy()
                     for dt, cl in ((torch.int16, False), (torch.float32, False), (torch.bfloat16, True), (torch.float16, False))}
            for gi, G in enumerate(games):
                if pi >= len(G["plies"]):
                    assert ncl[gi] == -1  # move -1 is illegal: state untouched
                    continue
                P = G["plies"][pi]
                assert ncl[gi] == len(P["closed"]), (L, C, gi, pi)
                got = [[int(lc[gi][2 * j]), int(lc[gi][2 * j + 1])] for j in range(ncl[gi])]
                assert got == P["closed"]
                rec = _state_rec(eng, h[gi])
                for k in ("board", "to_play", "just_played", "btc2", "hash0"):
                    assert rec[k] == P[k], (L, C, gi, pi, k)
                assert int(res[gi]) == P["result"]
                assert np.array_equal(valid[gi], unhex(P["valid"], np.uint8).astype(bool))
                ref_f = unhex(P["features"], np.int8).astype(np.float32)
                for f in feats.values():
                    assert np.array_equal(f[gi].ravel(), ref_f)
        # illegal moves: the reference raises ValueError, the engine reports -1 and leaves the state alone
        before = eng.states_to_numpy(st)
        bad = np.array([G["illegal"][0][0] if G["illegal"] else -1 for G in games], np.int32)
        ncl, _ = eng.play(st, bad)
        assert (ncl.cpu().numpy() == -1).all()
        assert np.array_equal(eng.states_to_numpy(st), before)
        eng.close()


def _check_root(eng, ref, where):
    vis = eng.root_visits().cpu().numpy()[0]
    W, P, S, U = (x.cpu().numpy()[0] for x in eng.root_children())
    st, rW, q = (x.cpu().numpy()[0] for x in eng.tree_stats())
    assert vis.tolist() == ref["visits"], where
    assert np.array_equal(W, unhex(ref["W"], np.float32)), where
    assert np.array_equal(P, unhex(ref["priors"], np.float64)), where
    assert int(st[0]) == ref["root_N"], where
    assert np.float32(rW) == np.float32(ref["root_W"]), where
    assert [int(st[1]), int(st[2]), int(st[3])] == ref["stats"][:3], (where, st, ref["stats"])
    assert np.float32(q) == np.float32(ref["stats"][3]), where
    assert bool(st[4]) == ref["is_expanded"] and bool(st[5]) == ref["is_terminal"], where
    v = np.array(ref["visits"]) > 0
    assert np.array_equal(S[v], np.array(ref["sign"])[v]), where
    assert np.array_equal(U, unhex(ref["ucb"], np.float64)), where
    rec = _state_rec(eng, eng.states_to_numpy(eng.root_states())[0])
    for k in ("board", "to_play", "just_played", "btc2"):
        assert rec[k] == ref["state"][k], (where, k)


def test_mcts_sessions_vs_reference_fixtures(mods):
    """Every recorded UCT_search / init_mcts_tree session of the reference, one tree at a time: the sequential ones
    (max_pending_evals = 1) and the ones with 4..64 simulations in flight."""
    engine, oracle = mods
    MCTS = load("mcts") + load("mcts_pending")
    engines = {}
    for si, S in enumerate(MCTS):
        key = (S["L"], S["C"])
        if key not in engines:
            engines[key] = engine.Engine(key, n_games=1, max_nodes=8192, cpuct=S["cpuct"], max_pending=64)
        eng = engines[key]
        st = eng.new_states(1)
        for m in S["pre_moves"]:
            ncl, _ = eng.play(st, [m])
            assert int(ncl[0]) >= 0
        eng.reset_roots(st)
        ev = engine.FakeNetEvaluator(S["kind"])
        for i, step in enumerate(S["steps"]):
            if step["op"] == "search":
                noise = None
                if "noise" in step:
                    noise = torch.from_numpy(unhex(step["noise"], np.float64)).reshape(1, -1)
                eng.run_search(step["num_reads"], ev, noise=noise, coeff=step["coeff"], pending=S.get("max_pending", 1),
                               graph_waves=4 if si % 5 == 0 else 0)
            else:
                eng.advance_roots([step["move"]], reuse=step["reuse"])
            _check_root(eng, step["root"], (si, i, step["op"]))
        eng.status()
    for e in engines.values():
        e.close()


@pytest.mark.parametrize("board,n_games,sims,kind,pending", [((3, 3), 256, 800, 0, 1), ((3, 3), 128, 200, 1, 1), ((5, 5), 64, 300, 0, 1),
                                                             ((2, 2), 64, 100, 1, 1), ((4, 4), 32, 150, 0, 1),
                                                             ((3, 3), 128, 800, 0, 8), ((3, 3), 64, 800, 1, 64), ((5, 5), 32, 300, 0, 16)])
def test_batched_search_vs_oracle(mods, board, n_games, sims, kind, pending):
    """Many trees in lock-step from different seeded start positions, several moves with tree reuse;
    visit counts, W, priors and tree stats of every tree must equal the oracle's bit for bit."""
    engine, oracle = mods
    L, C = board
    eng = engine.Engine(board, n_games=n_games, max_nodes=4096, max_pending=pending)
    rng = np.random.RandomState(42)
    og = [oracle.OracleGame(L, C) for _ in range(n_games)]
    st = eng.new_states(n_games)
    n_edges = L * (C + 1) + C * (L + 1)
    for ply in range(max(1, n_edges // 3)):
        mv = np.full(n_games, -1, np.int32)
        for g in range(n_games):
            if ply < (g % (n_edges // 3 + 1)) and og[g].result() is None:
                legal = np.flatnonzero(og[g].valid_moves())
                mv[g] = rng.choice(legal)
                og[g].play_(int(mv[g]))
        eng.play(st, mv)
    trees = [oracle.OracleTree(L, C, og[g].s, kind=kind) for g in range(n_games)]
    eng.reset_roots(st)
    ev = engine.FakeNetEvaluator(kind)
    for move_i in range(4):
        reads = np.array([sims if g % 5 else sims // 2 for g in range(n_games)], np.int32)
        term = np.array([t.root()["is_terminal"] for t in trees])
        reads[term] = -1
        eng.run_search(torch.from_numpy(reads), ev, max_reads=sims, pending=pending, graph_waves=8 if move_i % 2 else 0)
        vis = eng.root_visits().cpu().numpy()
        W, P, S, U = (x.cpu().numpy() for x in eng.root_children())
        stt, rW, q = (x.cpu().numpy() for x in eng.tree_stats())
        moves = np.full(n_games, -1, np.int32)
        for g in range(n_games):
            if term[g]:
                continue
            ov = trees[g].search(int(reads[g]), max_pending=pending)
            r = trees[g].root()
            assert np.array_equal(vis[g], ov), (move_i, g)
            assert np.array_equal(W[g], r["W"]) and np.array_equal(P[g], r["priors"]), (move_i, g)
            assert [int(x) for x in stt[g][:4]] == [r["root_N"]] + r["stats"][:3], (move_i, g)
            assert np.float32(rW[g]) == np.float32(r["root_W"])
            # a move that is sometimes the most visited and sometimes an unvisited legal one
            legal = np.flatnonzero(r["state"].valid_moves())
            moves[g] = int(np.argmax(ov)) if g % 3 else int(legal[(g + move_i) % len(legal)])
            trees[g].reroot(int(moves[g]), True)
        eng.advance_roots(moves, reuse=True)
        s = eng.status()
        assert s["errors"] == 0
    eng.close()


def test_random_rollouts_vs_oracle(mods):
    engine, oracle = mods
    for (L, C), n in (((5, 5), 2048), ((3, 3), 1024), ((2, 3), 256)):
        eng = engine.Engine((L, C), n_games=1, max_nodes=4)
        st = eng.new_states(n)
        plies, moves = eng.random_rollout(st, seed=12345, game0=7, record_moves=True)
        plies, moves = plies.cpu().numpy(), moves.cpu().numpy()
        h = eng.states_to_numpy(st)
        res = eng.result(st).cpu().numpy()
        for g in range(0, n, 7):
            og = oracle.OracleGame(L, C)
            mv = og.random_rollout(12345, 7 + g)
            assert plies[g] == len(mv) and moves[g][:len(mv)].tolist() == mv, g
            rec = _state_rec(eng, h[g])
            orec = og.record()
            for k in ("board", "to_play", "just_played", "btc2", "hash0"):
                assert rec[k] == orec[k]
            assert int(res[g]) == orec["result"] and orec["result"] != 2
        eng.close()


def test_pool_exhaustion_is_reported(mods):
    engine, _ = mods
    eng = engine.Engine((3, 3), n_games=4, max_nodes=16)
    eng.reset_roots()
    eng.run_search(100, engine.FakeNetEvaluator(0))
    with pytest.raises(engine.EngineError):
        eng.status()
    eng.close()


def test_illegal_reroot_is_reported(mods):
    engine, _ = mods
    eng = engine.Engine((3, 3), n_games=2, max_nodes=64)
    eng.reset_roots()
    eng.run_search(10, engine.FakeNetEvaluator(0))
    eng.advance_roots([3, 0])  # action 3 is a padding cell on 3x3
    with pytest.raises(engine.EngineError):
        eng.status()
    eng.close()
