"""Pins oracle/dbaz_oracle.c against outputs recorded from the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_io import load, unhex
from oracle import oracle

GAMES = load("games")
MCTS = load("mcts")
MCTS_PENDING = load("mcts_pending")
SELFPLAY = load("selfplay")


def _cmp_state(rec, ref, where, full_hash=True):
    for k in ("board", "to_play", "just_played", "btc2", "hash1_x2", "result"):
        assert rec[k] == ref[k], (where, k, rec[k], ref[k])
    if full_hash:
        assert rec["hash0"] == ref["hash0"], where
    else:  # inside mcts.py moves are np.int64 and the reference's own hash wraps; low 32 bits are exact
        assert int(rec["hash0"]) & 0xFFFFFFFF == int(ref["hash0"]) & 0xFFFFFFFF, where


@pytest.mark.parametrize("gi", range(len(GAMES)))
def test_game_rules(gi):
    G = GAMES[gi]
    g = oracle.OracleGame(G["L"], G["C"])
    _cmp_state(g.record(), G["init"], "init")
    assert np.array_equal(g.valid_moves(), unhex(G["init_valid"], np.uint8).astype(bool))
    assert np.array_equal(g.features().ravel(), unhex(G["init_features"], np.int8))
    for pi, P in enumerate(G["plies"]):
        closed = g.play_(P["move"])
        assert [list(x) for x in closed] == P["closed"], (pi, closed)
        _cmp_state(g.record(), P, pi)
        assert np.array_equal(g.valid_moves(), unhex(P["valid"], np.uint8).astype(bool))
        assert np.array_equal(g.features().ravel().astype(np.int8), unhex(P["features"], np.int8))
    for a, raised in G["illegal"]:
        assert raised == 1
        with pytest.raises(ValueError):
            g.copy().play_(a)


def _check_root(r, ref, where):
    assert r["visits"].tolist() == ref["visits"], where
    assert np.array_equal(r["W"], unhex(ref["W"], np.float32)), where
    assert np.array_equal(r["priors"], unhex(ref["priors"], np.float64)), where
    assert r["root_N"] == ref["root_N"], where
    assert np.float32(r["root_W"]) == np.float32(ref["root_W"]), where
    assert r["stats"][:3] == ref["stats"][:3], (where, r["stats"], ref["stats"])
    assert np.float32(r["stats"][3]) == np.float32(ref["stats"][3]), where
    assert r["is_terminal"] == ref["is_terminal"] and r["is_expanded"] == ref["is_expanded"], where
    assert r["priors_f64"] == (ref["priors_dtype"] == "float64"), where
    if ref["root_N"] > 0 or not ref["is_expanded"]:
        # sign of never-visited children is +1 in both; compare only where it can matter
        vis = np.array(ref["visits"]) > 0
        assert np.array_equal(r["sign"][vis], np.array(ref["sign"])[vis]), where
    assert np.array_equal(r["ucb"], unhex(ref["ucb"], np.float64)), where
    _cmp_state(r["state"].record(), ref["state"], where, full_hash=False)


@pytest.mark.parametrize("si", range(len(MCTS) + len(MCTS_PENDING)))
def test_mcts_sessions(si):
    S = MCTS[si] if si < len(MCTS) else MCTS_PENDING[si - len(MCTS)]
    start = oracle.OracleGame(S["L"], S["C"])
    for m in S["pre_moves"]:
        start.play_(m)
    t = oracle.OracleTree(S["L"], S["C"], start.s, kind=S["kind"])
    assert len(S["steps"]) > 0
    for i, st in enumerate(S["steps"]):
        if st["op"] == "search":
            noise = unhex(st["noise"], np.float64) if "noise" in st else None
            t.search(st["num_reads"], cpuct=S["cpuct"], noise=noise, coeff=st["coeff"], max_pending=S.get("max_pending", 1))
        else:
            t.reroot(st["move"], st["reuse"])
        _check_root(t.root(S["cpuct"]), st["root"], (si, i, st["op"]))


def test_python_nn_seam_matches_builtin_fake():
    def nn(g):
        return g.fake_nn(0)
    a = oracle.OracleTree(3, 3, nn=nn)
    b = oracle.OracleTree(3, 3, kind=0)
    assert np.array_equal(a.search(200), b.search(200))


@pytest.mark.parametrize("gi", range(len(SELFPLAY)))
def test_selfplay_trajectory(gi):
    """SelfPlay.play_game (self_play.py:51-74) replayed on the oracle with the host-side
    RNG protocol of SURVEY 8c: per move the legacy MT19937 stream yields the Dirichlet
    draw, then the single uniform consumed by np.random.choice."""
    import math
    G = SELFPLAY[gi]
    rs = np.random.RandomState(G["seed"])
    t = oracle.OracleTree(G["L"], G["C"], kind=G["kind"])
    A = t.A
    temp = None
    moves, visits = [], []
    i = -1
    while not t.root()["is_terminal"]:
        i += 1
        if str(i) in G["temperature"]:
            temp = G["temperature"][str(i)]
        r = t.root()
        valid = r["state"].valid_moves()
        n = min(4 * math.factorial(int(valid.sum())), G["num_read"])
        alpha, coeff = G["noise"]
        noise = rs.dirichlet(np.ones(A) * alpha, 1).ravel() * valid
        vc = t.search(n, noise=noise, coeff=coeff)
        probs = (vc / vc.max()) ** (1 / temp)
        probs = probs / probs.sum()
        mv = int(rs.choice(A, 1, p=probs)[0])
        moves.append(mv)
        visits.append(vc.tolist())
        t.reroot(mv, True)
    assert moves == G["moves"]
    assert visits == G["visits"]
    assert t.root()["state"].result() == G["z"]
