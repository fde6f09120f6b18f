"""Pins oracle/py_port.py (the Python restatement timed as the CPU baseline) against the fixtures
recorded from the real reference.  CPU only; the longest sessions are left to the C oracle."""
import asyncio
import warnings

import numpy as np
import pytest

from golden_io import load, unhex
from oracle import py_port as pp

MCTS = load("mcts")
SELFPLAY = load("selfplay")
GAMES = load("games")
SHORT = [i for i, S in enumerate(MCTS) if sum(st.get("num_reads", 0) for st in S["steps"]) <= 5000]


def _nn(kind):
    async def nn(board):
        return pp.fake_nn_eval(board, kind)
    return nn


@pytest.mark.parametrize("gi", range(0, len(GAMES), 3))
def test_board_rules(gi):
    G = GAMES[gi]
    pp.Board.configure(G["L"], G["C"])
    b = pp.Board()
    for P in G["plies"]:
        closed = b.apply(P["move"])
        assert [list(x) for x in closed] == P["closed"]
        assert bytes(b.cells.ravel().tolist()).hex() == P["board"]
        assert (b.to_play, -1 if b.just_played is None else b.just_played) == (P["to_play"], P["just_played"])
        assert [int(round(2 * x)) for x in b.need] == P["btc2"]
        assert (2 if b.result() is None else b.result()) == P["result"]
        assert np.array_equal(b.planes().ravel().astype(np.int8), unhex(P["features"], np.int8))
        assert str(int(b.key[0])) == P["hash0"]


@pytest.mark.parametrize("si", SHORT)
def test_mcts_sessions(si):
    warnings.filterwarnings("ignore")
    S = MCTS[si]
    pp.Board.configure(S["L"], S["C"])
    b = pp.Board()
    for m in S["pre_moves"]:
        b.apply(m)
    root = pp.new_root(b)
    if S["seed"] is not None:
        np.random.seed(S["seed"])
    for i, st in enumerate(S["steps"]):
        if st["op"] == "search":
            asyncio.run(pp.uct_search(root, st["num_reads"], _nn(S["kind"]), cpuct=tuple(S["cpuct"]), max_pending=1,
                                      dirichlet=(st["alpha"], st["coeff"])))
        else:
            root = pp.reroot(root, st["move"], st["reuse"])
        ref = st["root"]
        assert root.child_N.tolist() == ref["visits"], (si, i)
        assert np.array_equal(root.child_W, unhex(ref["W"], np.float32)), (si, i)
        assert np.array_equal(np.asarray(root.prior, dtype=np.float64), unhex(ref["priors"], np.float64)), (si, i)
        assert int(root.N_own) == ref["root_N"]
        stats = root.parent.stats()
        assert [int(x) for x in stats[:3]] == ref["stats"][:3], (si, i)


def test_selfplay_trajectory():
    warnings.filterwarnings("ignore")
    G = SELFPLAY[0]
    pp.Board.configure(G["L"], G["C"])
    np.random.seed(G["seed"])
    seq = asyncio.run(pp.play_game(_nn(G["kind"]), num_read=G["num_read"], noise=tuple(G["noise"]), max_pending=1))
    assert [int(n.move) for n in seq[1:]] == G["moves"]
    assert [n.child_N.tolist() for n in seq[:-1]] == G["visits"]
