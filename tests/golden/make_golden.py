#!/usr/bin/env python
"""Generate tests/golden/*.json by running the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The fixtures pin the oracle (oracle/dbaz_oracle.c) and, through it and directly,
the CUDA engine.  Everything recorded here is an output of the reference's own
code: dots_boxes/dots_boxes_game.py (BoxesState), mcts.py (UCT_search,
init_mcts_tree, TreeRoot.get_tree_stats) and self_play.py (SelfPlay.play_game,
get_datasets).  The only things injected are deterministic fake NNs through the
`async_nn` seam (mcts.py:187) and `np.random.seed`.
"""
import asyncio
import json
import os
import random
import sys
import warnings

REF = os.environ.get("DBAZ_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from dots_boxes.dots_boxes_game import BoxesState  # noqa: E402
import mcts  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CPUCT = (1.25, 19652)


def set_board(L, C):
    BoxesState.init_static_fields(((L, C),))


# ---------------------------------------------------------------- fake NNs
def fake_nn_eval(state, kind):
    """kind 0: hash-seeded priors/values (SURVEY 8a KAT).  kind 1: uniform prior
    1/A and a 5-level value -- maximises exact ties (first index must win)."""
    A = state.get_actions_size()
    h = int(state.get_hash()[0]) & 0xFFFFFFFF
    if kind == 0:
        raw = np.array([float((h * 2654435761 + i * 40503) % 1024) + 1 for i in range(A)], dtype=np.float32)
        p = raw / raw.sum()
        v = np.array([((h % 2001) - 1000) / 1000], dtype=np.float32)
    else:
        p = np.full(A, np.float32(1.0) / np.float32(A), dtype=np.float32)
        v = np.array([(((h * 31) % 5) - 2) / 2], dtype=np.float32)
    return p, v


def make_nn(kind, yielding=False):
    """yielding: suspend once, like a real batched net would; with max_pending_evals = K the reference's event loop
    then runs the simulations in deterministic waves of K select_leaf()s followed by their K backups."""
    async def nn(state):
        if yielding:
            await asyncio.sleep(0)
        return fake_nn_eval(state, kind)
    return nn


def f(x):
    return float(x)


def hx(arr, dtype):
    """Arrays are stored as little-endian hex of `dtype` (exact, compact)."""
    return np.ascontiguousarray(np.asarray(arr), dtype=dtype).tobytes().hex()


def state_record(s):
    return {
        "board": bytes(s.board.ravel().tolist()).hex(),
        "to_play": int(s.to_play),
        "just_played": -1 if s.just_played is None else int(s.just_played),
        "btc2": [int(round(2 * b)) for b in s.boxes_to_close],
        "hash0": str(int(s.hash[0])),
        "hash1_x2": int(round(2 * s.hash[1])),
        "result": 2 if s.get_result() is None else int(s.get_result()),
    }


# ------------------------------------------------------------------- games
def gen_games():
    out = []
    rnd = random.Random(1234)
    for (L, C), n_games in (((3, 3), 24), ((5, 5), 10), ((2, 2), 8), ((4, 4), 6), ((2, 3), 8), ((1, 1), 2), ((3, 5), 4)):
        set_board(L, C)
        for gi in range(n_games):
            s = BoxesState()
            rec = {"L": L, "C": C, "init": state_record(s), "plies": [],
                   "init_valid": hx(s.get_valid_moves(), np.uint8),
                   "init_features": hx(s.get_features().ravel(), np.int8)}
            play_past_terminal = gi % 2 == 0
            while True:
                legal = s.get_valid_moves(as_indices=True)
                if not legal or (s.get_result() is not None and not play_past_terminal):
                    break
                mv = rnd.choice(legal)
                closed = s.play_(mv)
                r = state_record(s)
                r["move"] = int(mv)
                r["closed"] = [[int(a), int(b)] for a, b in closed]
                r["valid"] = hx(s.get_valid_moves(), np.uint8)
                r["features"] = hx(s.get_features().ravel(), np.int8)
                assert s.get_features().dtype.itemsize <= 2
                rec["plies"].append(r)
            # an illegal move must raise ValueError (dots_boxes_game.py:63-65)
            bad = [a for a in range(s.get_actions_size()) if not s.get_valid_moves()[a]]
            illegal = []
            for a in bad[:4]:
                try:
                    s.play(a)
                    illegal.append([a, 0])
                except ValueError:
                    illegal.append([a, 1])
            rec["illegal"] = illegal
            out.append(rec)
    return out


# -------------------------------------------------------------------- mcts
def root_record(root):
    st = root.get_tree_stats()
    tv = root.total_value
    tv = float(np.asarray(tv).ravel()[0])
    return {
        "visits": [int(x) for x in root.child_number_visits],
        "W": hx(root.child_total_value, np.float32),
        "priors": hx(root.child_priors, np.float64),
        "priors_dtype": str(np.asarray(root.child_priors).dtype),
        "sign": [int(x) for x in root.child_player_changed],
        "root_N": int(root.number_visits),
        "root_W": tv,
        "stats": [int(st.max_deepness), int(st.tree_size), int(st.terminal_count), f(st.q_value)],
        "ucb": hx(root.children_ucb_score(), np.float64),
        "state": state_record(root.game_state),
        "is_terminal": bool(root.is_terminal),
        "is_expanded": bool(root.is_expanded),
    }


class NoiseTap:
    """Records what np.random.dirichlet returned (the reference multiplies it by
    the legal mask afterwards, mcts.py:222-223)."""

    def __init__(self):
        self.last = None
        self._orig = np.random.dirichlet

    def __enter__(self):
        def tap(*a, **k):
            r = self._orig(*a, **k)
            self.last = np.array(r).ravel().copy()
            return r
        np.random.dirichlet = tap
        return self

    def __exit__(self, *e):
        np.random.dirichlet = self._orig


def run_session(L, C, pre_moves, script, kind, seed=None, max_pending=1):
    """script: list of ("search", num_reads, (alpha, coeff)) | ("reroot", move|"argmax", reuse)."""
    set_board(L, C)
    s = BoxesState()
    for m in pre_moves:
        s.play_(int(m))
    root = mcts.create_root_uct_node(s)
    nn = make_nn(kind, yielding=max_pending > 1)
    if seed is not None:
        np.random.seed(seed)
    steps = []
    with NoiseTap() as tap:
        for op in script:
            if root.is_terminal and op[0] == "reroot":
                break
            if op[0] == "search":
                _, n, dirichlet = op
                tap.last = None
                asyncio.run(mcts.UCT_search(root, n, nn, cpuct=CPUCT, max_pending_evals=max_pending, dirichlet=dirichlet))
                rec = {"op": "search", "num_reads": n, "alpha": dirichlet[0], "coeff": dirichlet[1]}
                if tap.last is not None:
                    rec["noise"] = hx(tap.last * root.game_state.get_valid_moves(), np.float64)
                rec["root"] = root_record(root)
                steps.append(rec)
            else:
                _, mv, reuse = op
                if mv == "argmax":
                    mv = int(np.argmax(root.child_number_visits))
                root = mcts.init_mcts_tree(root, int(mv), reuse_tree=reuse)
                steps.append({"op": "reroot", "move": int(mv), "reuse": bool(reuse), "root": root_record(root)})
    return {"L": L, "C": C, "pre_moves": [int(m) for m in pre_moves], "kind": kind, "seed": seed,
            "cpuct": list(CPUCT), "steps": steps, "max_pending": max_pending}


def full_game_script(n, dirichlet=(0.0, 0.0), max_moves=80, reuse=True):
    sc = []
    for _ in range(max_moves):
        sc.append(("search", n, dirichlet))
        sc.append(("reroot", "argmax", reuse))
    return sc


def load_csv_positions():
    rows = []
    with open(os.path.join(REF, "test", "test_boards.csv")) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#") or line.startswith("id;"):
                continue
            i, moves, nxt, z = line.split(";")
            rows.append((int(i), [int(x) for x in moves.split()], [int(x) for x in nxt.split()], z))
    return rows


def gen_mcts():
    sessions = []
    # KAT from SURVEY 8a + full argmax games with tree reuse
    sessions.append(run_session(3, 3, [], full_game_script(800), 0))
    sessions.append(run_session(3, 3, [5, 17], full_game_script(800), 1))
    sessions.append(run_session(3, 3, [0, 4, 16], full_game_script(300, reuse=False), 0))
    sessions.append(run_session(5, 5, [], full_game_script(800), 0))
    sessions.append(run_session(5, 5, [1, 40, 7], full_game_script(150), 1))
    sessions.append(run_session(2, 2, [], full_game_script(200), 0))
    sessions.append(run_session(2, 2, [], full_game_script(200), 1))
    sessions.append(run_session(4, 4, [], full_game_script(120), 0))
    sessions.append(run_session(2, 3, [], full_game_script(200), 1))
    sessions.append(run_session(3, 5, [2], full_game_script(100), 0))
    # the 3-sim KAT (uniform prior, root N=4)
    sessions.append(run_session(3, 3, [], [("search", 3, (0.0, 0.0))], 1))
    # Dirichlet path (host-supplied noise), repeated searches on one root, mixed reuse
    sessions.append(run_session(3, 3, [], full_game_script(200, (0.8, 0.25)), 0, seed=7))
    sessions.append(run_session(5, 5, [], [("search", 300, (0.8, 0.25)), ("search", 200, (0.8, 0.25)),
                                           ("reroot", "argmax", True), ("search", 300, (0.3, 0.4)),
                                           ("search", 100, (0.0, 0.0)), ("reroot", "argmax", False),
                                           ("search", 100, (0.0, 0.3)), ("search", 100, (0.8, 0.25))], 0, seed=11))
    sessions.append(run_session(3, 3, [9], [("search", 100, (0.0, 0.0)), ("search", 100, (0.0, 0.0)),
                                            ("search", 50, (0.0, 0.25)), ("reroot", 2, True),
                                            ("search", 64, (0.0, 0.0))], 1))
    # re-root onto a never-visited child, then search (unexpanded reused root)
    sessions.append(run_session(3, 3, [], [("search", 5, (0.0, 0.0)), ("reroot", 27, True),
                                           ("search", 40, (0.0, 0.0))], 0))
    # terminal-heavy endgames: every position of the reference's test_boards.csv
    for i, moves, _nxt, _z in load_csv_positions():
        s = run_session(3, 3, moves, [("search", 800 if i % 3 == 0 else 200, (0.0, 0.0))], i % 2)
        s["csv_id"] = i
        sessions.append(s)
    return sessions


# --------------------------------------------------------------- self-play
def gen_selfplay():
    import self_play
    from utils.utils import DotDict
    out = []
    for (L, C), n_read, seeds, kind in (((3, 3), 100, (0, 1, 2), 0), ((3, 3), 800, (5,), 1), ((2, 2), 60, (3, 4), 0),
                                        ((5, 5), 60, (9,), 0)):
        set_board(L, C)
        params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": (0.8, 0.25),
                                        "mcts": {"mcts_num_read": n_read, "mcts_cpuct": CPUCT,
                                                 "temperature": {0: 1.0, 12: 0.02}, "max_async_searches": 1}}})
        for seed in seeds:
            sp = self_play.SelfPlay(make_nn(kind), params)
            np.random.seed(seed)
            asyncio.run(sp.play_game(BoxesState(), seed))
            idx, seq, z = sp.played_games[0]
            df = sp.get_datasets(3, with_features=True).reset_index()
            rec = {"L": L, "C": C, "num_read": n_read, "seed": seed, "kind": kind, "noise": [0.8, 0.25],
                   "temperature": {"0": 1.0, "12": 0.02}, "z": int(z),
                   "moves": [int(n.move) for n in seq[1:]],
                   "visits": [[int(x) for x in n.child_number_visits] for n in seq[:-1]],
                   "columns": list(df.columns),
                   "rows": [[f(x) for x in row] for row in df.to_numpy(dtype=np.float64)]}
            out.append(rec)
    return out


# -------------------------------------------------------------- symmetries
def gen_symmetries():
    """Outputs of the reference's SymmetriesGenerator (dots_boxes_nn.py:11-58) for each of its 8 choices."""
    import importlib
    import torch
    ref_nn = importlib.import_module("dots_boxes.dots_boxes_nn")
    gen = ref_nn.SymmetriesGenerator()
    out = []
    for L in (3, 5):
        r = L + 1
        torch.manual_seed(L)
        b = torch.randint(0, 2, (5, 3, r, r)).float()
        b[:, 0, :, -1] = 0; b[:, 1, -1, :] = 0
        b[:, 2] = torch.randint(1, 9, (5, 1, 1)).float()
        p = torch.rand(5, 2, r, r)
        p[:, 0, :, -1] = 0; p[:, 1, -1, :] = 0   # policies carry no mass on padding slots
        p = p.reshape(5, -1)
        rec = {"L": L, "boards": hx(b.numpy(), np.float32), "policies": hx(p.numpy(), np.float32), "out": []}
        for i in range(8):
            orig = random.randint
            random.randint = lambda a, b_, i=i: i
            try:
                rb, rp = gen(b.clone(), p.clone())
            finally:
                random.randint = orig
            rec["out"].append({"boards": hx(rb.numpy(), np.float32), "policies": hx(rp.reshape(5, -1).numpy(), np.float32)})
        out.append(rec)
    return out


def gen_mcts_pending():
    """max_pending_evals > 1 (the reference ships 64, configuration.py:35) with a net that yields once per call."""
    out = []
    out.append(run_session(3, 3, [], full_game_script(800), 0, max_pending=8))
    out.append(run_session(3, 3, [6, 20], full_game_script(800), 0, max_pending=64))
    out.append(run_session(3, 3, [], full_game_script(203, (0.8, 0.25)), 1, seed=3, max_pending=5))
    out.append(run_session(5, 5, [], full_game_script(400, max_moves=12), 0, max_pending=64))
    out.append(run_session(5, 5, [3], full_game_script(150, (0.8, 0.25)), 0, seed=4, max_pending=16))
    out.append(run_session(2, 2, [], full_game_script(100), 1, max_pending=64))
    out.append(run_session(3, 3, [], [("search", 5, (0.0, 0.0)), ("reroot", 27, True), ("search", 40, (0.0, 0.0)),
                                      ("search", 7, (0.0, 0.0)), ("reroot", "argmax", False), ("search", 33, (0.0, 0.0))], 0,
                           max_pending=4))
    for i, moves, _nxt, _z in load_csv_positions()[::4]:
        s = run_session(3, 3, moves, [("search", 300, (0.0, 0.0))], i % 2, max_pending=8 if i % 3 else 64)
        s["csv_id"] = i
        out.append(s)
    return out


def node_record(node, depth):
    """A node of the reference's tree with its children (UCTNode.children, mcts.py:50-60), `depth` levels down."""
    rec = {"move": None if node.move is None else int(node.move),
           "visits": [int(x) for x in node.child_number_visits], "W": hx(node.child_total_value, np.float32),
           "priors": hx(np.asarray(node.child_priors, dtype=np.float64), np.float64),
           "sign": [int(x) for x in node.child_player_changed],
           "N": int(np.asarray(node.number_visits).ravel()[0]), "own_W": float(np.asarray(node.total_value).ravel()[0]),
           "ucb": hx(np.asarray(node.children_ucb_score(), dtype=np.float64), np.float64),
           "is_terminal": bool(node.is_terminal), "is_expanded": bool(node.is_expanded), "state": state_record(node.game_state)}
    if depth > 0:
        rec["children"] = {str(a): node_record(c, depth - 1) for a, c in sorted(node.children.items())}
    return rec


def gen_treewalk():
    """Trees of the reference walked from Python: root, children and grandchildren after a search (and after a re-root
    with reuse), for UCTNode.children / print_mcts_tree of the drop-in."""
    out = []
    for (L, C, pre, n, kind) in ((3, 3, [0, 4, 16], 300, 0), (3, 3, [], 150, 1), (2, 2, [0, 1], 120, 0), (5, 5, [3, 40], 200, 0)):
        set_board(L, C)
        s = BoxesState()
        for m in pre:
            s.play_(int(m))
        root = mcts.create_root_uct_node(s)
        nn = make_nn(kind)
        asyncio.run(mcts.UCT_search(root, n, nn, cpuct=CPUCT, max_pending_evals=1, dirichlet=(0.0, 0.0)))
        first = node_record(root, 2)
        mv = int(np.argmax(root.child_number_visits))
        root = mcts.init_mcts_tree(root, mv, reuse_tree=True)
        asyncio.run(mcts.UCT_search(root, n, nn, cpuct=CPUCT, max_pending_evals=1, dirichlet=(0.0, 0.0)))
        out.append({"L": L, "C": C, "pre_moves": pre, "num_reads": n, "kind": kind, "first": first, "reroot_move": mv,
                    "second": node_record(root, 2)})
    return out


def dump(name, obj):
    import gzip
    with gzip.GzipFile(os.path.join(OUT, name + ".json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(obj, separators=(",", ":")).encode())


def main():
    which = sys.argv[1:] or ["games", "mcts", "selfplay", "symmetries", "mcts_pending", "treewalk"]
    if "games" in which:
        dump("games", gen_games())
    if "mcts" in which:
        dump("mcts", gen_mcts())
    if "selfplay" in which:
        dump("selfplay", gen_selfplay())
    if "symmetries" in which:
        dump("symmetries", gen_symmetries())
    if "mcts_pending" in which:
        dump("mcts_pending", gen_mcts_pending())
    if "treewalk" in which:
        dump("treewalk", gen_treewalk())


if __name__ == "__main__":
    main()
