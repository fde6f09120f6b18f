"""ctypes binding of oracle/dbaz_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see the header of dbaz_oracle.c).  The product
package dotsboxesaz_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "dbaz_oracle.c")
_LIB = os.path.join(_HERE, "liborc.so")
NONE = 2  # orc_result() code for "get_result() is None"
MAX_A = 128


def build(force=False):
    """gcc -O2, no FMA contraction (the reference's float64 UCB has none)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-o", _LIB, _SRC, "-lm"])
    return _LIB


class State(C.Structure):
    _fields_ = [("board", C.c_uint8 * MAX_A), ("to_play", C.c_int32), ("just_played", C.c_int32),
                ("btc2", C.c_int32 * 2), ("hash_lo", C.c_uint64), ("hash_hi", C.c_uint64),
                ("hash_btc2", C.c_int32), ("pad_", C.c_int32)]


class Game(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("L", "C", "rows", "cols", "plane", "A", "nboxes")]


NN_FN = C.CFUNCTYPE(None, C.POINTER(Game), C.POINTER(State), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        assert L.orc_sizeof_state() == C.sizeof(State)
        L.orc_play.restype = C.c_int
        L.orc_result.restype = C.c_int
        L.orc_philox_u32.restype = C.c_uint32
        L.orc_philox_u32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_random_rollout.argtypes = [C.POINTER(Game), C.POINTER(State), C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_tree_new.restype = C.c_void_p
        L.orc_tree_new.argtypes = [C.c_int, C.c_int, C.POINTER(State)]
        L.orc_tree_free.argtypes = [C.c_void_p]
        L.orc_uct_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                     C.c_void_p, C.c_double]
        L.orc_uct_search_k.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                       C.c_void_p, C.c_double, C.c_int]
        L.orc_reroot.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_root_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_root_state.argtypes = [C.c_void_p, C.POINTER(State)]
        L.orc_root_flags.argtypes = [C.c_void_p]
        L.orc_root_N.argtypes = [C.c_void_p]
        L.orc_root_N.restype = C.c_int64
        L.orc_root_W.argtypes = [C.c_void_p]
        L.orc_root_W.restype = C.c_float
        L.orc_tree_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_tree_stats.restype = C.c_float
        L.orc_tree_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_root_ucb.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.orc_selfplay_argmax.argtypes = [C.c_int, C.c_int, C.POINTER(State), C.c_int, C.c_int, C.c_double,
                                          C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_fake_nn.argtypes = [C.POINTER(Game), C.POINTER(State), C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleGame:
    """BoxesState restated (dots_boxes_game.py:10-118) on the C oracle."""

    def __init__(self, L, Cc, state=None):
        self.g = Game()
        lib().orc_game_init(C.byref(self.g), L, Cc)
        self.L, self.C, self.A = L, Cc, self.g.A
        self.s = State()
        if state is None:
            lib().orc_state_init(C.byref(self.g), C.byref(self.s))
        else:
            C.memmove(C.byref(self.s), C.byref(state), C.sizeof(State))

    def copy(self):
        return OracleGame(self.L, self.C, self.s)

    def valid_moves(self):
        out = np.zeros(self.A, dtype=np.uint8)
        lib().orc_valid_moves(C.byref(self.g), C.byref(self.s), out.ctypes.data_as(C.c_void_p))
        return out.astype(bool)

    def play_(self, move):
        closed = (C.c_int32 * 4)()
        n = lib().orc_play(C.byref(self.g), C.byref(self.s), int(move), closed)
        if n < 0:
            raise ValueError("Illegal move: %d" % move)
        return [(closed[2 * i], closed[2 * i + 1]) for i in range(n)]

    def result(self):
        r = lib().orc_result(C.byref(self.g), C.byref(self.s))
        return None if r == NONE else r

    def features(self):
        out = np.zeros(3 * self.g.plane, dtype=np.int16)
        lib().orc_features(C.byref(self.g), C.byref(self.s), out.ctypes.data_as(C.c_void_p))
        return out.reshape(3, self.g.rows, self.g.cols)

    def board(self):
        return np.frombuffer(bytes(self.s.board)[:self.A], dtype=np.uint8).reshape(2, self.g.rows, self.g.cols)

    def hash0(self):
        return int(self.s.hash_lo) | (int(self.s.hash_hi) << 64)

    def record(self):
        return {"board": bytes(self.s.board)[:self.A].hex(), "to_play": self.s.to_play,
                "just_played": self.s.just_played, "btc2": [self.s.btc2[0], self.s.btc2[1]],
                "hash0": str(self.hash0()), "hash1_x2": self.s.hash_btc2,
                "result": lib().orc_result(C.byref(self.g), C.byref(self.s))}

    def random_rollout(self, seed, game):
        moves = np.zeros(self.A, dtype=np.int32)
        n = lib().orc_random_rollout(C.byref(self.g), C.byref(self.s), seed, game, moves.ctypes.data_as(C.c_void_p))
        return moves[:n].tolist()

    def fake_nn(self, kind=0):
        p = np.zeros(self.A, dtype=np.float32)
        v = np.zeros(1, dtype=np.float32)
        k = C.c_int(kind)
        lib().orc_fake_nn(C.byref(self.g), C.byref(self.s), p.ctypes.data_as(C.c_void_p),
                          v.ctypes.data_as(C.c_void_p), C.cast(C.byref(k), C.c_void_p))
        return p, v


class OracleTree:
    """mcts.py (UCTNode/TreeRoot/UCT_search/init_mcts_tree) restated, sequential sims."""

    def __init__(self, L, Cc, root_state=None, nn=None, kind=0):
        self.L, self.C = L, Cc
        g = OracleGame(L, Cc, root_state)
        self.A = g.A
        self.t = lib().orc_tree_new(L, Cc, C.byref(g.s))
        self._kind = C.c_int(kind)
        self._nn_py = nn
        self._cb = None
        if nn is not None:
            A = self.A

            def tramp(gp, sp, pp, vp, _u):
                p, v = nn(OracleGame(L, Cc, sp.contents))
                for i in range(A):
                    pp[i] = float(p[i])
                vp[0] = float(np.asarray(v).ravel()[0])
            self._cb = NN_FN(tramp)

    def __del__(self):
        if getattr(self, "t", None):
            lib().orc_tree_free(self.t)
            self.t = None

    def search(self, num_reads, cpuct=(1.25, 19652), noise=None, coeff=0.0, max_pending=1):
        fn = C.cast(self._cb, C.c_void_p) if self._cb is not None else C.cast(lib().orc_fake_nn, C.c_void_p)
        user = None if self._cb is not None else C.cast(C.byref(self._kind), C.c_void_p)
        nz = None
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float64)
            assert nz.shape == (self.A,)
        lib().orc_uct_search_k(self.t, int(num_reads), fn, user, float(cpuct[0]), float(cpuct[1]),
                               None if nz is None else nz.ctypes.data_as(C.c_void_p), float(coeff), int(max_pending))
        return self.root()["visits"]

    def reroot(self, move, reuse=True):
        if lib().orc_reroot(self.t, int(move), int(bool(reuse))) != 0:
            raise ValueError("illegal re-root move %d" % move)

    def root(self, cpuct=(1.25, 19652)):
        A = self.A
        N = np.zeros(A, np.int32); W = np.zeros(A, np.float32); P = np.zeros(A, np.float64); S = np.zeros(A, np.int32)
        lib().orc_root_arrays(self.t, *(x.ctypes.data_as(C.c_void_p) for x in (N, W, P, S)))
        st = np.zeros(3, np.int64)
        q = lib().orc_tree_stats(self.t, st.ctypes.data_as(C.c_void_p))
        ucb = np.zeros(A, np.float64)
        lib().orc_root_ucb(self.t, float(cpuct[0]), float(cpuct[1]), ucb.ctypes.data_as(C.c_void_p))
        s = State()
        lib().orc_root_state(self.t, C.byref(s))
        fl = lib().orc_root_flags(self.t)
        return {"visits": N, "W": W, "priors": P, "sign": S, "root_N": int(lib().orc_root_N(self.t)),
                "root_W": float(lib().orc_root_W(self.t)), "stats": [int(st[0]), int(st[1]), int(st[2]), float(q)],
                "ucb": ucb, "state": OracleGame(self.L, self.C, s), "is_expanded": bool(fl & 1),
                "is_terminal": bool(fl & 2), "priors_f64": bool(fl & 4)}

    def counters(self):
        c = np.zeros(2, np.int64)
        lib().orc_tree_counters(self.t, c.ctypes.data_as(C.c_void_p))
        return int(c[0]), int(c[1])


def selfplay_argmax(L, Cc, num_reads, max_moves=128, start=None, cpuct=(1.25, 19652), kind=0):
    """Bulk fake-NN argmax self-play; returns (visits[m, A], moves[m], sims, path_nodes)."""
    g = OracleGame(L, Cc, start)
    vis = np.zeros((max_moves, g.A), np.int32)
    mv = np.zeros(max_moves, np.int32)
    cnt = np.zeros(2, np.int64)
    m = lib().orc_selfplay_argmax(L, Cc, C.byref(g.s), num_reads, max_moves, float(cpuct[0]), float(cpuct[1]),
                                  vis.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p),
                                  cnt.ctypes.data_as(C.c_void_p), int(kind))
    return vis[:m], mv[:m], int(cnt[0]), int(cnt[1])
