"""BoxesState: drop-in for the reference's dots_boxes/dots_boxes_game.py:10-155.

A BoxesState is a host-side view of ONE packed 32-byte state (include/dbaz_b200.h: dbaz_state).
Every rule -- legal mask, move application with box closure and the extra turn, result,
feature planes -- is computed by the CUDA kernels through the C ABI as a batch of one; nothing
is re-implemented on the host and there is no CPU fallback.  Bulk work should use
dotsboxesaz_b200.engine.Engine directly (thousands of states per call).
"""
import numpy as np
import torch

from .. import engine as _engine
from .._capi import STATE_DTYPE, RESULT_NONE
from ..game import GameState

_RULES = {}


def _rules(dim):
    """The shared batch-of-one rules engine for a board size (lazily created on the current device)."""
    e = _RULES.get(dim)
    if e is None:
        e = _RULES[dim] = _engine.Engine(dim, n_games=1, max_nodes=2)
    return e


class BoxesState(GameState):
    __slots__ = ("_s",)
    BOARD_DIM = (3, 3)
    FEATURES_SHAPE = (3, 4, 4)
    NB_ACTIONS = 32
    NB_BOXES = 9

    @staticmethod
    def init_static_fields(dims):
        """Called as init_static_fields(((L, C),)) -- the reference reads dims[0] (dots_boxes_game.py:21-28)."""
        L, C = dims[0]
        BoxesState.BOARD_DIM = (L, C)
        BoxesState.FEATURES_SHAPE = (3, L + 1, C + 1)
        BoxesState.NB_ACTIONS = 2 * (L + 1) * (C + 1)
        BoxesState.NB_BOXES = L * C

    def __init__(self, _packed=None):
        if _packed is not None:
            self._s = np.array(_packed, dtype=STATE_DTYPE).reshape(1).copy()
            self._s["flags"] = 0; self._s["depth"] = 0; self._s["parent"] = -1; self._s["parent_action"] = -1
            return
        eng = _rules(self.BOARD_DIM)
        self._s = eng.states_to_numpy(eng.new_states(1))

    # ---- plumbing
    @classmethod
    def from_packed(cls, packed):
        return cls(_packed=packed)

    def packed(self):
        return self._s.copy()

    def _dev(self):
        eng = _rules(self.BOARD_DIM)
        return eng, eng.states_from_numpy(self._s)

    def __deepcopy__(self, memo):
        return BoxesState(_packed=self._s)

    def __copy__(self):
        return BoxesState(_packed=self._s)

    # ---- reference attributes
    @property
    def to_play(self):
        return int(self._s["to_play"][0])

    @property
    def just_played(self):
        jp = int(self._s["just_played"][0])
        return None if jp < 0 else jp

    @property
    def boxes_to_close(self):
        return [int(self._s["btc2"][0][0]) / 2, int(self._s["btc2"][0][1]) / 2]

    def _edges(self):
        return int(self._s["edges"][0][0]) | (int(self._s["edges"][0][1]) << 64)

    @property
    def board(self):
        """uint8[2, L+1, C+1]: 0 free, 1 padding, 255 played (dots_boxes_game.py:30-39,67)."""
        L, C = self.BOARD_DIM
        e = self._edges()
        b = np.zeros((2, L + 1, C + 1), dtype=np.uint8)
        b[1, L, :] = 1
        b[0, :, C] = 1
        flat = b.reshape(-1)
        for a in range(flat.size):
            if (e >> a) & 1:
                flat[a] = 255
        return b

    @property
    def hash(self):
        e = self._edges()
        if e == 0:
            return (0, 0)
        return (e, int(self._s["btc2"][0][self.to_play]) / 2)

    # ---- GameState interface, each a batch-of-one call into the CUDA rules kernels
    def get_actions_size(self):
        return self.NB_ACTIONS

    def get_valid_moves(self, as_indices=False):
        eng, st = self._dev()
        m = eng.valid_moves(st)[0].cpu().numpy()
        return np.argwhere(m).ravel().tolist() if as_indices else m

    def get_result(self):
        eng, st = self._dev()
        r = int(eng.result(st)[0])
        return None if r == RESULT_NONE else r

    def play_(self, move):
        eng, st = self._dev()
        ncl, lc = eng.play(st, [int(move)])
        n = int(ncl[0])
        if n < 0:
            L, C = self.BOARD_DIM
            plc = np.unravel_index(int(move), (2, L + 1, C + 1)) if 0 <= int(move) < self.NB_ACTIONS else None
            raise ValueError("Illegal move: " + str(move) + "->" + str(plc) + "\n" + str(self))
        self._s = eng.states_to_numpy(st)
        lc = lc[0].cpu().tolist()
        return [(lc[2 * i], lc[2 * i + 1]) for i in range(n)]

    def play(self, move):
        nxt = BoxesState(_packed=self._s)
        nxt.play_(move)
        return nxt

    def get_features(self):
        eng, st = self._dev()
        return eng.features(st, torch.int16)[0].cpu().numpy()

    def get_hash(self):
        return self.hash

    def __hash__(self):
        return self.hash.__hash__()

    def __eq__(self, other):
        return self.hash == other.hash

    def __repr__(self):
        b = self.board
        _, lines, cols = b.shape
        out = ["-" * 30, "Just played = " + str(self.just_played), "To play = " + str(self.to_play),
               "Boxes to close = " + str(self.boxes_to_close), "Result = " + str(self.get_result())]
        for l in range(lines):
            out.append("+" + "".join("---+" if b[0, l, c] == 255 else "   +" for c in range(cols - 1)))
            out.append("".join("|   " if b[1, l, c] == 255 else "    " for c in range(cols)) if l < lines - 1 else "")
        return "\n".join(out)


def nn_batch_builder(*game_states):
    """(N, 3, L+1, C+1) features of N states, each passed as a 1-tuple as AsyncBatchedProxy does
    (dots_boxes_game.py:148-155, utils/proxies.py:63) -- one batched kernel call."""
    eng = _rules(BoxesState.BOARD_DIM)
    packed = np.concatenate([gs[0]._s for gs in game_states])
    return eng.features(eng.states_from_numpy(packed), torch.int16).cpu().numpy()
