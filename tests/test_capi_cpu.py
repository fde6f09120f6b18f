"""CPU: the C-ABI library loads and exports every symbol include/dbaz_b200.h declares; host logic that
needs no GPU (configs, Elo, sharding); the engine fails loudly without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dbaz_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dbaz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dotsboxesaz_b200 import _capi
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 25
    assert sorted(_capi.SYMBOLS) == names, set(names) ^ set(_capi.SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.dbaz_abi_version() == 1
    assert lib.dbaz_sizeof_state() == 32 == _capi.STATE_DTYPE.itemsize


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dotsboxesaz_b200 import _capi, engine
    with pytest.raises(RuntimeError):
        engine.Engine((3, 3), n_games=4)
    lib = _capi.load()
    cfg = _capi.Config(1, 0, 3, 3, 4, 16, 0, 0, 1.25, 19652.0)
    h = ctypes.c_void_p()
    assert lib.dbaz_engine_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.dbaz_last_error(None)
    bad = _capi.Config(1, 0, 9, 9, 4, 16, 0, 0, 1.25, 19652.0)
    assert lib.dbaz_engine_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"too large" in lib.dbaz_last_error(None)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dotsboxesaz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), (dirpath, f)


def test_dotdict_and_elo():
    from dotsboxesaz_b200.utils.utils import DictWithDefault, DotDict, elo_rating2
    d = DotDict({"a": {"b": 1, "s": "x/_exp_/y"}, "c": 2})
    assert d.a.b == 1 and d.missing is None
    d.merge({"a": {"b": 3, "z": {"q": 1}}, "e": 5})
    assert d.a.b == 3 and d.a.z.q == 1 and d.e == 5 and d.a.s == "x/_exp_/y"
    d.rewrite_str("_exp_", "run1")
    assert d.a.s == "x/run1/y"
    d.k = {"n": 1}
    assert isinstance(d.k, DotDict) and d.k.n == 1
    calls = []
    dd = DictWithDefault(lambda k: calls.append(k) or k * 2)
    assert dd[3] == 6 and dd[3] == 6 and calls == [3]
    e0, e1 = elo_rating2(1200, 1200, 12, 8, K=30)
    assert abs(e0 - 1260) < 1e-9 and abs(e1 - 1140) < 1e-9


def test_shard_game_indices_partition():
    from dotsboxesaz_b200.self_play import shard_game_indices
    for n, w in ((10, 1), (10, 3), (4096, 8), (5, 8)):
        parts = [shard_game_indices(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_dataset_frame_matches_reference_columns():
    """_dataset builds the DataFrame of self_play.py:95-156 (columns, dtypes, MultiIndex)."""
    from dotsboxesaz_b200.self_play import _dataset
    from golden_io import load
    G = load("selfplay")[0]
    cols = G["columns"]
    rows = []
    ref = np.array(G["rows"], dtype=np.float64)
    ix = {c: i for i, c in enumerate(cols)}
    for r in ref:
        rows.append({"game_idx": int(r[ix["game_idx"]]), "move_idx": int(r[ix["move_idx"]]),
                     "move": None if r[ix["move"]] < 0 else int(r[ix["move"]]), "player": int(r[ix["player"]]),
                     "features": r[[ix[c] for c in cols if c.startswith("x_")]].astype(np.int16),
                     "pi": r[[ix[c] for c in cols if c.startswith("pi_")]], "z": r[ix["z"]],
                     "stats": (int(r[ix["max_deepness"]]), int(r[ix["tree_size"]]), int(r[ix["terminal_count"]]),
                               np.float32(r[ix["q_value"]]))})
    df = _dataset(rows, 3, True)
    assert df.index.names == ["generation", "game_idx", "move_idx"]
    assert str(df["move"].dtype) == "int16" and str(df["player"].dtype) == "int8" and str(df["q_value"].dtype) == "float32"
    flat = df.reset_index()
    assert list(flat.columns) == cols
    assert np.array_equal(flat.to_numpy(dtype=np.float64), ref)


def torch_has_cuda():
    import torch
    return torch.cuda.is_available()


def test_adaptive_ladder_and_default_cache_size():
    """Host logic of the adaptive wave loop that needs no GPU: the batch ladder (Engine._ladder) and the default eval
    cache size (self_play.default_eval_cache)."""
    from dotsboxesaz_b200 import engine, self_play
    from dotsboxesaz_b200.utils.utils import DotDict
    e = object.__new__(engine.Engine)
    for n, steps in ((4096, 16), (4096, 8), (192, 8), (64, 8), (100, 8), (16384, 16), (1, 8)):
        e.n_games, e.LADDER_STEPS = n, steps
        lad = e._ladder()
        assert lad[0] == n and lad == sorted(set(lad), reverse=True)
        assert all(r <= n for r in lad) and (min(lad) <= max(64, n // steps) or n <= 64)
        assert all(r % 8 == 0 or r == n for r in lad)            # 16-byte rows for the library kernels
        for busy in (1, n // 3, n - 1, n):                        # every busy count has a rung that holds it
            assert min(r for r in lad if r >= busy) >= busy

    # an evaluator with a batch quantum (the tower kernel: boards per tile x SMs) gets rungs at its multiples
    class Ev:
        batch_quantum = 2664
    e.n_games, e.LADDER_STEPS = 16384, 16
    lad = e._ladder(Ev())
    assert lad[0] == 16384 and lad[1:7] == [15984, 13320, 10656, 7992, 5328, 2664] and min(lad) == 64
    assert lad == sorted(set(lad), reverse=True) and all(r % 8 == 0 for r in lad)
    e.n_games = 1000                                              # fewer games than one quantum: the plain ladder
    assert e._ladder(Ev()) == e._ladder()

    class G33:
        BOARD_DIM = (3, 3)

    class G77:
        BOARD_DIM = (7, 7)
    p = DotDict({"self_play": {"mcts": {"max_async_searches": 64}}, "game": {"clazz": G33}})
    k = self_play.default_eval_cache(p)
    assert k == 24 and (1 << k) * 16 * 32 <= (8 << 30)            # 2^24 entries of 512 bytes = 8 GiB
    p.self_play.eval_cache_log2 = 0
    assert self_play.default_eval_cache(p) == 0
    p2 = DotDict({"self_play": {"mcts": {}}, "game": {"clazz": G77}})
    k7 = self_play.default_eval_cache(p2)                          # A = 128: 2 KB per entry
    assert k7 == 22 and (1 << k7) * 16 * 128 <= (8 << 30)
    # sized for an engine: ~4096 slots per tree, never more than the memory budget (8 GB when there is no device to ask)
    p3 = DotDict({"self_play": {"mcts": {}}, "game": {"clazz": G33}})
    assert self_play.default_eval_cache(p3, 64, 1024) == 18        # 64 trees: 2^18 entries
    assert self_play.default_eval_cache(p3, 1024, 4096) == 22
    assert self_play.eval_cache_log2_for(32, 1, 64) == 16           # floor
    big = self_play.eval_cache_log2_for(32, 32768, 4096)
    assert 24 <= big <= 27 and (big == 24 or torch_has_cuda())


def test_pick_rows_prefers_cheap_rungs_and_holds_the_rows():
    """Engine._pick_rows (host logic of the adaptive loop): with a measured evaluator-time curve that has tile-wave steps,
    the chosen rung holds the expected rows unless a rung at most 10 % short is at least 5 % cheaper per row served."""
    from dotsboxesaz_b200 import engine
    e = object.__new__(engine.Engine)
    e.n_games, e.LADDER_STEPS = 4096, 16
    lad = e._ladder()
    curve = {4096: 223.0, 3840: 209.2, 3584: 212.8, 3328: 181.1, 3072: 180.6, 2816: 179.3, 2560: 177.2, 2304: 146.4, 2048: 141.2,
             1792: 138.8, 1536: 130.1, 1280: 130.0, 1024: 103.8, 768: 80.6, 512: 70.9, 256: 56.4, 128: 52.5, 64: 51.1}
    e._eval_us = {(1, r): curve[r] for r in lad}
    assert e._pick_rows(lad, 4096, 1) == 4096
    assert e._pick_rows(lad, 2400, 1) == 2304      # 2560 would cost a second wave of tiles for 96 more rows
    assert e._pick_rows(lad, 2000, 1) == 2048
    assert e._pick_rows(lad, 10, 1) == 64
    for want in range(1, 4097, 37):
        r = e._pick_rows(lad, want, 1)
        assert r in lad and r >= engine.Engine.UNDERSIZE * want
    # whatever the curve, the choice serves at least as many rows per microsecond as the smallest rung that holds them all
    for us in (curve, {r: 50.0 + 0.15 * r for r in lad}):
        e._eval_us = {(1, r): us[r] for r in lad}
        score = lambda r, want: min(r, want) / (us[r] + engine.Engine.WAVE_OVERHEAD_US)
        for want in range(1, 4097, 53):
            r = e._pick_rows(lad, want, 1)
            assert score(r, want) >= score(min(x for x in lad if x >= want), want) - 1e-12
            if r < want:
                assert score(r, want) > engine.Engine.UNDERSIZE_GAIN * max(score(x, want) for x in lad if x >= want)
    # without measurements: the smallest rung that holds the rows
    e._eval_us = {}
    assert e._pick_rows(lad, 2400, 1) == 2560


def test_device_loop_rule_equals_the_host_rule():
    """The rung choice of the one-graph-per-search loop (k_loop_decide -> loop_pick, csrc/dbaz_loop.cuh) is the host's
    Engine._pick_rows: same rung for every number of rows wanted, on a measured evaluator curve with tile-wave steps and
    on a smooth one."""
    import ctypes as C
    from dotsboxesaz_b200 import _capi, engine
    lib = _capi.load()
    e = object.__new__(engine.Engine)
    e.n_games, e.LADDER_STEPS = 4096, 16
    lad = e._ladder()
    curve = {4096: 223.0, 3840: 209.2, 3584: 212.8, 3328: 181.1, 3072: 180.6, 2816: 179.3, 2560: 177.2, 2304: 146.4, 2048: 141.2,
             1792: 138.8, 1536: 130.1, 1280: 130.0, 1024: 103.8, 768: 80.6, 512: 70.9, 256: 56.4, 128: 52.5, 64: 51.1}
    for us in (curve, {r: 50.0 + 0.15 * r for r in lad}):
        e._eval_us = {(1, r): us[r] for r in lad}
        rows = (C.c_int32 * len(lad))(*lad)
        t = (C.c_float * len(lad))(*[us[r] for r in lad])
        for want in range(1, 4097, 7):
            got = lib.dbaz_search_loop_pick(rows, t, len(lad), engine.Engine.UNDERSIZE, engine.Engine.UNDERSIZE_GAIN,
                                            engine.Engine.WAVE_OVERHEAD_US, want)
            assert lad[got] == e._pick_rows(lad, want, 1), (want, lad[got], e._pick_rows(lad, want, 1))
