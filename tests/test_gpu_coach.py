"""GPU: full coach iterations (self-play on the engine -> device-resident replay window -> train -> checkpoint -> the
next generation plays with the new weights through the SAME evaluator plan and CUDA graphs), single process; the
multi-GPU variant is tools/coach_bench.py under torchrun."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _small_params(tmp_path):
    from dotsboxesaz_b200 import configuration
    from dotsboxesaz_b200.dots_boxes.dots_boxes_game import BoxesState
    BoxesState.init_static_fields(((3, 3),))
    params = configuration.simple
    params.rewrite_str("data/_exp_", str(tmp_path))
    params.self_play.num_games = 64
    params.self_play.concurrent_games = 64
    params.self_play.max_nodes_per_tree = 1024
    params.self_play.mcts.mcts_num_read = 50
    params.nn.pytorch_device = "cuda:0"
    params.nn.train_params.nb_epochs = 1
    params.nn.train_params.train_batch_size = 256
    params.nn.train_params.val_batch_size = 64
    return params


def test_two_generations(tmp_path):
    """learn_to_play(0, 1) runs generations 0 AND 1 (to_generation is inclusive, coach.py:143), writes both checkpoints in
    the reference's format and exports every rank's rows to the replay store."""
    from dotsboxesaz_b200 import coach
    from dotsboxesaz_b200.utils.utils import ReplayStore
    params = _small_params(tmp_path)
    params.self_play.export_frames = True
    timings = coach.learn_to_play(params, 0, 1)
    assert [t["generation"] for t in timings] == [0, 1]
    assert all(t["rows"] >= 64 * 12 and t["gather_s"] >= 0 for t in timings)
    for g in (0, 1):
        ck = torch.load(os.path.join(str(tmp_path), "model_gen%d.pt" % g), map_location="cpu")
        assert set(ck) == {"last_batch_idx", "model_dict", "optimizer_dict"}
    assert ck["last_batch_idx"] > 0  # generation 1 trains min(2*1, nb_epochs) = 1 epoch
    data = ReplayStore(params.hdf_file).read("fresh")
    gens = sorted(set(data.index.get_level_values("generation")))
    assert gens == [0, 1]
    per_gen = data.groupby(level="generation").size()
    assert (per_gen >= 64 * 12).all()  # a 3x3 game has at least 13 searched roots ... allow margin
    assert [int(v) for v in per_gen] == [t["rows"] for t in timings]   # the exported rows are the gathered rows
    pi = data[[c for c in data.columns if c.startswith("pi_")]].to_numpy()
    assert np.allclose(pi.sum(1), 1.0)
    assert set(np.unique(data["z"])) <= {-1.0, 0.0, 1.0}
    # the reference's store-based training step still works on what was exported
    last, _ = coach.train_nn(params, 1, 0)
    assert last > 0


def test_device_samples_equal_the_dataframe_and_pos_average_matches_pandas(tmp_path):
    """samples.batch_from_selfplay (device tensors) row for row against get_datasets() (self_play.py:95-156), a short
    last chunk padded with throw-away games, and the device pos_average against `df.groupby(features).mean()`
    (utils/utils.py:72-73)."""
    from dotsboxesaz_b200 import engine, samples, self_play
    params = _small_params(tmp_path)
    params.self_play.export_frames = True
    eng = engine.Engine((3, 3), n_games=48, max_nodes=1024, eval_cache=14)
    try:
        ev = engine.FakeNetEvaluator(0)
        indices = list(range(100, 100 + 70))      # 48 + 22: the second chunk is padded with 26 throw-away games
        batch, frames, info = self_play.play_shard(params, 3, eng, ev, indices, want_frames=True)
        import pandas as pd
        df = pd.concat(frames)
        assert info["chunks"] == 2 and sorted(set(df.index.get_level_values("game_idx"))) == indices
        assert len(batch) == len(df)
        ref = samples.batch_from_frame(df, eng.device)
        order = lambda b: np.lexsort((b.meta[:, 2].cpu().numpy(), b.meta[:, 1].cpu().numpy()))
        oa, ob = order(batch), order(ref)
        for f in samples.SampleBatch.FIELDS:
            a, b = getattr(batch, f).cpu().numpy()[oa], getattr(ref, f).cpu().numpy()[ob]
            assert np.array_equal(a, b), f
        # pos_average: same groups, same means as pandas (float64 mean, float32 result)
        avg = samples.pos_average(batch)
        flat = df.reset_index()
        fcols = [c for c in flat.columns if c.startswith("x_")]
        pcols = [c for c in flat.columns if c.startswith("pi_")]
        g = flat.groupby(fcols)[pcols + ["z"]].mean().reset_index()
        assert len(avg) == len(g) < len(batch)
        key = lambda x: [tuple(r) for r in x]
        got = dict(zip(key(avg.planes.cpu().numpy().astype(np.int64)), zip(avg.pi.cpu().numpy(), avg.z.cpu().numpy())))
        for xr, pr, zr in zip(key(g[fcols].to_numpy(dtype=np.int64)), g[pcols].to_numpy(dtype=np.float32), g["z"].to_numpy(dtype=np.float32)):
            p, z = got[xr]
            assert np.allclose(p, pr, rtol=0, atol=1e-7) and abs(float(z) - float(zr)) <= 1e-7
        # the training dataset built from it
        ds = samples.DeviceDataset(batch, (3, 4, 4), n_samples=10 ** 9, pos_avg=True)
        assert len(ds) == len(g) and ds.features.shape[1:] == (3, 4, 4) and ds.features.dtype == torch.float32
        nb = sum(1 for _ in ds.batches(64))
        assert nb == len(ds) // 64
    finally:
        eng.close()


def test_evaluator_reload_keeps_graphs_and_matches_a_fresh_plan():
    """FusedSimpleNN.load / FusedResNetZero.load: new weights into the same plan (same tensor addresses, so the engine's
    captured wave-loop graphs stay valid) give exactly what a freshly built plan gives."""
    from dotsboxesaz_b200 import engine
    from dotsboxesaz_b200.nn import FusedResNetZero, FusedSimpleNN, ResNetZero, resnet_zero_parameters
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    from dotsboxesaz_b200.utils.utils import DotDict
    for net in ("simple", "resnet"):
        eng = engine.Engine((3, 3), n_games=256, max_nodes=256, eval_cache=12)
        try:
            mk = (lambda: SimpleNN(board=(3, 3))) if net == "simple" else (
                lambda: ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters((3, 3), nb_blocks=3)}})))
            torch.manual_seed(1)
            m1 = mk()
            torch.manual_seed(2)
            m2 = mk()
            cls = FusedSimpleNN if net == "simple" else FusedResNetZero
            plan = cls(m1, eng)
            assert next(m1.parameters()).device.type == "cpu" and next(m1.parameters()).dtype == torch.float32  # caller's module untouched
            eng.reset_roots()
            eng.run_search(64, plan, graph_waves=4, adaptive=True)
            v1 = eng.root_visits().clone()
            n_graph_sets = len(eng._graphs)
            plan.load(m2)
            eng.clear_eval_cache()
            eng.reset_roots()
            eng.run_search(64, plan, graph_waves=4, adaptive=True)
            v2 = eng.root_visits().clone()
            assert len(eng._graphs) == n_graph_sets            # nothing was re-captured
            fresh = cls(m2, eng)
            eng.clear_eval_cache()
            eng.reset_roots()
            eng.run_search(64, fresh, graph_waves=4, adaptive=True)
            v3 = eng.root_visits()
            assert torch.equal(v2, v3) and not torch.equal(v1, v2)
        finally:
            eng.close()
