"""GPU: one full coach iteration (self-play on the engine -> replay store -> train -> checkpoint -> next generation
plays with the new weights), single process; the 2-GPU variant is tools/coach_bench.py under torchrun."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_two_generations(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dotsboxesaz_b200 import coach, configuration
    from dotsboxesaz_b200.dots_boxes.dots_boxes_game import BoxesState
    from dotsboxesaz_b200.utils.utils import ReplayStore
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
    timings = coach.learn_to_play(params, 0, 2)
    assert [t["generation"] for t in timings] == [0, 1]
    for g in (0, 1):
        ck = torch.load(os.path.join(str(tmp_path), "model_gen%d.pt" % g), map_location="cpu")
        assert set(ck) == {"last_batch_idx", "model_dict", "optimizer_dict"}
    assert ck["last_batch_idx"] > 0  # generation 1 trains min(2*1, nb_epochs) = 1 epoch
    data = ReplayStore(params.hdf_file).read("data")
    gens = sorted(set(data.index.get_level_values("generation")))
    assert gens == [0, 1]
    assert set(np.unique(data["training"])) <= {-1, 1}
    per_gen = data.groupby(level="generation").size()
    assert (per_gen >= 64 * 12).all()  # a 3x3 game has at least 13 searched roots ... allow margin
    pi = data[[c for c in data.columns if c.startswith("pi_")]].to_numpy()
    assert np.allclose(pi.sum(1), 1.0)
    assert set(np.unique(data["z"])) <= {-1.0, 0.0, 1.0}
