"""Parameter presets (reference: configuration.py:13-162), same keys and values; the engine-specific additions are
`self_play.concurrent_games` (trees per GPU) and `self_play.max_nodes_per_tree`."""
from functools import partial

from .dots_boxes.dots_boxes_game import BoxesState, nn_batch_builder
from .dots_boxes.dots_boxes_nn import SimpleNN, SymmetriesGenerator
from .nn import GenerationLrScheduler, ResNetZero, resnet_zero_parameters
from .utils.utils import DotDict


def _preset(model_class, model_parameters, lr_schedule, lr, board=(3, 3)):
    return DotDict({
        "data_root": "data/_exp_",
        "hdf_file": "data/_exp_/sp_data",  # a directory of parquet parts; name it *.hdf to use pandas HDFStore
        "tensorboard_log": "data/tboard/_exp_",
        "game": {"clazz": BoxesState, "init": partial(BoxesState.init_static_fields, (board,))},
        "self_play": {
            "num_games": 2000, "n_workers": 20, "games_per_workers": 25, "reuse_mcts_tree": True,
            "noise": (0.8, 0.25), "nn_batch_size": 48, "nn_batch_timeout": 0.05, "nn_batch_builder": nn_batch_builder,
            "pytorch_devices": ["cuda:0"], "concurrent_games": 4096, "max_nodes_per_tree": 8192,
            "mcts": {"mcts_num_read": 800, "mcts_cpuct": (1.25, 19652), "temperature": {0: 1.0, 12: 0.02}, "max_async_searches": 64},
        },
        "elo": {"hdf_file": "data/_exp_/elo_data", "n_games": 20, "n_workers": 10, "games_per_workers": 2,
                "self_play_override": {"reuse_mcts_tree": False, "noise": (0.0, 0.0), "mcts": {"mcts_num_read": 1200}}},
        "nn": {
            "model_class": model_class, "pytorch_device": "cuda:0", "chkpts_filename": "data/_exp_/model_gen{}.pt",
            "train_params": {"pos_average": True, "symmetries": SymmetriesGenerator(), "nb_epochs": 10,
                             "max_samples_per_gen": 100 * 4096, "train_split": 0.9, "train_batch_size": 4096,
                             "val_batch_size": 4096, "lr_scheduler": GenerationLrScheduler(lr_schedule), "lr": lr,
                             "optimizer_params": {"momentum": 0.9, "weight_decay": 1e-4}},
            "lr_scheduler": GenerationLrScheduler(lr_schedule),
            "model_parameters": model_parameters,
        },
    })


simple = _preset(SimpleNN, None, {0: 1e-2, 20: 1e-3, 50: 1e-4}, 1e-2)
resnet = _preset(ResNetZero, resnet_zero_parameters((3, 3)), {0: 0.1, 30: 0.01, 50: 0.001}, 0.1)
params = resnet
