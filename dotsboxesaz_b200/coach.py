"""Generation loop: self-play -> train -> (broadcast) -> next generation (reference: coach.py:11-162).

One process per GPU (torchrun); every rank plays its shard of the generation's games on its own engine, the sample
rows are gathered to rank 0, rank 0 trains exactly as the reference does (single device), saves model_gen{g}.pt in
the reference's checkpoint format and broadcasts the new weights with one NCCL collective -- the reference's workers
re-read the checkpoint from disk instead (self_play.py:188-190).  The Elo arena (coach.py:98-121) is not part of the
self-play hot path and is not provided.
"""
import argparse
import os
import time

import numpy as np
import torch

from . import self_play
from .utils.utils import ReplayDataset, ReplayStore


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def selfplay(params, generation, model=None, engine=None, evaluator=None):
    """coach.py:11-30.  Returns (DataFrame on rank 0 / None elsewhere, seconds)."""
    tick = time.time()
    rank, _ = _dist()
    store = ReplayStore(params.hdf_file)
    if evaluator is None and model is not None:
        from . import engine as _engine
        from .nn import make_evaluator
        if engine is None:
            n = max(1, min(int(params.self_play.get("concurrent_games", 4096) or 4096), params.self_play.num_games))
            engine = _engine.Engine(tuple(params.game.clazz.BOARD_DIM), n_games=n,
                                    max_nodes=int(params.self_play.get("max_nodes_per_tree", 8192) or 8192),
                                    eval_cache=self_play.default_eval_cache(params))
        evaluator = make_evaluator(model, engine)
    df = self_play.generate_games(params.hdf_file, generation, params.nn.model_class, params.self_play.num_games, params,
                                  engine=engine, evaluator=evaluator, writer=lambda f, k, d: store.append(k, d))
    if rank == 0:
        print(f"Selfplay finished. Generation of {params.self_play.num_games} games took {time.time() - tick:.0f} sec.", flush=True)
    return df, time.time() - tick


def train_nn(params, generation, min_generation, writer=None, model=None):
    """coach.py:33-95 on rank 0: move `fresh` rows into `data` with a train/validation split, build the datasets of
    the training window, run NeuralNetWrapper.train, write model_gen{generation}.pt."""
    from .nn import NeuralNetWrapper
    tick = time.time()
    tp = params.nn.train_params
    store = ReplayStore(params.hdf_file)
    if store.has("fresh"):
        fresh = store.read("fresh")
        train = fresh.sample(frac=tp.train_split)
        fresh = fresh.assign(training=np.where(fresh.index.isin(train.index), 1, -1).astype(np.int8))
        store.append("data", fresh)
        store.delete("fresh")
    else:
        print("No new training data! Is it normal?", flush=True)
    df = store.read("data", min_generation)
    shape = tuple(params.game.clazz.FEATURES_SHAPE)
    n = tp.max_samples_per_gen
    train_ds = ReplayDataset(df, True, shape, int(n * tp.train_split), tp.pos_average)
    val_ds = ReplayDataset(df, False, shape, int(n * (1 - tp.train_split)), tp.pos_average)
    if model is None:
        model = params.nn.model_class(params)
    wrapper = NeuralNetWrapper(model, params)
    if params.nn.lr_scheduler is not None:
        params.nn.train_params.lr = params.nn.lr_scheduler(generation)
    if len(train_ds) < tp.train_batch_size:  # drop_last would leave nothing to train on
        tp.train_batch_size = max(1, len(train_ds))
    if len(val_ds) < tp.val_batch_size:
        tp.val_batch_size = max(1, len(val_ds))
    last = wrapper.train(train_ds, val_ds if len(val_ds) else None, writer, generation)
    print(f"Training finished in {time.time() - tick:.0f} sec. (batch_idx={last})", flush=True)
    return last, wrapper.model


def learn_to_play(params, from_generation, to_generation, writer=None, start_train=False):
    """coach.py:123-162 without the Elo step.  Returns per-generation timings."""
    rank, world = _dist()
    dev = torch.device("cuda", torch.cuda.current_device())
    model = params.nn.model_class(params).to(dev)
    if from_generation > 0:
        model.load_parameters(from_generation - 1, to_device=dev)
    self_play.broadcast_model(model)  # generation 0: every rank plays with rank 0's initial weights
    timings = []
    for generation in range(from_generation, to_generation):
        t = {"generation": generation}
        if not (start_train and generation == from_generation):
            _, t["selfplay_s"] = selfplay(params, generation, model=model)
        ws = min(max(4 + (generation - 4) // 2, 4), 20)  # training window, coach.py:148-149
        tick = time.time()
        if rank == 0:
            _, trained = train_nn(params, generation, generation - ws, writer, model=model)
            model = trained
        t["train_s"] = time.time() - tick
        tick = time.time()
        self_play.broadcast_model(model)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t["broadcast_s"] = time.time() - tick
        timings.append(t)
        if rank == 0:
            print(t, flush=True)
    return timings


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("from_gen", type=int)
    ap.add_argument("to_gen", type=int)
    ap.add_argument("exp")
    ap.add_argument("-t", "--start_train", action="store_true")
    ap.add_argument("-p", "--params", default="simple")
    args = ap.parse_args()
    from . import configuration
    params = getattr(configuration, args.params)
    params.rewrite_str("_exp_", args.exp)
    os.makedirs(params.data_root, exist_ok=True)
    if "RANK" in os.environ:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    learn_to_play(params, args.from_gen, args.to_gen, start_train=args.start_train)


if __name__ == "__main__":
    main()
