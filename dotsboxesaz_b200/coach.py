"""Generation loop: self-play -> train -> broadcast -> next generation (reference: coach.py:11-162).

One process per GPU (torchrun).  Every rank plays its shard of the generation's games on ONE engine that lives for the
whole run (sized from the shard, closed explicitly), with an evaluator whose folded weights are refreshed IN PLACE after
every training step, so the CUDA graphs of the wave loop are captured once.  The (features, pi, z) rows stay in HBM:
they reach the training rank through one NCCL gather per generation (samples.gather_batches), sit in a device-resident
replay window, are de-duplicated (`pos_average`) by a device sort + segment mean and feed the optimizer as device
batches; rank 0 trains exactly as the reference does (single device, SGD, one random symmetry per batch), saves
model_gen{g}.pt in the reference's checkpoint format and broadcasts the new weights with one NCCL collective -- the
reference's workers re-read the checkpoint from disk instead (self_play.py:188-190).  DataFrames / the replay store on
disk are export only (`params.self_play.export_frames`).  The store-based functions of the reference (`selfplay`,
`train_nn`) are kept for drop-in use.
"""
import argparse
import os
import time

import numpy as np
import torch

from . import samples, self_play
from .utils.utils import ReplayDataset, ReplayStore


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def selfplay(params, generation, model=None, engine=None, evaluator=None, return_batch=False):
    """coach.py:11-30.  Returns (DataFrame of this rank's rows or None, seconds) -- with return_batch
    (df, gathered samples.SampleBatch on rank 0, info dict, seconds)."""
    tick = time.time()
    rank, world = _dist()
    own_engine = False
    if evaluator is None and model is not None:
        from .nn import make_evaluator
        if engine is None:
            engine = self_play.shard_engine(params, len(self_play.shard_game_indices(params.self_play.num_games, rank, world)))
            own_engine = True
        evaluator = make_evaluator(model, engine)
    try:
        df, batch, info = self_play.generate_games(params.hdf_file, generation, params.nn.model_class, params.self_play.num_games,
                                                   params, engine=engine, evaluator=evaluator, return_batch=True)
    finally:
        if own_engine:
            engine.close()
    if rank == 0:
        print(f"Selfplay finished. Generation of {params.self_play.num_games} games took {time.time() - tick:.0f} sec.", flush=True)
    return (df, batch, info, time.time() - tick) if return_batch else (df, time.time() - tick)


def train_nn(params, generation, min_generation, writer=None, model=None):
    """coach.py:33-95 on rank 0, from the replay store on disk: move `fresh` rows into `data` with a train/validation
    split, build the datasets of the training window, run NeuralNetWrapper.train, write model_gen{generation}.pt."""
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
    return _train(params, generation, train_ds, val_ds, writer, model, tick)


def train_nn_device(params, generation, min_generation, window, writer=None, model=None):
    """train_nn from the device-resident replay window (samples.ReplayWindow): the same sampling cap, pos_average and
    split semantics, no DataFrame, no host copy."""
    tick = time.time()
    tp = params.nn.train_params
    window.drop_before(min_generation)
    shape = tuple(params.game.clazz.FEATURES_SHAPE)
    n = tp.max_samples_per_gen
    train_ds = samples.DeviceDataset(window.select(True, min_generation), shape, int(n * tp.train_split), tp.pos_average)
    val_ds = samples.DeviceDataset(window.select(False, min_generation), shape, int(n * (1 - tp.train_split)), tp.pos_average)
    return _train(params, generation, train_ds, val_ds, writer, model, tick)


def _train(params, generation, train_ds, val_ds, writer, model, tick):
    from .nn import NeuralNetWrapper
    tp = params.nn.train_params
    if model is None:
        model = params.nn.model_class(params)
    wrapper = NeuralNetWrapper(model, params)
    if params.nn.lr_scheduler is not None:
        params.nn.train_params.lr = params.nn.lr_scheduler(generation)
    saved = (tp.train_batch_size, tp.val_batch_size)
    if len(train_ds) < tp.train_batch_size:  # drop_last would leave nothing to train on
        tp.train_batch_size = max(1, len(train_ds))
    if len(val_ds) < tp.val_batch_size:
        tp.val_batch_size = max(1, len(val_ds))
    try:
        last = wrapper.train(train_ds, val_ds if len(val_ds) else None, writer, generation)
    finally:
        tp.train_batch_size, tp.val_batch_size = saved
    print(f"Training finished in {time.time() - tick:.0f} sec. (batch_idx={last}, {len(train_ds)} training positions)", flush=True)
    return last, wrapper.model


def compute_elo(params, generations, elos, models=None, engine=None):
    """coach.py:98-121: a match between two generations (self_play.compute_elo plays it on one engine)."""
    return self_play.compute_elo(params.elo, [params, params], generations, elos, models=models, engine=engine)


def learn_to_play(params, from_generation, to_generation, writer=None, start_train=False, with_elo=False):
    """coach.py:123-162: generations from_generation .. to_generation INCLUSIVE (`while from_generation <=
    to_generation`, coach.py:143).  Returns per-generation timings (self-play, gather, train, broadcast seconds)."""
    import copy
    from .nn import make_evaluator
    rank, world = _dist()
    dev = torch.device("cuda", torch.cuda.current_device())
    model = params.nn.model_class(params).to(dev)
    if from_generation > 0:
        model.load_parameters(from_generation - 1, to_device=dev)
    self_play.broadcast_model(model)  # generation 0: every rank plays with rank 0's initial weights
    shard = self_play.shard_game_indices(params.self_play.num_games, rank, world)
    engine = self_play.shard_engine(params, len(shard), device=dev)
    evaluator = make_evaluator(model, engine)
    window = samples.ReplayWindow(params.nn.train_params.train_split, seed=int(params.self_play.get("seed", 0) or 0))
    elo = 1200.0
    prev_model = None
    timings = []
    try:
        for generation in range(from_generation, to_generation + 1):
            t = {"generation": generation}
            if not (start_train and generation == from_generation):
                if generation > from_generation:
                    if hasattr(evaluator, "load"):
                        evaluator.load(model)      # new weights into the same plan: the captured graphs stay valid
                    else:
                        evaluator = make_evaluator(model, engine)
                _, batch, info, t["selfplay_s"] = selfplay(params, generation, engine=engine, evaluator=evaluator, return_batch=True)
                t.update({k: info[k] for k in ("play_s", "gather_s", "export_s", "sims") if k in info})
                if rank == 0:
                    window.add(generation, batch)
                    t["rows"] = 0 if batch is None else len(batch)
            ws = min(max(4 + (generation - 4) // 2, 4), 20)  # training window, coach.py:148-149
            tick = time.time()
            if rank == 0:
                if with_elo:
                    prev_model = copy.deepcopy(model)
                _, model = train_nn_device(params, generation, generation - ws, window, writer, model=model)
            t["train_s"] = time.time() - tick
            tick = time.time()
            self_play.broadcast_model(model)
            torch.cuda.synchronize()
            t["broadcast_s"] = time.time() - tick
            if with_elo and rank == 0 and generation > 0 and prev_model is not None:
                old, elo, _ = self_play.compute_elo(params.elo, [params, params], [generation - 1, generation], (elo, elo),
                                                    models=[prev_model, model])
                t["elo"] = elo
            timings.append(t)
            if rank == 0:
                print(t, flush=True)
    finally:
        engine.close()
    return timings


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("from_gen", type=int)
    ap.add_argument("to_gen", type=int)
    ap.add_argument("exp")
    ap.add_argument("-t", "--start_train", action="store_true")
    ap.add_argument("-e", "--elo", action="store_true", help="play the Elo match against the previous generation after training")
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
    learn_to_play(params, args.from_gen, args.to_gen, start_train=args.start_train, with_elo=args.elo)


if __name__ == "__main__":
    main()
