#!/usr/bin/env python
"""BASELINE configs[4]: full coach iterations (self-play + gather + train + weight broadcast), sharded over the
GPUs of one box.  Launch with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/coach_bench.py --board 5x5
Rank 0 prints one JSON line per generation."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--board", default="5x5")
    ap.add_argument("--net", default="resnet", choices=["simple", "resnet"])
    ap.add_argument("--games", type=int, default=1024, help="games per generation (whole job)")
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--generations", type=int, default=2)
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--export", action="store_true", help="also write every rank's rows to the replay store (DataFrame + parquet)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from dotsboxesaz_b200 import coach, configuration
    from dotsboxesaz_b200.dots_boxes.dots_boxes_game import BoxesState
    from dotsboxesaz_b200.nn import resnet_zero_parameters
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    L, C = (int(x) for x in args.board.split("x"))
    BoxesState.init_static_fields(((L, C),))
    params = configuration.simple if args.net == "simple" else configuration.resnet
    if args.net == "resnet":
        params.nn.model_parameters = resnet_zero_parameters((L, C), nb_blocks=args.blocks)
    root = tempfile.mkdtemp(prefix="dbaz_coach_") if rank == 0 else None
    if world > 1:
        box = [root]
        dist.broadcast_object_list(box, src=0)
        root = box[0]
    params.rewrite_str("data/_exp_", root)
    params.self_play.num_games = args.games
    params.self_play.concurrent_games = max(1, args.games // world)
    params.self_play.max_nodes_per_tree = 4096
    params.self_play.mcts.mcts_num_read = args.sims
    params.nn.pytorch_device = "cuda:%d" % torch.cuda.current_device()
    params.nn.train_params.nb_epochs = 2
    params.nn.train_params.train_batch_size = 1024
    params.nn.train_params.val_batch_size = 1024
    t0 = time.time()
    params.self_play.export_frames = bool(args.export)
    timings = coach.learn_to_play(params, 0, args.generations - 1)  # to_generation is inclusive (coach.py:143)
    if rank == 0:
        for t in timings:
            t.update(board=args.board, net=args.net, games=args.games, sims_per_move=args.sims, n_gpus=world)
            print(json.dumps(t), flush=True)
        print(json.dumps({"total_s": time.time() - t0, "n_gpus": world}), flush=True)
        shutil.rmtree(root, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
