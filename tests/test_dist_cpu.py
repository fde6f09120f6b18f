"""CPU, world_size 2, gloo: the only collectives of the path -- weight broadcast after a coach iteration
and the gather of sample rows (self_play.broadcast_model / gather_samples) -- plus the index sharding."""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    from dotsboxesaz_b200.self_play import broadcast_model, gather_samples, shard_game_indices
    torch.manual_seed(100 + rank)  # every rank starts from different weights (reference gen-0 behaviour)
    model = SimpleNN(board=(3, 3))
    with torch.no_grad():
        model.bn0.running_mean.fill_(float(rank + 1))
    before = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    broadcast_model(model, src=0)
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    sums = torch.stack([after.double().sum(), model.bn0.running_mean.double().sum(),
                        model.bn0.num_batches_tracked.double()])
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    changed = (rank == 0) == torch.equal(before, after)
    mine = shard_game_indices(11, rank, world)
    df = pd.DataFrame({"generation": np.int16(1), "game_idx": np.asarray(mine, dtype=np.int16),
                       "move_idx": np.int16(0), "z": np.asarray(mine, dtype=np.float64) * 0.5})
    df = df.set_index(["generation", "game_idx", "move_idx"])
    full = gather_samples(df, dst=0)
    ok_gather = True
    if rank == 0:
        ok_gather = (full is not None and sorted(full.index.get_level_values("game_idx")) == list(range(11))
                     and np.allclose(full["z"].to_numpy(), np.arange(11) * 0.5))
    else:
        ok_gather = full is None
    out.put((rank, bool(same), bool(changed), bool(ok_gather)))
    dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, changed, ok_gather in res:
        assert same, "weights differ across ranks after broadcast"
        assert changed, "rank 0 must keep its weights, the others must receive them"
        assert ok_gather
