"""CPU: the host-side logic of the device sample pipeline (dotsboxesaz_b200/samples.py) -- fixed-width records for the
collective, pos_average against pandas, the replay window's split -- and the world-2 gloo gather of variable row counts
(the NCCL path of a coach iteration, self_play.py:264-265)."""
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.multiprocessing as mp


def _random_batch(R, F=48, A=32, seed=0, n_positions=None):
    from dotsboxesaz_b200 import samples
    g = torch.Generator().manual_seed(seed)
    n_positions = n_positions or R
    pos = torch.randint(0, n_positions, (R,), generator=g)
    planes_all = torch.randint(0, 2, (n_positions, F), generator=g).to(torch.uint8)
    planes_all[:, 2 * F // 3:] = torch.randint(0, 19, (n_positions, 1), generator=g).to(torch.uint8)
    planes = planes_all[pos]
    plane = F // 3
    key = torch.zeros((R, 3), dtype=torch.int64)
    for a in range(2 * plane):
        key[:, a >> 6] |= planes[:, a].long() << (a & 63)
    key[:, 2] = planes[:, 2 * plane].long()
    pi = torch.rand((R, A), generator=g)
    pi = (pi / pi.sum(1, keepdim=True)).float()
    z = torch.randint(-1, 2, (R,), generator=g).float()
    meta = torch.stack([torch.full((R,), 7), torch.arange(R), torch.randint(0, 24, (R,), generator=g)], 1).to(torch.int32)
    return samples.SampleBatch(planes, pi, z, key, meta)


def test_pack_unpack_roundtrip():
    from dotsboxesaz_b200 import samples
    for F, A in ((48, 32), (108, 72), (27, 18)):
        b = _random_batch(100, F, A, seed=F)
        rec = b.pack()
        assert rec.shape == (100, samples.SampleBatch.record_bytes(F, A)) and rec.dtype == torch.uint8
        u = samples.SampleBatch.unpack(rec, F, A)
        for f in samples.SampleBatch.FIELDS:
            assert torch.equal(getattr(b, f), getattr(u, f)), f


def test_pos_average_matches_pandas_groupby():
    from dotsboxesaz_b200 import samples
    b = _random_batch(3000, seed=3, n_positions=400)
    avg = samples.pos_average(b)
    F, A = 48, 32
    df = pd.DataFrame(np.concatenate([b.planes.numpy().astype(np.float64), b.pi.numpy().astype(np.float64), b.z.numpy().astype(np.float64)[:, None]], 1),
                      columns=["x_%d" % i for i in range(F)] + ["pi_%d" % i for i in range(A)] + ["z"])
    g = df.groupby(["x_%d" % i for i in range(F)]).mean().reset_index()
    assert len(avg) == len(g) <= 400
    got = {tuple(r): (p, z) for r, p, z in zip(avg.planes.numpy().astype(np.int64), avg.pi.numpy(), avg.z.numpy())}
    for r, p, z in zip(g[["x_%d" % i for i in range(F)]].to_numpy(dtype=np.int64), g[["pi_%d" % i for i in range(A)]].to_numpy(dtype=np.float32),
                       g["z"].to_numpy(dtype=np.float32)):
        gp, gz = got[tuple(r)]
        assert np.allclose(gp, p, rtol=0, atol=1e-7) and abs(float(gz) - float(z)) <= 1e-7
    # batch_from_frame rebuilds the same keys from the feature columns
    df2 = df.assign(generation=np.int16(1), game_idx=np.arange(len(df), dtype=np.int32), move_idx=np.int16(0)).set_index(["generation", "game_idx", "move_idx"])
    b2 = samples.batch_from_frame(df2, torch.device("cpu"))
    assert torch.equal(b2.key, b.key) and torch.equal(b2.planes, b.planes)


def test_replay_window_split_and_dataset():
    from dotsboxesaz_b200 import samples
    w = samples.ReplayWindow(train_split=0.9, seed=1)
    for gen in range(6):
        w.add(gen, _random_batch(1000, seed=gen))
    assert w.rows() == 6000
    tr, va = w.select(True, 2), w.select(False, 2)
    assert len(tr) == 4 * 900 and len(va) == 4 * 100
    assert set(tr.meta[:, 0].tolist()) == {7}  # meta carries what the producer wrote
    w.drop_before(4)
    assert sorted(w.gens) == [4, 5]
    ds = samples.DeviceDataset(w.select(True), (3, 4, 4), n_samples=500)
    assert len(ds) == 500 and ds.value.shape == (500, 1)
    f, p, z = ds[3]
    assert f.shape == (3, 4, 4) and p.shape == (32,) and z.shape == (1,)
    seen = sum(x.shape[0] for x, _, _ in ds.batches(128, drop_last=False))
    assert seen == 500


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
    from dotsboxesaz_b200 import samples
    ok = True
    for rows in ((37, 91), (0, 12), (5, 0)):           # unequal, and an empty shard on either side
        mine = _random_batch(rows[rank], seed=10 + rank) if rows[rank] else None
        full = samples.gather_batches(mine, 48, 32, dst=0, device=torch.device("cpu"))
        if rank == 0:
            want = samples.SampleBatch.cat([_random_batch(rows[r], seed=10 + r) for r in range(world) if rows[r]])
            ok = ok and full is not None and len(full) == sum(rows)
            ok = ok and all(torch.equal(getattr(full, f), getattr(want, f)) for f in samples.SampleBatch.FIELDS)
        else:
            ok = ok and full is None
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_batches_world2_gloo():
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
    assert all(ok for _, ok in res)
