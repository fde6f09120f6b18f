"""Device-resident sample pipeline: self-play rows -> NCCL gather -> replay window -> pos_average -> training batches.

Reference path (all host side): every worker appends its get_datasets() DataFrame to the HDF store under a lock
(self_play.py:264-265), coach.train_nn moves `fresh` into `data` with a random train/validation split
(coach.py:57-67), HDFStoreDataset samples the window, merges rows with identical features by averaging pi and z
(`pos_average`, utils/utils.py:72-73) and hands float32 arrays to a DataLoader (nn.py:177-181); the 8-fold board
symmetry is applied per batch on the training device (nn.py:212, dots_boxes_nn.py:11-58).

Here the rows never leave HBM between self-play and the optimizer: `SampleBatch` is a handful of device tensors,
`gather_batches` moves them to the training rank with ONE NCCL collective per generation (all ranks' rows packed into
fixed-width byte records; row counts exchanged first), `ReplayWindow` keeps the last generations on the training GPU,
`pos_average` is a device sort + segment mean over the position key (get_hash(): exactly what get_features() shows,
dots_boxes_game.py:96-112), and `DeviceDataset.batches()` yields shuffled float32 batches without a DataLoader.
DataFrames are only built for export (self_play.BatchedSelfPlay.get_datasets / utils.ReplayStore).
"""
import numpy as np
import torch


class SampleBatch:
    """Rows of (features, pi, z) with their position key and bookkeeping, all on one device.

    planes  uint8  [R, F]      get_features().ravel(): 0/1 edge planes and the constant third plane (2 * boxes_to_close)
    pi      float32 [R, A]     visit distribution of the searched root
    z       float32 [R]        game outcome from the point of view of the player to move
    key     int64  [R, 3]      (low edge word, high edge word, 2 * boxes_to_close[to_play]) = get_hash()
    meta    int32  [R, 3]      (generation, game_idx, move_idx)
    """

    FIELDS = ("planes", "pi", "z", "key", "meta")

    def __init__(self, planes, pi, z, key, meta):
        self.planes, self.pi, self.z, self.key, self.meta = planes, pi, z, key, meta

    def __len__(self):
        return int(self.planes.shape[0])

    @property
    def device(self):
        return self.planes.device

    def to(self, device):
        return SampleBatch(*(getattr(self, f).to(device) for f in self.FIELDS))

    def index(self, idx):
        return SampleBatch(*(getattr(self, f)[idx] for f in self.FIELDS))

    @staticmethod
    def cat(batches):
        batches = [b for b in batches if b is not None and len(b)]
        if not batches:
            return None
        return SampleBatch(*(torch.cat([getattr(b, f) for b in batches]) for f in SampleBatch.FIELDS))

    @staticmethod
    def empty(F, A, device):
        return SampleBatch(torch.zeros((0, F), dtype=torch.uint8, device=device), torch.zeros((0, A), dtype=torch.float32, device=device),
                           torch.zeros((0,), dtype=torch.float32, device=device), torch.zeros((0, 3), dtype=torch.int64, device=device),
                           torch.zeros((0, 3), dtype=torch.int32, device=device))

    # ---- fixed-width byte records for the collective: [planes F | pad to 4 | pi 4A | z 4 | key 24 | meta 12]
    @staticmethod
    def record_bytes(F, A):
        return (F + 3) // 4 * 4 + 4 * A + 4 + 24 + 12

    def pack(self):
        R, F, A = len(self), self.planes.shape[1], self.pi.shape[1]
        Fp = (F + 3) // 4 * 4
        rec = torch.zeros((R, self.record_bytes(F, A)), dtype=torch.uint8, device=self.device)
        rec[:, :F] = self.planes
        o = Fp
        rec[:, o:o + 4 * A] = self.pi.contiguous().view(torch.uint8).reshape(R, 4 * A); o += 4 * A
        rec[:, o:o + 4] = self.z.contiguous().view(torch.uint8).reshape(R, 4); o += 4
        rec[:, o:o + 24] = self.key.contiguous().view(torch.uint8).reshape(R, 24); o += 24
        rec[:, o:o + 12] = self.meta.contiguous().view(torch.uint8).reshape(R, 12)
        return rec

    @staticmethod
    def unpack(rec, F, A):
        R = rec.shape[0]
        Fp = (F + 3) // 4 * 4
        o = Fp
        planes = rec[:, :F].contiguous()
        pi = rec[:, o:o + 4 * A].contiguous().view(torch.float32).reshape(R, A); o += 4 * A
        z = rec[:, o:o + 4].contiguous().view(torch.float32).reshape(R); o += 4
        key = rec[:, o:o + 24].contiguous().view(torch.int64).reshape(R, 3); o += 24
        meta = rec[:, o:o + 12].contiguous().view(torch.int32).reshape(R, 3)
        return SampleBatch(planes, pi, z, key, meta)


def batch_from_selfplay(bsp, generation):
    """SampleBatch of the last play_games_device() / play_games_async() call of a BatchedSelfPlay, straight from its
    device history (self_play.py:95-156 produces the same rows as a DataFrame)."""
    h, eng = bsp._device_hist, bsp.eng
    planes, pi, z, slot, mi = bsp.device_samples()
    R = planes.shape[0]
    n = eng.n_games
    states = torch.stack(h["states"]).reshape(-1, 4)[mi * n + slot]          # the packed root states, row for row
    raw = states.view(torch.uint8).reshape(R, 32)
    btc2 = raw[:, 16:20].contiguous().view(torch.int16).reshape(R, 2).long()
    to_play = raw[:, 20].long()
    key = torch.stack([states[:, 0], states[:, 1], btc2.gather(1, to_play.unsqueeze(1)).squeeze(1)], 1)
    games = torch.as_tensor(np.asarray(h["games_idxs"], dtype=np.int64), device=planes.device)[slot]
    keep = games >= 0                                                          # padding slots of a short last chunk
    meta = torch.stack([torch.full_like(games, int(generation)), games, mi], 1).to(torch.int32)
    out = SampleBatch(planes.reshape(R, -1).to(torch.uint8), pi.float(), z.float(), key, meta)
    return out if bool(keep.all()) else out.index(torch.nonzero(keep).reshape(-1))


def batch_from_frame(df, device, generation=None):
    """SampleBatch from a get_datasets() DataFrame (self_play.py:95-156) -- the host-RNG / drop-in paths and files
    written by the reference.  The position key is rebuilt from the feature columns: bit a of the edge words is
    x_a for the two edge planes, the third component is the constant third plane."""
    flat = df.reset_index()
    fcols = [c for c in flat.columns if c.startswith("x_")]
    pcols = [c for c in flat.columns if c.startswith("pi_")]
    x = flat[fcols].to_numpy(dtype=np.int64)
    R, F = x.shape
    plane = F // 3
    bits = x[:, :2 * plane].astype(np.uint64)
    w = np.zeros((R, 2), dtype=np.uint64)
    for a in range(2 * plane):
        w[:, a >> 6] |= bits[:, a] << np.uint64(a & 63)
    key = np.stack([w[:, 0].view(np.int64), w[:, 1].view(np.int64), x[:, 2 * plane]], 1)
    gen = flat["generation"].to_numpy(dtype=np.int64) if generation is None else np.full(R, int(generation), dtype=np.int64)
    meta = np.stack([gen, flat["game_idx"].to_numpy(dtype=np.int64), flat["move_idx"].to_numpy(dtype=np.int64)], 1).astype(np.int32)
    t = lambda a, dt: torch.as_tensor(np.array(a, copy=True)).to(device=device, dtype=dt)
    return SampleBatch(t(x, torch.uint8), t(flat[pcols].to_numpy(dtype=np.float64), torch.float32),
                       t(flat["z"].to_numpy(dtype=np.float64), torch.float32), t(key, torch.int64), t(meta, torch.int32))


def gather_batches(batch, F, A, dst=0, device=None):
    """Every rank's rows on rank `dst` -- the NCCL form of the reference's locked HDF append (self_play.py:264-265).
    Row counts are exchanged with one all_gather, then ONE gather of fixed-width byte records padded to the largest
    count (NCCL needs equal shapes).  Returns the concatenated SampleBatch on `dst`, None elsewhere.  Works on gloo too
    (CPU tensors) for the host-side tests."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return batch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else (batch.device if batch is not None else torch.device("cpu"))
    if batch is None:
        batch = SampleBatch.empty(F, A, dev)
    counts = [torch.zeros((1,), dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(batch)], dtype=torch.int64, device=dev))
    counts = [int(c) for c in counts]
    rows = max(max(counts), 1)
    rec = torch.zeros((rows, SampleBatch.record_bytes(F, A)), dtype=torch.uint8, device=dev)
    if len(batch):
        rec[:len(batch)] = batch.pack()
    out = [torch.empty_like(rec) for _ in range(world)] if rank == dst else None
    dist.gather(rec, out, dst=dst)
    if rank != dst:
        return None
    return SampleBatch.cat([SampleBatch.unpack(o[:c], F, A) for o, c in zip(out, counts) if c > 0])


def pos_average(batch):
    """`df.groupby(features).mean()` (utils/utils.py:72-73) on the device: rows with identical features merge into one
    with the mean pi and the mean z.  The position key get_hash() = (edge set, boxes_to_close[to_play]) determines the
    features (dots_boxes_game.py:96-112), so the grouping is a lexicographic sort of three int64 columns
    (torch.unique) and two scatter-adds in float64.  Output order: ascending key (pandas: ascending feature columns --
    the same groups, another order; training shuffles)."""
    if batch is None or len(batch) == 0:
        return batch
    uniq, inv, cnt = torch.unique(batch.key, dim=0, return_inverse=True, return_counts=True)
    G = uniq.shape[0]
    c = cnt.double().unsqueeze(1)
    pi = torch.zeros((G, batch.pi.shape[1]), dtype=torch.float64, device=batch.device).index_add_(0, inv, batch.pi.double()) / c
    z = torch.zeros((G,), dtype=torch.float64, device=batch.device).index_add_(0, inv, batch.z.double()) / c.squeeze(1)
    first = torch.full((G,), len(batch), dtype=torch.int64, device=batch.device).scatter_reduce_(
        0, inv, torch.arange(len(batch), device=batch.device), reduce="amin")
    return SampleBatch(batch.planes[first], pi.float(), z.float(), uniq, batch.meta[first])


class DeviceDataset:
    """HDFStoreDataset (utils/utils.py:61-91) on device tensors: float32 features [R, 3, L+1, C+1], pi [R, A], z [R, 1].
    Indexable like the reference's dataset (a DataLoader works) and iterable in shuffled device batches."""

    def __init__(self, batch, features_shape, n_samples=int(1e12), pos_avg=False, generator=None):
        if batch is not None and len(batch) > n_samples:
            perm = torch.randperm(len(batch), device=batch.device, generator=generator)[:n_samples]
            batch = batch.index(perm)
        if pos_avg:
            batch = pos_average(batch)
        self.batch = batch
        n = 0 if batch is None else len(batch)
        self.features = (batch.planes.float().reshape(n, *features_shape) if n else torch.zeros((0,) + tuple(features_shape)))
        self.policy = batch.pi if n else torch.zeros((0, 1))
        self.value = batch.z.reshape(n, 1) if n else torch.zeros((0, 1))

    def __len__(self):
        return int(self.features.shape[0])

    def __getitem__(self, i):
        return self.features[i], self.policy[i], self.value[i]

    def batches(self, batch_size, shuffle=True, drop_last=True, generator=None):
        n = len(self)
        order = torch.randperm(n, device=self.features.device, generator=generator) if shuffle else torch.arange(n, device=self.features.device)
        stop = n - n % batch_size if drop_last else n
        for lo in range(0, stop, batch_size):
            idx = order[lo:lo + batch_size]
            yield self.features[idx], self.policy[idx], self.value[idx]


class ReplayWindow:
    """The replay store on the training GPU: one SampleBatch per generation with its train(+1) / validation(-1) flag
    drawn once when the generation arrives (coach.py:57-63: `fresh.sample(frac=train_split)`), the last `window`
    generations kept (coach.py:148-149 moves the window's lower end)."""

    def __init__(self, train_split=0.9, seed=0):
        self.train_split = float(train_split)
        self.gens = {}          # generation -> (SampleBatch, training flag int8 [R])
        self.seed = int(seed)

    def add(self, generation, batch):
        if batch is None or len(batch) == 0:
            return
        g = torch.Generator(device=batch.device)
        g.manual_seed(self.seed + 7919 * int(generation))
        R = len(batch)
        n_train = int(round(R * self.train_split))
        flag = torch.full((R,), -1, dtype=torch.int8, device=batch.device)
        flag[torch.randperm(R, device=batch.device, generator=g)[:n_train]] = 1
        if generation in self.gens:
            old, oflag = self.gens[generation]
            batch, flag = SampleBatch.cat([old, batch]), torch.cat([oflag, flag])
        self.gens[generation] = (batch, flag)

    def drop_before(self, min_generation):
        for g in [g for g in self.gens if g < min_generation]:
            del self.gens[g]

    def select(self, train, min_generation=None):
        want = 1 if train else -1
        parts = []
        for g in sorted(self.gens):
            if min_generation is not None and g < min_generation:
                continue
            b, flag = self.gens[g]
            parts.append(b.index(torch.nonzero(flag == want).reshape(-1)))
        return SampleBatch.cat(parts)

    def rows(self):
        return sum(len(b) for b, _ in self.gens.values())
