"""Batching front-end for a Python-side net (reference: utils/proxies.py:18-75).

Same surface -- `AsyncBatchedProxy(func, batch_size, timeout, batch_builder, max_queue_size, cache_size, cache_hash)`,
`await proxy(game_state) -> (p, v)`, `proxy.run()` as a background task -- so it can stand between the drop-in
`UCT_search(max_pending_evals=K)` (which awaits K leaves concurrently) and `NeuralNetWrapper`.  One deliberate
difference: the reference dispatches a batch only when its OLDEST request is older than `timeout` (there is no
"batch full" trigger), which costs >= `timeout` of wall clock per batch; here a batch is also dispatched as
soon as `batch_size` requests are waiting.  Results are cached in an LRU keyed by `get_hash()` as in the reference.
"""
import asyncio
import collections
import time


def default_batch_builder(*states_batch):
    import numpy as np
    return np.stack([gs[0].get_features() for gs in states_batch], axis=0)


class AsyncBatchedProxy:
    def __init__(self, func, batch_size, timeout=None, batch_builder=None, max_queue_size=None, cache_size=400000,
                 cache_hash=lambda args: args[0].get_hash()):
        self.func = func
        self.batch_size = batch_size
        self.timeout = timeout if timeout is not None else 0.05
        self.batch_builder = batch_builder or default_batch_builder
        self.with_cache = cache_size > 0
        self.cache_size = cache_size
        self.cache = collections.OrderedDict()
        self.cache_hash = cache_hash
        self.max_queue_size = max_queue_size if max_queue_size else 2 * batch_size
        self.queue = None
        self.n_batches = 0
        self.n_evals = 0

    def _q(self):
        if self.queue is None:  # created lazily inside the running loop
            self.queue = asyncio.Queue(maxsize=self.max_queue_size)
        return self.queue

    async def __call__(self, *args):
        if self.with_cache:
            key = self.cache_hash(args)
            hit = self.cache.get(key)
            if hit is not None:
                self.cache.move_to_end(key)
                return hit
        fut = asyncio.get_running_loop().create_future()
        await self._q().put((time.time(), args, fut))
        res = await fut
        if self.with_cache:
            self.cache[key] = res
            if len(self.cache) > self.cache_size:
                self.cache.popitem(last=False)
        return res

    async def run(self):
        q = self._q()
        pending = []
        try:
            while True:
                wait = self.timeout
                if pending:
                    wait = max(0.0, self.timeout - (time.time() - pending[0][0]))
                try:
                    item = await asyncio.wait_for(q.get(), timeout=wait if pending else None)
                    pending.append(item)
                    while len(pending) < self.batch_size and not q.empty():
                        pending.append(q.get_nowait())
                except asyncio.TimeoutError:
                    pass
                if pending and (len(pending) >= self.batch_size or time.time() - pending[0][0] >= self.timeout):
                    batch, pending = pending[:self.batch_size], pending[self.batch_size:]
                    ps, vs = await self.func(self.batch_builder(*[b[1] for b in batch]))
                    self.n_batches += 1
                    self.n_evals += len(batch)
                    for i, (_, _, fut) in enumerate(batch):
                        if not fut.done():
                            fut.set_result((ps[i], vs[i]))
        except asyncio.CancelledError:
            return
