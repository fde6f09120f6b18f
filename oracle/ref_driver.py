"""Driver of the REAL reference (oracle/_ref, staged by oracle/make_ref.sh) for bench.py's CPU legs.

TEST INFRASTRUCTURE: only bench.py (`--impl reference`, `cpu_baseline`) and tests/ may import this; the product never
does.  Every function below calls the reference's own, unmodified code -- `mcts.UCT_search`, `BoxesState`,
`utils.proxies.AsyncBatchedProxy`, `nn.NeuralNetWrapper`, `dots_boxes_nn.SimpleNN` -- exactly as its self-play worker
wires them (self_play.py:166-234, configuration.py:13-40: batch 48, 50 ms age trigger, 400 k LRU,
max_async_searches 64, Dirichlet (0.8, 0.25), cpuct (1.25, 19652)).
"""
import asyncio
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "mcts.py"))


def _import_ref():
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `sh oracle/make_ref.sh` where /root/reference exists")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import mcts                                           # noqa: E402  (the reference's)
    from dots_boxes.dots_boxes_game import BoxesState, nn_batch_builder
    from utils.proxies import AsyncBatchedProxy
    from utils.utils import DotDict
    return mcts, BoxesState, nn_batch_builder, AsyncBatchedProxy, DotDict


def worker_search(args):
    """Bounded sample of bench.py's workload on the reference: `n_pos` synthetic roots (0..plies random plies), one
    UCT_search(num_read) each.  Returns (simulations run = UCTNode.backup calls, seconds).  Same job tuple as
    oracle.py_port.worker_search."""
    board_dims, num_read, n_pos, seed, net, plies = args
    import numpy as np
    import torch
    torch.set_num_threads(1)                              # the reference's process layout: one worker per core
    mcts, BoxesState, nn_batch_builder, AsyncBatchedProxy, DotDict = _import_ref()
    BoxesState.init_static_fields((tuple(board_dims),))
    np.random.seed(seed)
    rng = np.random.RandomState(seed)
    counter = [0]
    orig_backup = mcts.UCTNode.backup

    def counting_backup(self, *a, **k):                   # "count sims by wrapping UCTNode.backup"
        counter[0] += 1
        return orig_backup(self, *a, **k)
    mcts.UCTNode.backup = counting_backup

    async def main():
        task = None
        if net == "simple":
            from nn import NeuralNetWrapper
            from dots_boxes.dots_boxes_nn import SimpleNN
            torch.manual_seed(0)
            params = DotDict({"nn": {"pytorch_device": "cpu"}})
            wrapper = NeuralNetWrapper(SimpleNN(), params)
            nnet = AsyncBatchedProxy(wrapper, batch_size=48, timeout=0.05, batch_builder=nn_batch_builder, cache_size=400000)
            task = asyncio.ensure_future(nnet.run())
            pending = 64
        else:
            def _fake(state):
                h = state.get_hash()[0] & 0xffffffff
                i = np.arange(BoxesState.NB_ACTIONS, dtype=np.uint64)
                raw = ((np.uint64(h) * np.uint64(2654435761) + i * np.uint64(40503)) % np.uint64(1024)).astype(np.float32) + np.float32(1)
                return raw / raw.sum(), np.array([((h % 2001) - 1000) / 1000.0], dtype=np.float32)

            async def nnet(state):
                return _fake(state)
            pending = 1
        t0 = time.time()
        for _ in range(n_pos):
            state = BoxesState()
            for _ in range(int(rng.randint(0, plies + 1))):
                state.play_(int(rng.choice(state.get_valid_moves(as_indices=True))))
            root = mcts.create_root_uct_node(state)
            await mcts.UCT_search(root, num_read, nnet, cpuct=(1.25, 19652), max_pending_evals=pending, dirichlet=(0.8, 0.25))
        dt = time.time() - t0
        if task is not None:
            task.cancel()
        return counter[0], dt
    try:
        return asyncio.run(main())
    finally:
        mcts.UCTNode.backup = orig_backup


def worker_game(args):
    """BASELINE configs[0], the reference's own CPU-runnable case: whole 3x3 self-play games at `num_read` sims/move
    through the reference's SelfPlay.play_game (self_play.py:51-74) with its dots_boxes_nn (SimpleNN) on the CPU behind
    AsyncBatchedProxy, tree reuse, Dirichlet (0.8, 0.25), temperature {0: 1.0, 12: 0.02}, max_async_searches 64.
    Returns (simulations, seconds, games, moves)."""
    board_dims, num_read, n_games, seed, net = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    mcts, BoxesState, nn_batch_builder, AsyncBatchedProxy, DotDict = _import_ref()
    import self_play as ref_self_play                     # the reference's
    BoxesState.init_static_fields((tuple(board_dims),))
    np.random.seed(seed)
    counter = [0]
    orig_backup = mcts.UCTNode.backup

    def counting_backup(self, *a, **k):
        counter[0] += 1
        return orig_backup(self, *a, **k)
    mcts.UCTNode.backup = counting_backup
    params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": (0.8, 0.25),
                                    "mcts": {"mcts_num_read": int(num_read), "mcts_cpuct": (1.25, 19652),
                                             "temperature": {0: 1.0, 12: 0.02}, "max_async_searches": 64 if net == "simple" else 1}},
                      "nn": {"pytorch_device": "cpu"}})

    async def main():
        task = None
        if net == "simple":
            from nn import NeuralNetWrapper
            from dots_boxes.dots_boxes_nn import SimpleNN
            torch.manual_seed(0)
            nnet = AsyncBatchedProxy(NeuralNetWrapper(SimpleNN(), params), batch_size=48, timeout=0.05, batch_builder=nn_batch_builder,
                                     cache_size=400000)
            task = asyncio.ensure_future(nnet.run())
        else:
            async def nnet(state):
                h = state.get_hash()[0] & 0xffffffff
                i = np.arange(BoxesState.NB_ACTIONS, dtype=np.uint64)
                raw = ((np.uint64(h) * np.uint64(2654435761) + i * np.uint64(40503)) % np.uint64(1024)).astype(np.float32) + np.float32(1)
                return raw / raw.sum(), np.array([((h % 2001) - 1000) / 1000.0], dtype=np.float32)
        sp = ref_self_play.SelfPlay(nnet, params)
        t0 = time.time()
        for g in range(n_games):
            await sp.play_game(BoxesState(), g)
        dt = time.time() - t0
        if task is not None:
            task.cancel()
        moves = sum(len(seq) - 1 for _, seq, _ in sp.played_games)
        return counter[0], dt, n_games, moves
    try:
        return asyncio.run(main())
    finally:
        mcts.UCTNode.backup = orig_backup


def worker_rollouts(args):
    """BASELINE configs[2] on the reference: uniformly random legal playouts to the end with BoxesState.play_.
    Returns (plies, seconds)."""
    board_dims, n_games, seed = args
    import numpy as np
    _, BoxesState, _, _, _ = _import_ref()
    BoxesState.init_static_fields((tuple(board_dims),))
    rng = np.random.RandomState(seed)
    plies = 0
    t0 = time.time()
    for _ in range(n_games):
        s = BoxesState()
        while s.get_result() is None:
            mv = s.get_valid_moves(as_indices=True)
            s.play_(int(mv[rng.randint(len(mv))]))
            plies += 1
    return plies, time.time() - t0


if __name__ == "__main__":
    print(worker_search(((3, 3), 100, 2, 0, "simple", 12)))
    print(worker_search(((3, 3), 800, 1, 0, "fake", 12)))
    print(worker_rollouts(((5, 5), 20, 0)))
    print(worker_game(((3, 3), 100, 1, 0, "simple")))
