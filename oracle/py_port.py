"""Python/NumPy restatement of the reference's self-play path -- TEST/BENCH INFRASTRUCTURE ONLY.

The reference is pure Python, so its "own CPU implementation" cannot be compiled into
oracle/_ref and cannot travel to the GPU box.  This module restates it in the reference's
own language and data types (NumPy uint8 board, per-node NumPy child arrays, asyncio
fan-out of simulations, an age-triggered batching proxy with an LRU in front of a torch
net on CPU) so that bench.py can time "the reference's CPU MCTS" on the box's host cores.
It is pinned against the same fixtures as the C oracle (tests/test_py_port_golden.py).

Only tests/ and bench.py (cpu_baseline / --impl reference) may import it.

Citations are to files under the reference root:
  Board           dots_boxes/dots_boxes_game.py:10-118
  Node / Tree     mcts.py:21-180
  uct_search      mcts.py:183-244
  BatchingProxy   utils/proxies.py:18-75   (dispatch only when the oldest request is older than `timeout`)
  play_game       self_play.py:27-74
"""
import asyncio
import collections
import math
import time

import numpy as np

VIRTUAL_LOSS = 1


class Board:
    DIM = (3, 3)

    @classmethod
    def configure(cls, L, C):
        cls.DIM = (L, C)

    def __init__(self, src=None):
        if src is not None:
            self.cells = src.cells.copy()
            self.to_play, self.just_played = src.to_play, src.just_played
            self.need = list(src.need)
            self.key = src.key
            return
        L, C = Board.DIM
        self.cells = np.zeros((2, L + 1, C + 1), dtype=np.uint8)
        self.cells[1, L, :] = 1
        self.cells[0, :, C] = 1
        self.to_play, self.just_played = 0, None
        self.need = [L * C / 2, L * C / 2]
        self.key = (0, 0)

    @property
    def n_actions(self):
        return self.cells.size

    def legal(self, as_indices=False):
        m = self.cells.ravel() == 0
        return np.flatnonzero(m).tolist() if as_indices else m

    def result(self):
        a, b = self.need[self.to_play], self.need[1 - self.to_play]
        if a == 0 and b == 0:
            return 0
        if a < 0:
            return 1
        if b < 0:
            return -1
        return None

    def _closed(self, l, c):
        q = self.cells
        return int(q[0, l, c]) + int(q[0, l + 1, c]) + int(q[1, l, c]) + int(q[1, l, c + 1]) == 1020

    def apply(self, move):
        p, l, c = np.unravel_index(move, self.cells.shape)
        if self.cells[p, l, c] != 0:
            raise ValueError("Illegal move: %s" % move)
        self.cells[p, l, c] = 255
        rows, cols = self.cells.shape[1:]
        if p == 0:
            cand = ([(l - 1, c)] if l > 0 else []) + ([(l, c)] if l < rows - 1 else [])
        else:
            cand = ([(l, c - 1)] if c > 0 else []) + ([(l, c)] if c < cols - 1 else [])
        done = [(int(a), int(b)) for a, b in cand if self._closed(a, b)]
        self.just_played = self.to_play
        if done:
            self.need[self.to_play] -= len(done)
        else:
            self.to_play = 1 - self.to_play
        self.key = (self.key[0] + (1 << int(move)), self.need[self.to_play])
        return done

    def after(self, move):
        nb = Board(self)
        nb.apply(move)
        return nb

    def planes(self):
        b = self.cells // 255
        k = np.full_like(b[0], self.need[self.to_play] * 2, dtype=np.int8)
        return np.concatenate((b, k[None]), axis=0)


class Tree:
    """TreeRoot: the first node's own statistics and the per-search tree stats."""

    def __init__(self):
        self.first = None
        self.W = collections.defaultdict(float)
        self.N = collections.defaultdict(int)
        self.sign = collections.defaultdict(int)
        self.depth_fix = 0
        self.depth = 0
        self.max_depth = 0
        self.terminals = 0
        self.size = 0

    # the first node addresses these through parent.child_*[move] with move None
    @property
    def child_W(self):
        return self.W

    @property
    def child_N(self):
        return self.N

    @property
    def child_sign(self):
        return self.sign

    def stats(self):
        q = self.first.W_own / (1 + self.first.N_own)
        return (self.max_depth - self.depth_fix, int(self.size), self.terminals, q if isinstance(q, float) else q[0])


class Node:
    __slots__ = ("board", "move", "parent", "expanded", "terminal", "kids", "prior", "child_W", "child_N", "child_sign", "depth")
    CPUCT, CPUCT_BASE = 1.25, 19652

    def __init__(self, board, move, parent):
        A = board.n_actions
        self.board, self.move, self.parent = board, move, parent
        self.expanded = False
        self.terminal = board.result() is not None
        self.kids = {}
        self.prior = np.zeros(A, dtype=np.float32)
        self.child_W = np.zeros(A, dtype=np.float32)
        self.child_N = np.zeros(A, dtype=np.int32)
        self.child_sign = np.ones(A, dtype=np.int32)
        self.depth = parent.depth + 1

    @property
    def N_own(self):
        return self.parent.child_N[self.move]

    @property
    def W_own(self):
        return self.parent.child_W[self.move]

    def kid(self, move):
        k = self.kids.get(move)
        if k is None:
            k = self.kids[move] = Node(self.board.after(move), move, self)
        return k

    def scores(self):
        n_own = self.N_own
        c = math.log((n_own + Node.CPUCT_BASE + 1) / Node.CPUCT_BASE) + Node.CPUCT
        c *= math.sqrt(n_own) / (self.child_N + 1)
        u = c * self.prior
        q = self.child_W / (1 + self.child_N)
        q *= self.child_sign
        return u + q

    def pick(self):
        blocked = 1 - self.board.legal()
        return int(np.argmax(-1e12 * blocked + self.scores()))

    def descend(self):
        cur, path = self, [self]
        while cur.expanded and not cur.terminal:
            cur.parent.child_W[cur.move] -= VIRTUAL_LOSS
            cur = cur.kid(cur.pick())
            path.append(cur)
        return cur, path

    def expand(self, prior):
        self.expanded = True
        self.prior = prior
        self.parent.child_sign[self.move] = 1 if self.board.to_play == self.board.just_played else -1

    def backup(self, path, value):
        me = self.board.to_play
        for n in path:
            v = value * (1 if n.board.to_play == me else -1)
            n.parent.child_W[n.move] += v + VIRTUAL_LOSS
            n.parent.child_N[n.move] += 1
        tree = path[0].parent
        tree.terminals += self.terminal
        tree.max_depth = max(tree.max_depth, self.depth)


def new_root(board):
    t = Tree()
    n = Node(board, None, t)
    t.first = n
    return n


def reroot(prev, move, reuse=True):
    if reuse:
        nxt = prev.kid(move)
        visits = prev.child_N[move]
        t = Tree()
        nxt.parent = t
        t.first = nxt
        t.depth_fix = nxt.depth
        t.size = visits
        prev.kids = None
        return nxt
    nxt = new_root(prev.kid(move).board)
    nxt.move = move
    return nxt


SIM_COUNTER = [0]


async def uct_search(root, num_reads, async_nn, cpuct=(1.25, 19652), max_pending=64, dirichlet=(0.0, 0.0), time_limit=None):
    async def one():
        leaf, path = root.descend()
        if not leaf.terminal:
            p, v = await async_nn(leaf.board)
            p = p * leaf.board.legal()
            s = p.sum()
            if s > 0 and s != 1.0:
                p /= s
        else:
            p, v = np.zeros(leaf.board.n_actions), leaf.board.result()
        leaf.expand(p)
        leaf.backup(path, v)
        SIM_COUNTER[0] += 1

    deadline = time.time() + (time_limit if time_limit else 120)
    Node.CPUCT, Node.CPUCT_BASE = cpuct
    if not root.expanded:
        await one()
    alpha, coeff = dirichlet
    s = root.prior.sum()
    probs = root.prior / root.prior.sum() if s != 0 else np.zeros(len(root.prior))
    if alpha > 0:
        conc = root.board.legal()
        conc[conc == 0] = 1e-60  # bool array: every entry becomes True (reference quirk)
        noise = np.random.dirichlet(conc * alpha, 1).ravel()
        noise *= root.board.legal()
    else:
        noise = 0.0
    root.prior = (1 - coeff) * probs + coeff * noise

    cap = min(max_pending, len(root.board.legal()))
    pending = set()
    for _ in range(num_reads):
        if time.time() > deadline:
            break
        if len(pending) >= cap:
            _, pending = await asyncio.wait(pending, return_when=asyncio.FIRST_COMPLETED)
            cap = max_pending
        pending.add(asyncio.ensure_future(one()))
    if pending:
        await asyncio.wait(pending)
    return root.child_N


class LRU:
    def __init__(self, cap):
        self.cap, self.d = cap, collections.OrderedDict()

    def get(self, k):
        v = self.d.get(k)
        if v is not None:
            self.d.move_to_end(k)
        return v

    def put(self, k, v):
        self.d[k] = v
        self.d.move_to_end(k)
        if len(self.d) > self.cap:
            self.d.popitem(last=False)


class BatchingProxy:
    """utils/proxies.py:18-75: requests queue up; a batch (<= batch_size) is dispatched only once the
    oldest queued request is older than `timeout`; results are cached by position key."""

    def __init__(self, predict_sync, batch_size=48, timeout=0.05, cache_size=400000):
        self.predict_sync, self.batch_size, self.timeout = predict_sync, batch_size, timeout
        self.cache = LRU(cache_size) if cache_size > 0 else None
        self.q = asyncio.Queue(maxsize=2 * batch_size)
        self.batches = 0
        self.evals = 0

    async def __call__(self, board):
        if self.cache is not None:
            hit = self.cache.get(board.key)
            if hit is not None:
                return hit
        fut = asyncio.get_event_loop().create_future()
        await self.q.put((time.time(), board, fut))
        res = await fut
        if self.cache is not None:
            self.cache.put(board.key, res)
        return res

    async def run(self):
        loop = asyncio.get_event_loop()
        ts, bs, fs = [], [], []
        try:
            while True:
                try:
                    async with asyncio.timeout(self.timeout):
                        t, b, f = await self.q.get()
                    ts.append(t); bs.append(b); fs.append(f)
                except asyncio.TimeoutError:
                    pass
                if fs and time.time() - ts[0] > self.timeout:
                    n = min(len(fs), self.batch_size)
                    X = np.stack([b.planes() for b in bs[:n]], axis=0)
                    ps, vs = await loop.run_in_executor(None, self.predict_sync, X)
                    self.batches += 1
                    self.evals += n
                    for i, f in enumerate(fs[:n]):
                        f.set_result((ps[i], vs[i]))
                    ts, bs, fs = ts[n:], bs[n:], fs[n:]
        except asyncio.CancelledError:
            return


async def play_game(async_nn, num_read=800, cpuct=(1.25, 19652), temperature=None, noise=(0.8, 0.25), max_pending=64,
                    reuse=True, start=None):
    """self_play.py:51-74 + 27-49.  Returns (roots..., terminal)."""
    temperature = temperature or {0: 1.0, 12: 0.02}
    root = new_root(start if start is not None else Board())
    seq, i, temp = [], -1, None
    while not root.terminal:
        i += 1
        temp = temperature.get(i, temp)
        k = len(root.board.legal(as_indices=True))
        n = min(4 * math.factorial(k), num_read)
        visits = await uct_search(root, n, async_nn, cpuct, max_pending, noise)
        probs = (visits / visits.max()) ** (1 / temp)
        probs = probs / probs.sum()
        move = np.random.choice(probs.shape[0], 1, p=probs)[0]
        seq.append(root)
        root = reroot(root, move, reuse)
    seq.append(root)
    return seq


# ------------------------------------------------------------ timing harness
def _torch_predict(board_dims, seed=0, threads=1):
    import torch
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    model = SimpleNN(board=board_dims).train(False)

    def predict_sync(X):
        with torch.no_grad():
            p, v = model(torch.tensor(X, dtype=torch.float32))
        return torch.exp(p).numpy(), v.numpy()
    return predict_sync


def fake_nn_eval(board, kind=0):
    A = board.n_actions
    h = int(board.key[0]) & 0xFFFFFFFF
    if kind == 0:
        raw = np.array([float((h * 2654435761 + i * 40503) % 1024) + 1 for i in range(A)], dtype=np.float32)
        return raw / raw.sum(), np.array([((h % 2001) - 1000) / 1000], dtype=np.float32)
    return (np.full(A, np.float32(1.0) / np.float32(A), dtype=np.float32),
            np.array([(((h * 31) % 5) - 2) / 2], dtype=np.float32))


def worker_selfplay(args):
    """One process of the reference's mp.Pool (self_play.py:242-270): plays whole games for `budget_s`
    seconds (finishing the game in progress) and returns (sims, seconds, games, moves)."""
    board_dims, num_read, budget_s, seed, net = args
    Board.configure(*board_dims)
    np.random.seed(seed)

    async def main():
        proxy = None
        if net == "simple":
            proxy = BatchingProxy(_torch_predict(board_dims, seed=0, threads=1))
            task = asyncio.ensure_future(proxy.run())
            nn = proxy
        else:
            async def nn(b):
                return fake_nn_eval(b, 0)
        SIM_COUNTER[0] = 0
        t0 = time.time()
        games = moves = 0
        while time.time() - t0 < budget_s:
            seq = await play_game(nn, num_read=num_read, max_pending=64 if net == "simple" else 1)
            games += 1
            moves += len(seq) - 1
        dt = time.time() - t0
        if proxy is not None:
            task.cancel()
        return SIM_COUNTER[0], dt, games, moves
    return asyncio.run(main())


def worker_search(args):
    """Bounded sample of the bench workload: `n_pos` synthetic roots, one uct_search(num_read) each."""
    board_dims, num_read, n_pos, seed, net, plies = args
    Board.configure(*board_dims)
    np.random.seed(seed)
    rng = np.random.RandomState(seed)

    async def main():
        proxy = None
        if net == "simple":
            proxy = BatchingProxy(_torch_predict(board_dims, seed=0, threads=1))
            task = asyncio.ensure_future(proxy.run())
            nn = proxy
        else:
            async def nn(b):
                return fake_nn_eval(b, 0)
        SIM_COUNTER[0] = 0
        t0 = time.time()
        for _ in range(n_pos):
            b = Board()
            for _ in range(int(rng.randint(0, plies + 1))):
                b.apply(int(rng.choice(b.legal(as_indices=True))))
            root = new_root(b)
            await uct_search(root, num_read, nn, max_pending=64 if net == "simple" else 1, dirichlet=(0.8, 0.25))
        dt = time.time() - t0
        if proxy is not None:
            task.cancel()
        return SIM_COUNTER[0], dt
    return asyncio.run(main())
