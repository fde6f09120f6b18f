"""Drop-in for the reference's mcts.py search API on top of the CUDA engine.

    root = create_root_uct_node(state)
    visits = await UCT_search(root, num_reads, async_nn, cpuct, max_pending_evals, dirichlet, time_limit)
    root = init_mcts_tree(root, move, reuse_tree=True)

Each root owns one single-tree engine handle; select / lazy child creation / expand / backup /
re-rooting all run in the sm_100a kernels (dbaz_search_step, dbaz_search_advance_roots).  The only
host work per simulation is what the reference's API forces: awaiting the caller's Python
`async_nn(game_state)` and handing its (p, v) back.  With max_pending_evals = 1 the simulations are
strictly sequential; with K the engine runs the waves the reference's event loop runs when the net
suspends each _search() once (K select_leaf()s leaving their virtual loss behind, then the K
expand/backup pairs in order) and awaits the K `async_nn` calls concurrently, exactly what a
batching proxy needs.  Throughput comes from dotsboxesaz_b200.self_play.BatchedSelfPlay, which runs
thousands of such trees in lock-step against a device-resident net.

Node objects are views: the CURRENT root reads live engine state; after init_mcts_tree() the old
root keeps a frozen snapshot of its arrays (what self_play.get_datasets reads).  `node.children`
(mcts.py:50,53-60) is a dict {action: node view} of the children that exist in the engine's node pool,
read through dbaz_search_node; such views are valid until the tree is re-rooted (re-rooting compacts the
pool and renumbers the nodes) and raise afterwards.
"""
import collections
import time
import weakref

import numpy as np
import torch

from . import engine as _engine
from .dots_boxes.dots_boxes_game import BoxesState

TreeStats = collections.namedtuple("TreeStats", ["max_deepness", "tree_size", "terminal_count", "q_value"])
VIRTUAL_LOSS = 1
DEFAULT_MAX_NODES = 32768
DEFAULT_MAX_PENDING = 64  # configuration.py:35 (max_async_searches)

_POOL = collections.defaultdict(list)


def _acquire(dim):
    pool = _POOL[dim]
    return pool.pop() if pool else _engine.Engine(dim, n_games=1, max_nodes=DEFAULT_MAX_NODES, max_pending=DEFAULT_MAX_PENDING)


def _release(dim, eng):
    if eng._h:
        _POOL[dim].append(eng)


class _Tree:
    """One engine handle with n_games == 1; returned to the pool when the last node view dies."""

    def __init__(self, dim):
        self.dim = dim
        self.eng = _acquire(dim)
        self.epoch = 0  # bumped by every re-root: node indices of earlier views are stale then
        weakref.finalize(self, _release, dim, self.eng)


class TreeRoot:
    """View of mcts.py:21-36 for the node that is (or was) the first node of the tree."""

    def __init__(self, node):
        self._node = node

    def get_tree_stats(self):
        s = self._node._read()
        return TreeStats(s["max_deepness"], s["tree_size"], s["terminal_count"], s["q_value"])


class UCTNode:
    CPUCT = 1.25
    CPUCT_BASE = 19652

    def __init__(self, tree, game_state, move):
        self._tree = tree
        self._frozen = None
        self.game_state = game_state
        self.move = move
        self.parent = TreeRoot(self)
        self.is_terminal = game_state.get_result() is not None  # mcts.py:52

    # ---- state access
    def _read(self):
        if self._frozen is not None:
            return self._frozen
        eng = self._tree.eng
        vis = eng.root_visits()
        W, P, S, U = eng.root_children()
        st, rW, q = eng.tree_stats()
        st = st[0].cpu().numpy()
        return {"visits": vis[0].cpu().numpy(), "W": W[0].cpu().numpy(), "priors": P[0].cpu().numpy(),
                "sign": S[0].cpu().numpy(), "ucb": U[0].cpu().numpy(), "root_N": int(st[0]), "max_deepness": int(st[1]),
                "tree_size": int(st[2]), "terminal_count": int(st[3]), "is_expanded": bool(st[4]), "is_terminal": bool(st[5]),
                "root_W": np.float32(rW[0].item()), "q_value": np.float32(q[0].item())}

    def _freeze(self):
        if self._frozen is None:
            self._frozen = self._read()
            self._tree = None

    @property
    def is_expanded(self):
        return self._read()["is_expanded"]

    @property
    def child_number_visits(self):
        return self._read()["visits"]

    @property
    def child_total_value(self):
        return self._read()["W"]

    @property
    def child_priors(self):
        return self._read()["priors"]

    @property
    def child_player_changed(self):
        return self._read()["sign"]

    @property
    def number_visits(self):
        return self._read()["root_N"]

    @property
    def total_value(self):
        return self._read()["root_W"]

    def children_ucb_score(self):
        return self._read()["ucb"]

    @property
    def children(self):
        """{action: child node} of the children created so far (mcts.py:50,53-60).  Only for the live root and views
        reached from it; a root that has been re-rooted away has released its subtree (mcts.py:175 `del children`)."""
        if self._tree is None:
            return {}
        return _children_of(self._tree, 0, self)

    def best_child(self):
        invalid = 1 - self.game_state.get_valid_moves()
        return np.argmax(-1e12 * invalid + self.children_ucb_score())

    def get_tree_stats(self):
        return self.parent.get_tree_stats()

    def __hash__(self):
        return self.game_state.__hash__()

    def __repr__(self):
        return "\n".join(["*" * 15, "Node: " + str(self.game_state.hash), "Move: " + str(self.move),
                          "#visits: " + str(self.number_visits), "Expanded: " + str(self.is_expanded),
                          "Terminal: " + str(self.is_terminal), "Total value: " + str(self.total_value), str(self.game_state)])


class _NodeView:
    """A non-root node of a live tree, read on demand from the engine's node pool (dbaz_search_node).  Same read-only
    attributes as UCTNode; valid until the tree is re-rooted."""

    def __init__(self, tree, index, move, parent):
        self._tree, self._index, self._epoch = tree, int(index), tree.epoch
        self.move = int(move)
        self.parent = parent
        self.game_state = BoxesState.from_packed(self._read()["state"])

    def _read(self):
        if self._tree.epoch != self._epoch:
            raise RuntimeError("this node view belongs to a tree that has been re-rooted since (node indices changed)")
        return self._tree.eng.node_view(0, self._index)

    is_terminal = property(lambda self: self._read()["is_terminal"])
    is_expanded = property(lambda self: self._read()["is_expanded"])
    child_number_visits = property(lambda self: self._read()["visits"])
    child_total_value = property(lambda self: self._read()["W"])
    child_priors = property(lambda self: self._read()["priors"])
    child_player_changed = property(lambda self: self._read()["sign"])
    number_visits = property(lambda self: self._read()["N"])
    total_value = property(lambda self: self._read()["own_W"])

    def children_ucb_score(self):
        return self._read()["ucb"]

    def best_child(self):
        invalid = 1 - self.game_state.get_valid_moves()
        return np.argmax(-1e12 * invalid + self.children_ucb_score())

    @property
    def children(self):
        return _children_of(self._tree, self._index, self)

    def __hash__(self):
        return self.game_state.__hash__()


def _children_of(tree, index, parent):
    idx = tree.eng.node_view(0, index)["child"]
    return {int(a): _NodeView(tree, int(idx[a]), int(a), parent) for a in np.flatnonzero(idx)}


def create_root_uct_node(game_state):
    """mcts.py:156-160"""
    tree = _Tree(tuple(type(game_state).BOARD_DIM))
    eng = tree.eng
    eng.reset_roots(eng.states_from_numpy(game_state.packed()))
    return UCTNode(tree, game_state, None)


def init_mcts_tree(previous_node, move, reuse_tree=True):
    """mcts.py:163-180: re-root on `move`, keeping the subtree (compacted in place on the device) or not."""
    tree = previous_node._tree
    if tree is None:
        raise RuntimeError("init_mcts_tree: this node is no longer the root of a live tree")
    previous_node._freeze()
    eng = tree.eng
    tree.epoch += 1
    eng.advance_roots([int(move)], reuse=bool(reuse_tree))
    try:
        eng.status()
    except _engine.EngineError as exc:
        raise ValueError("Illegal move: %s (%s)" % (move, exc)) from None
    state = BoxesState.from_packed(eng.states_to_numpy(eng.root_states()))
    return UCTNode(tree, state, int(move))


async def UCT_search(root_node, num_reads, async_nn, cpuct=(1.25, 19652), max_pending_evals=64, dirichlet=(0.0, 0.0),
                     time_limit=None):
    """mcts.py:183-244.  Returns root_node.child_number_visits (int32[A])."""
    tree = root_node._tree
    if tree is None:
        raise RuntimeError("UCT_search: this node is no longer the root of a live tree")
    eng = tree.eng
    end_time = time.time() + (time_limit if time_limit else 120)
    UCTNode.CPUCT, UCTNode.CPUCT_BASE = cpuct
    eng.set_cpuct(cpuct)

    dev = eng.device
    K = max(1, min(int(max_pending_evals), eng.max_pending))
    import asyncio

    async def drain(n_sims, pending):
        """Run waves until the budget is spent: every wave hands the caller's net the leaves of up to `pending`
        simulations at once (awaited concurrently) and feeds the answers to the next wave's backups."""
        first = min(pending, eng.A)
        waves = 2 + max(0, -(-(n_sims - first) // pending))  # upper bound, see Engine.run_search
        for _ in range(waves):
            eng.step()
            kinds = eng.leaf_kind.cpu().numpy()  # device -> host sync: the caller's Python net must see the leaves
            rows = np.flatnonzero(kinds == 1)
            if rows.size:
                packed = eng.states_to_numpy(eng.leaf_states)
                outs = await asyncio.gather(*(async_nn(BoxesState.from_packed(packed[r])) for r in rows))
                p = np.stack([np.asarray(o[0], dtype=np.float32).reshape(-1) for o in outs])
                v = np.asarray([np.asarray(o[1], dtype=np.float32).reshape(-1)[0] for o in outs], dtype=np.float32)
                idx = torch.from_numpy(rows).to(dev)
                eng.priors.index_copy_(0, idx, torch.from_numpy(p).to(dev))
                eng.values.index_copy_(0, idx, torch.from_numpy(v).to(dev))
            if time.time() > end_time:
                eng.step_flush()  # mcts.py:232-233: launch no more simulations; back up the pending ones
                return
        eng.step()  # flush the last backups

    if not root_node.is_expanded:
        eng.begin(-2, pending=1)  # mcts.py:207-208, before the noise is drawn (keeps the global RNG order)
        await drain(1, 1)
    alpha, coeff = dirichlet
    noise = None
    if alpha > 0:
        # same draw, from the same global legacy stream, as mcts.py:220-223 (bool mask quirk included)
        valid = root_node.game_state.get_valid_moves()
        conc = valid.copy()
        conc[conc == 0] = 1e-60
        noise = np.random.dirichlet(conc * alpha, 1).ravel() * valid
        noise = torch.from_numpy(noise).reshape(1, -1)
    num_reads = min(int(num_reads), 2_000_000_000)
    if time_limit:
        # a time-limited search (players.py:62-63 asks for 1e12 reads) also ends when the tree's node pool is used up:
        # every simulation creates at most one node
        st, _, _ = eng.tree_stats()
        num_reads = max(0, min(num_reads, eng.max_nodes - int(st[0, 6]) - K - 1))
    eng.begin(num_reads, noise, float(coeff), pending=K)
    await drain(num_reads, K)
    eng.status()
    return root_node.child_number_visits


def print_mcts_tree(node, max_level=10, prefix=" "):
    """mcts.py:247-272: the node, then its children recursively."""
    if max_level < 0:
        return

    def top3(arr):
        asc = max(arr) == 0
        return "; ".join(f"{i}->{arr[i]:.4f}" for i in np.argsort(arr)[::1 if asc else -1][:3])
    gs = node.game_state
    print(f"{prefix[:-1]}{node.move} ({node.number_visits}/{node.total_value}) -> {gs.to_play} {gs.get_result()}")
    print(f"{prefix[:-1]} - child values:{top3(node.child_total_value)}")
    print(f"{prefix[:-1]} - child visits:{top3(node.child_number_visits)}")
    print(f"{prefix[:-1]} - priors:{top3(node.child_priors)}")
    print(f"{prefix[:-1]} - ucb:{top3(node.children_ucb_score())}")
    for n in node.children.values():
        print_mcts_tree(n, max_level - 1, prefix + "   |")
