"""GPU, BASELINE.json sizes: properties that do not need the oracle to finish 3 M simulations --
conservation of visits, determinism (eager == CUDA-graph == repeated), a sampled subset of trees checked
exactly against the oracle, and rule invariants of a million random playouts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dotsboxesaz_b200 import engine
    from oracle import oracle
    return engine, oracle


def _roots(eng, n, seed):
    g = torch.Generator(device=eng.device)
    g.manual_seed(seed)
    st = eng.new_states(n)
    depth = torch.randint(0, 13, (n,), generator=g, device=eng.device)
    for ply in range(12):
        legal = eng.valid_moves(st).float()
        mv = torch.multinomial(legal + 1e-9, 1, generator=g).reshape(-1).int()
        eng.play(st, torch.where(depth > ply, mv, torch.full_like(mv, -1)))
    return st


def test_config1_4096_games_800_sims(mods):
    """BASELINE configs[1]: 3x3, 4096 concurrent games, 800 sims/move (the bit-exact visit-count parity run)."""
    engine, oracle = mods
    n, sims = 4096, 800
    eng = engine.Engine((3, 3), n_games=n, max_nodes=2048)
    roots = _roots(eng, n, 7)
    ev = engine.FakeNetEvaluator(0)
    runs = []
    for graph_waves in (0, 16, 16):
        eng.reset_roots(roots)
        eng.run_search(sims, ev, graph_waves=graph_waves)
        vis = eng.root_visits()
        W, P, _, _ = eng.root_children()
        st, _, _ = eng.tree_stats()
        runs.append((vis.clone(), W.clone(), st.clone()))
    for r in runs[1:]:  # eager == graph == repeated, bit for bit
        assert torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1]) and torch.equal(r[2], runs[0][2])
    vis, W, st = runs[0]
    # the production schedule -- eval cache, in-kernel chains, compact rows, adaptive batch ladder -- on a second engine
    # (twice: cold table, then a table that already holds every evaluation): same counts, W and statistics, bit for bit
    eng2 = engine.Engine((3, 3), n_games=n, max_nodes=2048, eval_cache=22)
    for rnd in range(2):
        eng2.reset_roots(roots)
        eng2.run_search(sims, ev, graph_waves=8, adaptive=True)
        W2, _, _, _ = eng2.root_children()
        st2, _, _ = eng2.tree_stats()
        assert torch.equal(eng2.root_visits(), vis) and torch.equal(W2, W) and torch.equal(st2, st), rnd
        info2 = eng2.status()
        assert info2["errors"] == 0 and info2["sims"] == n * (sims + 1)
        assert info2["cache_hits"] > (0.3 if rnd == 0 else 0.7) * info2["sims"]  # warm: all but terminal leaves and evicted entries
    eng2.close()
    assert (vis.sum(1) == sims).all()                      # every simulation after the root expansion visits one child
    assert (st[:, 0] == sims + 1).all()                    # root N counts the expansion sim too
    legal = eng.valid_moves(roots)
    assert (vis[~legal] == 0).all()                        # never an illegal action
    assert (st[:, 6] <= sims + 1).all() and (st[:, 6] >= 2).all() and (st[:, 7] == 0).all()
    assert eng.status()["errors"] == 0
    # a spread-out subset, exactly, against the oracle
    vis_np, W_np = vis.cpu().numpy(), W.cpu().numpy()
    roots_np = eng.states_to_numpy(roots)
    for g in range(0, n, 173):
        og = oracle.OracleGame(3, 3)
        e = int(roots_np["edges"][g][0])
        # rebuild the root by replaying its edges in any order is NOT valid (turn order matters): take the state as is
        og.s.to_play = int(roots_np["to_play"][g]); og.s.just_played = int(roots_np["just_played"][g])
        og.s.btc2[0] = int(roots_np["btc2"][g][0]); og.s.btc2[1] = int(roots_np["btc2"][g][1])
        for a in range(32):
            if (e >> a) & 1:
                og.s.board[a] = 255
        og.s.hash_lo = e
        og.s.hash_btc2 = og.s.btc2[og.s.to_play]
        t = oracle.OracleTree(3, 3, og.s)
        ov = t.search(sims)
        assert np.array_equal(vis_np[g], ov), g
        assert np.array_equal(W_np[g], t.root()["W"]), g
    # advancing every root on its most visited move keeps exactly that child's visits - 1 ... + its subtree
    best = vis.argmax(1).int()
    kept = vis.gather(1, best.long().unsqueeze(1)).reshape(-1)
    eng.advance_roots(best, reuse=True)
    st2, _, _ = eng.tree_stats()
    assert torch.equal(st2[:, 2], kept) and (st2[:, 0] == 0).all()           # tree_size = visits of the chosen child; own N restarts
    vis2 = eng.root_visits()
    exp_children = torch.where(st2[:, 5] == 1, torch.zeros_like(kept), (kept - 1).clamp_min(0))
    assert torch.equal(vis2.sum(1), exp_children)                            # the child's own expansion visit has no grandchild
    eng.close()


def test_config4_5x5_16384_games_production_schedule_equals_plain(mods):
    """BASELINE configs[3] sizes (5x5 boxes, 16384 concurrent games; 72 actions = 3 per lane, two mask words): the
    production schedule (eval cache, in-kernel chains, compact rows, adaptive batches) against the plain wave loop, bit for
    bit, on 200 simulations per tree, plus conservation of visits; a spread of trees exactly against the oracle."""
    engine, oracle = mods
    n, sims = 16384, 200
    ev = engine.FakeNetEvaluator(0)
    plain = engine.Engine((5, 5), n_games=n, max_nodes=sims + 8)
    g = torch.Generator(device=plain.device)
    g.manual_seed(11)
    st = plain.new_states(n)
    depth = torch.randint(0, 25, (n,), generator=g, device=plain.device)
    played = []
    for ply in range(24):
        legal = plain.valid_moves(st).float()
        mv = torch.multinomial(legal + 1e-9, 1, generator=g).reshape(-1).int()
        mv = torch.where(depth > ply, mv, torch.full_like(mv, -1))
        played.append(mv.cpu().numpy())
        plain.play(st, mv)
    plain.reset_roots(st)
    plain.run_search(sims, ev, graph_waves=8)
    vis, (W, P, _, U), (stats, rW, q) = plain.root_visits(), plain.root_children(), plain.tree_stats()
    plain.close()
    eng = engine.Engine((5, 5), n_games=n, max_nodes=sims + 8, eval_cache=20)
    eng.reset_roots(st)
    eng.run_search(sims, ev, graph_waves=8, adaptive=True)
    W2, P2, _, U2 = eng.root_children()
    stats2, rW2, q2 = eng.tree_stats()
    assert torch.equal(eng.root_visits(), vis) and torch.equal(W2, W) and torch.equal(P2, P) and torch.equal(U2, U)
    assert torch.equal(stats2, stats) and torch.equal(rW2, rW) and torch.equal(q2, q)
    info = eng.status()
    assert info["errors"] == 0 and info["cache_hits"] > 0
    terminal_root = stats[:, 5] == 1
    assert (vis.sum(1)[~terminal_root] == sims).all() and (vis.sum(1)[terminal_root] == 0).all()
    eng.close()
    vis_np = vis.cpu().numpy()
    for t in range(0, n, 1489):
        og = oracle.OracleGame(5, 5)
        for mv in played:
            if mv[t] >= 0:
                og.play_(int(mv[t]))
        if og.result() is not None:
            continue
        assert np.array_equal(oracle.OracleTree(5, 5, og.s).search(sims), vis_np[t]), t


def test_config3_million_random_playouts(mods):
    """BASELINE configs[2]: 5x5, 2^20 concurrent games, random legal moves to terminal."""
    engine, _ = mods
    n = 1 << 20
    eng = engine.Engine((5, 5), n_games=1, max_nodes=4)
    st = eng.new_states(n)
    plies = eng.random_rollout(st, seed=0)
    h = eng.states_to_numpy(st)
    res = eng.result(st).cpu().numpy()
    plies = plies.cpu().numpy()
    e0, e1 = h["edges"][:, 0], h["edges"][:, 1]
    pop = np.array([bin(int(x)).count("1") for x in e0[:4096]]) + np.array([bin(int(x)).count("1") for x in e1[:4096]])
    assert np.array_equal(pop, plies[:4096])                               # one edge per ply
    assert plies.min() >= 25 and plies.max() <= 60                          # >half of 25 boxes needs >= 25 edges... at most all 60
    assert ((res == 1) | (res == -1)).all()                                 # 25 boxes: no draw, never None at the end
    b = h["btc2"].astype(np.int64)
    assert (b.min(1) < 0).all() and ((b[:, 0] < 0) ^ (b[:, 1] < 0)).all()   # exactly one player passed the majority
    closed = 50 - b.sum(1)                                                  # 2 * boxes closed so far
    assert (closed % 2 == 0).all() and (closed // 2 >= 13).all() and (closed // 2 <= 25).all()
    winner_moved_last = h["just_played"] == h["to_play"]
    assert winner_moved_last.all()                                          # the deciding move closes a box: mover keeps the turn
    assert (res == 1).all()                                                 # ... so the state is a win for the player to move
    # idempotence: a second rollout of finished games changes nothing
    before = st.clone()
    p2 = eng.random_rollout(st, seed=1)
    assert torch.equal(st, before) and int(p2.sum()) == 0
    eng.close()
