"""The paths that produce BASELINE's games/hour -- BatchedSelfPlay.play_games_device / play_games_async (device RNG,
production schedule: eval cache, in-kernel chains, compact rows, adaptive batch ladder, CUDA graphs) -- and full
config-1 games with host RNG, replayed move by move on the CPU oracle with the same noise and the same sampled moves:
visit counts and tree statistics of EVERY search of EVERY game must be bit-equal.
Reference: self_play.py:27-74 (get_next_move / play_game), mcts.py:183-244 (UCT_search), mcts.py:163-180 (init_mcts_tree).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TEMPERATURE = {0: 1.0, 12: 0.02}
NOISE = (0.8, 0.25)
CPUCT = (1.25, 19652)


def _params(num_read, noise=NOISE, pending=1):
    from dotsboxesaz_b200.utils.utils import DotDict
    return DotDict({"self_play": {"reuse_mcts_tree": True, "noise": noise,
                                  "mcts": {"mcts_num_read": num_read, "mcts_cpuct": CPUCT, "temperature": dict(TEMPERATURE),
                                           "max_async_searches": pending}}})


def _n_searches(k, num_read):
    return min(4 * math.factorial(k), num_read) if k < 8 else num_read   # self_play.py:64-65


def _replay_device_history(bsp, board, num_read, games):
    """Every game of a device-resident batch again on oracle.OracleTree: same roots, same pre-drawn noise, same moves."""
    from oracle import oracle
    h = bsp._device_hist
    L, C = board
    moves = torch.stack(h["moves"]).cpu().numpy()       # [n_moves, n]
    active = torch.stack(h["active"]).cpu().numpy()
    visits = torch.stack(h["visits"]).cpu().numpy()
    stats = torch.stack(h["stats"]).cpu().numpy()       # root_N, max_deepness, tree_size, terminal_count, ...
    q = torch.stack(h["q"]).cpu().numpy()
    noise = h["noise"].cpu().numpy() if h["noise"] is not None else None
    res = h["result"].cpu().numpy()
    uniforms = h["uniforms"].cpu().numpy() if h.get("uniforms") is not None else None
    inv_temp = h["inv_temp"].cpu().numpy() if uniforms is not None else None
    searches = sims = draws = 0
    for g in games:
        tree = oracle.OracleTree(L, C)
        og = oracle.OracleGame(L, C)
        for m in range(moves.shape[0]):
            if og.result() is not None:
                assert not active[m, g] and moves[m, g] == -1
                continue
            assert active[m, g], (g, m)
            valid = og.valid_moves()
            reads = _n_searches(int(valid.sum()), num_read)
            ov = tree.search(reads, cpuct=CPUCT, noise=None if noise is None else noise[m, g] * valid, coeff=NOISE[1] if noise is not None else 0.0)
            assert np.array_equal(visits[m, g], ov), ("visit counts", g, m, visits[m, g], ov)
            r = tree.root()
            assert int(stats[m, g, 0]) == r["root_N"]
            assert [int(stats[m, g, 1]), int(stats[m, g, 2]), int(stats[m, g, 3])] == r["stats"][:3], ("tree stats", g, m)
            assert np.float32(q[m, g]) == np.float32(r["stats"][3]), ("root q", g, m)
            searches += 1
            sims += reads
            mv = int(moves[m, g])
            assert valid[mv]
            if uniforms is not None:
                # the draw of the move (k_selfplay_pick): np.random.choice's law, self_play.py:29-35, with the recorded uniform
                v = visits[m, g].astype(np.float64)
                cdf = np.cumsum((v / max(v.max(), 1.0)) ** inv_temp[m])
                target = uniforms[m, g] * cdf[-1]
                want = int(np.searchsorted(cdf, target, side="right"))
                assert want == mv or abs(cdf[min(want, mv)] - target) <= 1e-12 * cdf[-1], ("draw", g, m, want, mv)
                draws += 1
            tree.reroot(mv, reuse=True)
            og.play_(mv)
        assert og.result() == int(res[g])
    assert uniforms is None or draws == searches
    return searches, sims


@pytest.mark.parametrize("mode", ["device", "async"])
@pytest.mark.parametrize("board,n,num_read,check", [((3, 3), 256, 800, 96), ((5, 5), 64, 200, 24)])
def test_device_selfplay_replays_bit_exact_on_the_oracle(mode, board, n, num_read, check):
    """The games/hour path: whole batches of self-play games (3x3 @ 800 sims/move, 5x5 @ 200) through the production
    schedule with the deterministic fake net, then `check` games replayed on the oracle with the recorded device noise
    (noise_all[move, game]) and the sampled moves."""
    from dotsboxesaz_b200 import engine, self_play
    eng = engine.Engine(board, n_games=n, max_nodes=4096 if board == (3, 3) else 6144, eval_cache=16)
    try:
        bsp = self_play.BatchedSelfPlay(eng, engine.FakeNetEvaluator(0), _params(num_read), graph_waves=8 if mode == "device" else 4)
        assert bsp.adaptive and bsp.pending == 1
        info = (bsp.play_games_device if mode == "device" else bsp.play_games_async)(range(n), seed=11)
        assert info["errors"] == 0 and info["cache_hits"] > 0
        games = list(range(0, n, max(1, n // check)))[:check]
        searches, sims = _replay_device_history(bsp, board, num_read, games)
        assert searches >= 10 * len(games)
    finally:
        eng.close()


def test_config1_full_games_host_rng_vs_oracle():
    """64 fully-checked GAMES on 3x3 at 800 sims/move, Dirichlet (0.8, 0.25), temperature
    {0: 1.0, 12: 0.02}, tree reuse, one legacy-MT19937 stream per game (seed_g = base + g; per move: the Dirichlet draw,
    then the uniform of np.random.choice -- the reference's order), production schedule on.  The oracle plays the same
    games from the same seeds on its own: moves, visit counts of every search and results must be identical."""
    from dotsboxesaz_b200 import engine, self_play
    from oracle import oracle
    n, num_read, base = 64, 800, 4000
    eng = engine.Engine((3, 3), n_games=n, max_nodes=4096, eval_cache=16)
    try:
        bsp = self_play.BatchedSelfPlay(eng, engine.FakeNetEvaluator(0), _params(num_read), graph_waves=8)
        assert bsp.adaptive
        played = bsp.play_games(range(n), seeds=[base + g for g in range(n)])
        assert len(played) == n
        total = 0
        for g, (idx, moves, visits, z) in enumerate(played):
            rs = np.random.RandomState(base + g)
            tree, og = oracle.OracleTree(3, 3), oracle.OracleGame(3, 3)
            temperature, m = None, 0
            while og.result() is None:
                if m in TEMPERATURE:
                    temperature = TEMPERATURE[m]
                valid = og.valid_moves()
                noise = rs.dirichlet(np.ones(og.A) * NOISE[0], 1).ravel() * valid
                ov = tree.search(_n_searches(int(valid.sum()), num_read), cpuct=CPUCT, noise=noise, coeff=NOISE[1])
                assert np.array_equal(visits[m], ov), ("visit counts", g, m)
                probs = self_play._apply_temperature(ov, temperature)
                mv = int(rs.choice(og.A, 1, p=probs)[0])
                assert mv == moves[m], ("move", g, m)
                tree.reroot(mv, reuse=True)
                og.play_(mv)
                m += 1
                total += 1
            assert m == len(moves) and og.result() == z
        assert total > 15 * n
    finally:
        eng.close()


def test_reference_selfplay_fixture_at_800_sims():
    """tests/golden/selfplay.json.gz entry 3: a SelfPlay.play_game trajectory recorded from the real reference at 800
    sims/move (fake net kind 1, Dirichlet noise, temperature schedule); the batched engine path must reproduce its moves,
    visit counts, result and dataset rows from the same seed -- through the production schedule."""
    from golden_io import load
    from dotsboxesaz_b200 import engine, self_play
    from dotsboxesaz_b200.utils.utils import DotDict
    fixtures = [G for G in load("selfplay") if G["num_read"] == 800]
    assert fixtures, "no 800-sims fixture"
    for G in fixtures:
        eng = engine.Engine((G["L"], G["C"]), n_games=2, max_nodes=4096, eval_cache=14)  # one idle slot on purpose
        try:
            params = DotDict({"self_play": {"reuse_mcts_tree": True, "noise": tuple(G["noise"]),
                                            "mcts": {"mcts_num_read": G["num_read"], "mcts_cpuct": CPUCT,
                                                     "temperature": {int(k): v for k, v in G["temperature"].items()},
                                                     "max_async_searches": 1}}})
            bsp = self_play.BatchedSelfPlay(eng, engine.FakeNetEvaluator(G["kind"]), params, graph_waves=8)
            played = bsp.play_games([G["seed"]], seeds=[G["seed"]])
            idx, moves, visits, z = played[0]
            assert moves == G["moves"] and [v.tolist() for v in visits] == G["visits"] and z == G["z"]
            df = bsp.get_datasets(3, True).reset_index()
            assert list(df.columns) == G["columns"]
            assert np.array_equal(df.to_numpy(dtype=np.float64), np.array(G["rows"], dtype=np.float64))
        finally:
            eng.close()


class _LoggingEvaluator:
    """Wraps a real evaluator and remembers every (leaf state -> priors, value) it produced, so that the oracle can be
    given IDENTICAL net outputs (north_star: visit counts bit-exact given identical NN outputs)."""

    def __init__(self, ev, engine):
        self.ev, self.eng, self.log = ev, engine, []
        self.engine_launches = getattr(ev, "engine_launches", 0)

    def __call__(self, eng):
        self.ev(eng)
        self.log.append((eng.leaf_kind.clone(), eng.leaf_states.clone(), eng.priors.clone(), eng.values.clone()))

    def table(self):
        out = {}
        for kind, st, p, v in self.log:
            rows = torch.nonzero(kind > 0).reshape(-1)
            s = st[rows].cpu().numpy().view(np.uint8).reshape(-1, 32)
            pn, vn = p[rows].cpu().numpy(), v[rows].cpu().numpy()
            for i in range(len(rows)):
                out[bytes(s[i, :22])] = (pn[i].copy(), np.float32(vn[i]))   # edges, btc2, to_play, just_played
        return out


def test_config3_5x5_800sims_resnet20_bf16_vs_oracle():
    """BASELINE configs[3] at its real depth: 5x5 boxes, 800 simulations per move, ResNetZero with 20 residual blocks in
    bf16 through the tcgen05 tower kernel.  Every evaluation the net produced is logged; oracle trees fed the same
    outputs must end with identical visit counts, W and tree statistics."""
    from dotsboxesaz_b200 import engine
    from dotsboxesaz_b200.nn import FusedResNetZero, ResNetZero, resnet_zero_parameters
    from dotsboxesaz_b200.utils.utils import DotDict
    from oracle import oracle
    n, sims = 96, 800
    eng = engine.Engine((5, 5), n_games=n, max_nodes=sims + 8)
    try:
        torch.manual_seed(0)
        model = ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters((5, 5), nb_blocks=20)}}))
        plan = FusedResNetZero(model, eng, dtype=torch.bfloat16)
        assert plan.tower is not None and plan.tower[2] == 40
        plan.TOWER_MIN_ROWS = 0  # 96 leaves per wave: keep them on the tower kernel (tiny batches default to the library path)
        ev = _LoggingEvaluator(plan, eng)
        # roots: a few random plies each, built with the engine's own rules
        g = torch.Generator(device=eng.device).manual_seed(7)
        st = eng.new_states(n)
        for ply in range(10):
            legal = eng.valid_moves(st).float()
            mv = torch.multinomial(legal + 1e-9, 1, generator=g).reshape(-1).int()
            eng.play(st, torch.where(torch.arange(n, device=eng.device) % 11 > ply, mv, torch.full_like(mv, -1)))
        rs = np.random.RandomState(5)
        valid = eng.valid_moves(st).cpu().numpy()
        noise = rs.dirichlet(np.ones(eng.A) * NOISE[0], size=n) * valid
        eng.reset_roots(st)
        eng.run_search(sims, ev, noise=torch.from_numpy(noise), coeff=NOISE[1])   # plain wave loop: every leaf goes to the net, row == tree
        vis = eng.root_visits().cpu().numpy()
        W, P, S, U = (x.cpu().numpy() for x in eng.root_children())
        stats, rootW, q = (x.cpu().numpy() for x in eng.tree_stats())
        assert eng.status()["errors"] == 0
        table = ev.table()
        assert len(table) > 20000
        states_np = eng.states_to_numpy(st)
        misses = [0]
        checked = 0
        for t in range(0, n, 12):
            # the oracle tree of root t with a net that answers from the log
            root = oracle.OracleGame(5, 5)

            def lookup(og):
                b = og.board().ravel()
                e = [0, 0]
                for a in np.flatnonzero(b == 255):
                    e[int(a) >> 6] |= 1 << (int(a) & 63)
                key = np.zeros(22, dtype=np.uint8)
                key[:16] = np.frombuffer(np.array(e, dtype=np.uint64).tobytes(), dtype=np.uint8)
                key[16:20] = np.frombuffer(np.array([og.s.btc2[0], og.s.btc2[1]], dtype=np.int16).tobytes(), dtype=np.uint8)
                key[20] = og.s.to_play
                key[21] = np.uint8(og.s.just_played & 0xff)
                hit = table.get(bytes(key))
                if hit is None:
                    misses[0] += 1
                    return np.full(eng.A, 1.0 / eng.A, np.float32), np.zeros(1, np.float32)
                return hit[0], np.array([hit[1]], dtype=np.float32)
            og = _reach(root, states_np[t], eng)
            tree = oracle.OracleTree(5, 5, og.s, nn=lookup)
            ov = tree.search(sims, cpuct=CPUCT, noise=noise[t], coeff=NOISE[1])
            assert misses[0] == 0, "the oracle asked for a position the engine never evaluated"
            assert np.array_equal(vis[t], ov), ("visit counts", t)
            r = tree.root()
            assert np.array_equal(W[t], r["W"]) and np.array_equal(P[t], r["priors"])
            assert int(stats[t][0]) == r["root_N"] and [int(stats[t][1]), int(stats[t][2]), int(stats[t][3])] == r["stats"][:3]
            assert np.float32(rootW[t]) == np.float32(r["root_W"])
            checked += 1
        assert checked == 8
    finally:
        eng.close()


def _reach(og, packed, eng):
    """The oracle game with the engine's packed root state taken as is (the order in which its edges were played is not
    recoverable, and turn order matters, so the fields are copied rather than replayed)."""
    og.s.to_play = int(packed["to_play"])
    og.s.just_played = int(packed["just_played"])
    og.s.btc2[0], og.s.btc2[1] = int(packed["btc2"][0]), int(packed["btc2"][1])
    edges = [int(packed["edges"][0]), int(packed["edges"][1])]
    for a in range(eng.A):
        if (edges[a >> 6] >> (a & 63)) & 1:
            og.s.board[a] = 255
    og.s.hash_lo, og.s.hash_hi = edges[0], edges[1]
    og.s.hash_btc2 = og.s.btc2[og.s.to_play]
    return og


def test_elo_arena_model_routing_vs_oracle():
    """The Elo arena swaps the model on every move according to the player to move at the ROOT (self_play.py:237-239): a
    whole search is evaluated by that player's net.  DualEvaluator with the two deterministic fake nets: every tree must end
    with the visit counts of an oracle search that used the net owning its root -- and differ from the other net's."""
    from dotsboxesaz_b200 import engine, self_play
    from oracle import oracle
    n, sims = 64, 300
    eng = engine.Engine((3, 3), n_games=n, max_nodes=sims + 8)
    try:
        g = torch.Generator(device=eng.device).manual_seed(3)
        st = eng.new_states(n)
        for ply in range(8):
            legal = eng.valid_moves(st).float()
            mv = torch.multinomial(legal + 1e-9, 1, generator=g).reshape(-1).int()
            eng.play(st, torch.where(torch.arange(n, device=eng.device) % 9 > ply, mv, torch.full_like(mv, -1)))
        dual = self_play.DualEvaluator(engine.FakeNetEvaluator(0), engine.FakeNetEvaluator(1), eng)
        swap = torch.arange(n, device=eng.device) % 2 == 1           # colours alternate between games, as compute_elo does
        to_play = st.view(torch.uint8).reshape(n, 32)[:, 20].bool()
        dual.owner.copy_(to_play ^ swap)
        eng.reset_roots(st)
        eng.run_search(sims, dual)
        vis = eng.root_visits().cpu().numpy()
        owner = dual.owner.cpu().numpy()
        packed = eng.states_to_numpy(st)
        differ = 0
        for t in range(n):
            og = _reach(oracle.OracleGame(3, 3), packed[t], eng)
            mine = oracle.OracleTree(3, 3, og.s, kind=int(owner[t])).search(sims, cpuct=CPUCT)
            other = oracle.OracleTree(3, 3, og.s, kind=1 - int(owner[t])).search(sims, cpuct=CPUCT)
            assert np.array_equal(vis[t], mine), ("routing", t, int(owner[t]))
            differ += int(not np.array_equal(mine, other))
        assert differ > n // 2   # the two nets do lead to different searches, so the check above is not vacuous
    finally:
        eng.close()
