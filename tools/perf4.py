import sys, torch
sys.path.insert(0, '/root/repo')
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedSimpleNN
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
torch.backends.cudnn.benchmark = True
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for ng in (1024, 2048, 4096, 8192):
    eng = engine.Engine((3,3), n_games=ng, max_nodes=1024)
    torch.manual_seed(0)
    ev = FusedSimpleNN(SimpleNN(board=(3,3)), eng)
    ev(eng)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10): ev(eng)
    ms = timeit(lambda: g.replay())/10
    fk = engine.FakeNetEvaluator(0)
    def run(): eng.reset_roots(); eng.run_search(800, fk, graph_waves=16)
    t = timeit(run)/802
    def run2(): eng.reset_roots(); eng.run_search(800, ev, graph_waves=16)
    t2 = timeit(run2,2)/802
    print(f"games={ng}: net {ms*1e3:.1f} us/batch ({ms*1e3/ng*1000:.1f} ns/eval); tree+fake wave {t*1e3:.1f} us; full wave {t2*1e3:.1f} us -> {ng/t2/1e3:.2f} Msims/s", flush=True)
    eng.close()
# config 2: random rollouts 5x5, 1M games
eng = engine.Engine((5,5), n_games=1, max_nodes=4)
n = 1<<20
st0 = eng.new_states(n)
def roll():
    st = st0.clone(); return eng.random_rollout(st, seed=0)
pl = roll(); torch.cuda.synchronize()
st = st0.clone()
a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
a.record(); plies = eng.random_rollout(st, seed=0); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b); tot = int(plies.sum())
print(f"rollouts 5x5: {n} games, {tot} plies in {ms:.3f} ms -> {n/ms/1e3:.1f} M games/s, {tot/ms/1e6:.2f} G plies/s, mean plies {tot/n:.1f}; state bytes in+out {2*32*n/ms/1e6:.1f} GB/s")
