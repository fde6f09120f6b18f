import sys, time, torch
sys.path.insert(0, '/root/repo')
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import DeviceEvaluator, ResNetZero, resnet_zero_parameters
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

for board, ng, mn in (((3,3),4096,2048),((3,3),16384,2048),((5,5),4096,2048)):
    eng = engine.Engine(board, n_games=ng, max_nodes=mn)
    sims=800
    for gw in (0, 8, 32):
        ev = engine.FakeNetEvaluator(0)
        def run():
            eng.reset_roots(); eng.run_search(sims, ev, graph_waves=gw)
        ms = timeit(run)
        st = eng.status()
        print(f"board={board} games={ng} fake-net graph_waves={gw}: {ms:.1f} ms/search  {ng*sims/ms/1e3:.2f} Msims/s  P={st['path_nodes']/max(1,st['sims']):.2f} maxnodes={st['max_nodes_used']}", flush=True)
    if board==(3,3):
        for dt in (torch.bfloat16, torch.float32):
            torch.manual_seed(0)
            ev = DeviceEvaluator(SimpleNN(board=board), eng, dtype=dt, channels_last=True)
            for gw in (0, 16):
                def run():
                    eng.reset_roots(); eng.run_search(sims, ev, graph_waves=gw)
                ms = timeit(run, 2)
                print(f"board={board} games={ng} SimpleNN {dt} graph_waves={gw}: {ms:.1f} ms/search  {ng*sims/ms/1e3:.3f} Msims/s", flush=True)
            # net alone
            ms = timeit(lambda: [ev(eng) for _ in range(50)])/50
            print(f"   net forward alone: {ms*1e3:.1f} us per batch of {ng} -> {ng/ms/1e3:.2f} M evals/s", flush=True)
        eng.set_planes(torch.float32, False)
    eng.close()
