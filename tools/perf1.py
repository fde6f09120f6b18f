import sys, time, torch
sys.path.insert(0, '/root/repo')
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import DeviceEvaluator, FusedSimpleNN, FusedResNetZero, ResNetZero, resnet_zero_parameters
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
from dotsboxesaz_b200.utils.utils import DotDict
torch.backends.cudnn.benchmark = True
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for board, ng in (((3,3),4096),((5,5),4096)):
    eng = engine.Engine(board, n_games=ng, max_nodes=1024)
    sims=800
    torch.manual_seed(0)
    nets = {"simple": SimpleNN(board=board)} if board==(3,3) else {}
    nets["resnet"] = ResNetZero(DotDict({"nn": {"model_parameters": resnet_zero_parameters(board)}}))
    for name, model in nets.items():
        for label, ev in (("eager", DeviceEvaluator(model, eng, dtype=torch.bfloat16)), ("fused", (FusedSimpleNN if name=="simple" else FusedResNetZero)(model, eng, dtype=torch.bfloat16))):
            ev(eng)
            ms = timeit(lambda: [ev(eng) for _ in range(20)])/20
            print(f"board={board} {name} {label}: net alone {ms*1e3:.1f} us/batch {ng/ms/1e3:.2f} M evals/s", flush=True)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(10): ev(eng)
            ms = timeit(lambda: g.replay())/10
            print(f"      graphed: {ms*1e3:.1f} us/batch", flush=True)
            def run():
                eng.reset_roots(); eng.run_search(sims, ev, graph_waves=16)
            ms = timeit(run, 2)
            print(f"      search: {ms:.1f} ms  {ng*sims/ms/1e3:.3f} Msims/s", flush=True)
    eng.close()
