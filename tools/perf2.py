import sys, torch
sys.path.insert(0, '/root/repo')
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedSimpleNN
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.benchmark = True
eng = engine.Engine((3,3), n_games=4096, max_nodes=64)
torch.manual_seed(0)
ev = FusedSimpleNN(SimpleNN(board=(3,3)), eng, dtype=torch.bfloat16)
for _ in range(5): ev(eng)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(20): ev(eng)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
