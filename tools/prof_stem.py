import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dotsboxesaz_b200 import engine
from dotsboxesaz_b200.nn import FusedSimpleNN
from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN
eng = engine.Engine((3, 3), n_games=4096, max_nodes=64)
torch.manual_seed(0)
ev = FusedSimpleNN(SimpleNN(board=(3, 3)), eng, dtype=torch.bfloat16)
st = eng.new_states(4096)
eng.random_rollout(st, seed=1)
eng.leaf_states.copy_(st)
for _ in range(6):
    ev(eng)
torch.cuda.synchronize()
print("ok")
