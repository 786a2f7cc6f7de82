import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtc_b200 import ops
from vtc_b200.synthetic import make_retrieval_pair
T, V = make_retrieval_pair(10000, 125000, 512, seed=1023)
q, g = T.cuda(), V.cuda()
for _ in range(3):
    v, i = ops.sim_topk(q, g, 11, precision="bf16")
torch.cuda.synchronize()
print(v[0], i[0])
