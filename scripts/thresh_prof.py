"""One fused 'thresh' call at the bench shape (for ncu -k regex:bomp_tc_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lyssa_oracle as lo
from lyssandra_b200 import engine
N = 1 << 20
X = engine.as_device_matrix(torch.as_tensor(np.ascontiguousarray(lo.synthetic_patches(N, 64, seed=0)), device="cuda:0"), None)
D = engine.as_dictionary(torch.as_tensor(lo.synthetic_dictionary(1024, 64, seed=1), device="cuda:0"), X.device)
for _ in range(2):
    engine.thresh_encode(X, D, 5, dense=True)
torch.cuda.synchronize()
