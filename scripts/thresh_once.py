"""One call of each thresholding coder (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lyssa_oracle as lo
from lyssandra_b200 import engine
N = 1 << 20
X = torch.as_tensor(np.ascontiguousarray(lo.synthetic_patches(N, 64, seed=0)), device="cuda:0")
D = torch.as_tensor(lo.synthetic_dictionary(1024, 64, seed=1), device="cuda:0")
Xd = engine.as_device_matrix(X, X.device); Dd = engine.as_dictionary(D, X.device)
engine.thresh_encode(Xd, Dd, 5)
engine.iht_encode(Xd, Dd, 5, 0.2, 1)
torch.cuda.synchronize()
