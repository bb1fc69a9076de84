"""Time the 'thresh' and 'iht' coders at the bench workload's shape (n=64, K=1024, 1M signals)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lyssa_oracle as lo
from lyssandra_b200 import engine

N = 1 << 20
X = torch.as_tensor(np.ascontiguousarray(lo.synthetic_patches(N, 64, seed=0)), device="cuda:0")
D = torch.as_tensor(lo.synthetic_dictionary(1024, 64, seed=1), device="cuda:0")
Xd = engine.as_device_matrix(X, X.device); Dd = engine.as_dictionary(D, X.device)
for name, fn in (("thresh k=5", lambda: engine.thresh_encode(Xd, Dd, 5)),
                 ("thresh k=5 dense", lambda: engine.thresh_encode(Xd, Dd, 5, dense=True)),
                 ("thresh k=102", lambda: engine.thresh_encode(Xd, Dd, 102)),
                 ("iht k=5 n_iter=4", lambda: engine.iht_encode(Xd, Dd, 5, 0.2, 4)),
                 ("bomp k=5", lambda: engine.bomp_encode(Xd, Dd, 5)),
                 ("bomp k=5 dense", lambda: engine.bomp_encode(Xd, Dd, 5, dense=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-22s %8.3f ms per 1M signals  (%.3g signals/s)" % (name, ms, N / ms * 1e3), flush=True)
