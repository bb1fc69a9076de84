"""Small driver for ncu captures: W warm-up encodes + S profiled encodes of the bench workload
(optionally fewer signals).  Not a benchmark — numbers printed under a profiler are never
bench values."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyssandra_b200 import _native  # noqa: E402
from oracle import lyssa_oracle as lo  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--signals", type=int, default=1 << 20)
ap.add_argument("--atoms", type=int, default=1024)
ap.add_argument("--features", type=int, default=64)
ap.add_argument("--k", type=int, default=5)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--sparse-only", action="store_true")
a = ap.parse_args()

lib = _native.load()
dev = torch.device("cuda", 0)
n, K, k, N = a.features, a.atoms, a.k, a.signals
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=0).T)).to(dev).t()
D = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=1)).to(dev)
idx = torch.empty((N, k), dtype=torch.int32, device=dev); val = torch.empty((N, k), device=dev)
nsel = torch.empty((N,), dtype=torch.int32, device=dev)
Zt = None if a.sparse_only else torch.empty((N, K), device=dev)
G = torch.empty((K, K), device=dev)
wsb = lib.lys_bomp_workspace_bytes(n, K, N, k)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for _ in range(a.warmup + a.steps):
    _native.check(lib.lys_gram(D.data_ptr(), K, n, K, G.data_ptr(), st))
    _native.check(lib.lys_bomp_encode(X.data_ptr(), X.stride(0), X.stride(1), D.data_ptr(), K, G.data_ptr(), n, K, N, k,
                                      idx.data_ptr(), val.data_ptr(), nsel.data_ptr(),
                                      Zt.data_ptr() if Zt is not None else None, 1, K, ws.data_ptr(), wsb, st))
torch.cuda.synchronize()
print("launches per encode:", lib.lys_bomp_launch_count(n, K, N, k) + 1)
