"""Bring-up: cycles per phase of the K-SVD sweep's atom loop (bring-up build: LYSSA_B200_LIB=.../liblyssa_b200_bringup.so)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyssandra_b200 import _native, engine
from lyssandra_b200.sparse_coding import sparse_encoder
from oracle import lyssa_oracle as lo
lib = _native.load()
dev = torch.device("cuda", 0)
n, K, k, N = 64, 1024, 10, int(os.environ.get("SW_N", 2000000))
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=2000).T)).to(dev).t()
D = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=1)).to(dev)
codes = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False).encode_sparse(X, D)
R, _ = engine.residual(X, D, codes, want_residual=True, want_error=False)
rowptr, entries = engine.build_atom_csr(codes)
out = (ctypes.c_ulonglong * 8)()
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.approx_ksvd_sweep(R, D, codes, rowptr, entries, n_cycles=1); e1.record(); torch.cuda.synchronize()
    lib.lys_debug_sweep_timing(out)
    v = [int(x) for x in out]
names = ["patch + look-ahead loads (c+2)", "rows + sums + publish (c+1)", "poll reduced sums (c)", "barrier", "new atom", "phase 2", "end barrier"]
print("N=%d sweep %.2f ms; cycles per atom (CTA %d, thread 0): total %.0f" % (N, e0.elapsed_time(e1), 74, sum(v) / K))
for nm, x in zip(names, v):
    print("   %-34s %7.0f" % (nm, x / K))
