"""Bring-up: time the K-SVD sweep alone at cfg3 with the library named by LYSSA_B200_LIB."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyssandra_b200 import _native, engine
from lyssandra_b200.sparse_coding import sparse_encoder
from oracle import lyssa_oracle as lo
dev = torch.device("cuda", 0)
n, K, k, N = 64, 1024, 10, 2000000
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=2000).T)).to(dev).t()
D = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=1)).to(dev)
codes = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False).encode_sparse(X, D)
R, _ = engine.residual(X, D, codes, want_residual=True, want_error=False)
rowptr, entries = engine.build_atom_csr(codes)
ts = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.approx_ksvd_sweep(R, D, codes, rowptr, entries, n_cycles=1); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(os.path.basename(os.environ.get("LYSSA_B200_LIB", "default")), "sweep ms:", " ".join("%.2f" % t for t in ts), " D checksum %.6f" % float(D.double().abs().sum()))
