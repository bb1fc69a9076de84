"""ncu driver: one approximate-K-SVD sweep at cfg3 scale (fewer signals by default)."""
import argparse, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from oracle import lyssa_oracle as lo
ap = argparse.ArgumentParser(); ap.add_argument("--signals", type=int, default=2000000); ap.add_argument("--k", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda", 0)
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(a.signals, 64, seed=0).T)).to(dev).t()
D = torch.from_numpy(lo.synthetic_dictionary(1024, 64, seed=1)).to(dev)
codes = sparse_encoder("bomp", {"n_nonzero_coefs": a.k}, verbose=False).encode_sparse(X, D)
R, _ = engine.residual(X, D, codes, True, False)
rowptr, entries = engine.build_atom_csr(codes)
engine.approx_ksvd_sweep(R, D, codes, rowptr, entries, 1)
torch.cuda.synchronize()
print("done")
