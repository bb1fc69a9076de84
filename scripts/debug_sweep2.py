import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from oracle import lyssa_oracle as lo
dev = "cuda:0"
g = dict(np.load(os.path.join(ROOT, "tests/golden/ksvd_sweep.npz")))
Xh = g["X"]; Dh = lo.synthetic_dictionary(96, 64, seed=9); Dh[:, 10] = Dh[:, 2]
K = 96; N = 400; k = 4
X = torch.from_numpy(np.ascontiguousarray(Xh)).to(dev); D = torch.from_numpy(Dh).to(dev).clone()
codes = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False).encode_sparse(X, D)
idx = codes.idx.cpu().numpy(); val = codes.val.cpu().numpy().astype(np.float64)
Zd = np.zeros((K, N)); Zd[idx, np.arange(N)[:, None]] = val
Do = Dh.astype(np.float64).copy(); lo.approx_ksvd(Xh.astype(np.float64), Do, Zd)
R, _ = engine.residual(X, D, codes, True, False)
rowptr, entries = engine.build_atom_csr(codes)
engine.approx_ksvd_sweep(R, D, codes, rowptr, entries, 1)
torch.cuda.synchronize()
err = np.abs(D.cpu().numpy() - Do).max(axis=0)
print("first bad", np.flatnonzero(err > 1e-5)[:5], "max", err.max())
