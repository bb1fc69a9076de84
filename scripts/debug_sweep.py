import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from oracle import lyssa_oracle as lo
dev = "cuda:0"
def run(Xh, Dh, k, tag):
    K = Dh.shape[1]; N = Xh.shape[1]
    X = torch.from_numpy(np.ascontiguousarray(Xh)).to(dev); D = torch.from_numpy(np.ascontiguousarray(Dh)).to(dev).clone()
    codes = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False).encode_sparse(X, D)
    idx = codes.idx.cpu().numpy(); val = codes.val.cpu().numpy().astype(np.float64)
    Zd = np.zeros((K, N)); Zd[idx, np.arange(N)[:, None]] = val
    Do = Dh.astype(np.float64).copy()
    _, _, unused = lo.approx_ksvd(Xh.astype(np.float64), Do, Zd)
    R, _ = engine.residual(X, D, codes, True, False)
    rowptr, entries = engine.build_atom_csr(codes)
    flags = engine.approx_ksvd_sweep(R, D, codes, rowptr, entries, 1)
    torch.cuda.synchronize()
    err = np.abs(D.cpu().numpy() - Do).max(axis=0)
    bad = np.flatnonzero(err > 1e-5)
    rp = rowptr.cpu().numpy()
    print(tag, "N", N, "K", K, "unused", unused, "gpu unused", np.flatnonzero(flags.cpu().numpy()).tolist(), "first bad", bad[:5], "n_bad", len(bad), "max", err.max())
g = dict(np.load(os.path.join(ROOT, "tests/golden/ksvd_sweep.npz")))
run(g["X"], g["D"], 4, "golden-dups")
D2 = lo.synthetic_dictionary(96, 64, seed=9)
run(g["X"], D2, 4, "no-dups")
D3 = D2.copy(); D3[:, 10] = D3[:, 2]
run(g["X"], D3, 4, "dup-at-10")
run(np.ascontiguousarray(lo.synthetic_patches(6000, 64, seed=31)), lo.synthetic_dictionary(256, 64, seed=32), 6, "larger")
D4 = lo.synthetic_dictionary(256, 64, seed=32); D4[:, 100] = D4[:, 5]
run(np.ascontiguousarray(lo.synthetic_patches(6000, 64, seed=31)), D4, 6, "larger-dup100")
