"""Small run, for compute-sanitizer memcheck, of the kernels changed late in round 2: the fused 'thresh' kernel with the
two-pass scan and candidate lists (also sorted correlations: the prune path; k = 10: sub-piece maxima), the
double-buffered dense-row writer (bomp and thresh, tile tails), the ODL statistics kernel with the user bitmap."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200.dict_learning import online_dict_learn
from oracle import lyssa_oracle as lo
dev = "cuda:0"
for (K, N, k) in ((1024, 700, 5), (512, 333, 10), (256, 129, 1)):
    X = engine.as_device_matrix(torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, 64, seed=K))).to(dev), dev)
    D = engine.as_dictionary(torch.from_numpy(lo.synthetic_dictionary(K, 64, seed=K + 1)).to(dev), dev)
    a = engine.thresh_encode(X, D, k); b = engine.thresh_encode(X, D, k, dense=True)
    c = engine.bomp_encode(X, D, min(k, 5), dense=True)
    torch.cuda.synchronize()
    print("ok thresh/bomp dense", K, N, k, int(a.idx.sum()), float(b[1].abs().sum()), float(c[1].abs().sum()))
theta = np.linspace(1.2, 0.1, 1024); Ds = np.zeros((64, 1024), dtype=np.float32); Ds[0], Ds[1] = np.cos(theta), np.sin(theta)
Xs = np.zeros((64, 200), dtype=np.float32); Xs[0] = 1.0; Xs[:, 100:] = 0.0
s = engine.thresh_encode(engine.as_device_matrix(torch.from_numpy(Xs).to(dev), dev), engine.as_dictionary(torch.from_numpy(Ds).to(dev), dev), 5)
torch.cuda.synchronize()
print("ok thresh sorted/constant (prune path)", s.idx[0].tolist(), s.idx[150].tolist())
Xd = torch.from_numpy(np.ascontiguousarray(lo.synthetic_descriptors(2 * 1000, 128, seed=5))).to(dev)
Dd = torch.from_numpy(np.ascontiguousarray(lo.norm_cols(np.abs(np.random.default_rng(6).standard_normal((128, 512)))).astype(np.float32))).to(dev)
Do, A, B = online_dict_learn(Xd, 512, sparse_coder=sparse_encoder("bomp", {"n_nonzero_coefs": 5}, verbose=False), batch_size=1000,
                             D_init=Dd.clone(), beta=0.9, n_epochs=1)
torch.cuda.synchronize()
print("ok odl", float(Do.abs().sum()), float(A.abs().sum()))
