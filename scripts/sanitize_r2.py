"""Small run of the kernels that are new in round 2 for compute-sanitizer (memcheck / racecheck / synccheck):
look-ahead K-SVD sweep (also streamed users and two cycles), exact K-SVD sweep, omp coder, screen-mode fused encode,
order-deterministic ODL statistics and the tcgen05 D.A GEMM, top-k selection on a given Alpha."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200.dict_learning import approx_ksvd, ksvd, online_dict_learn
from lyssandra_b200.feature_encoding import soft_thresholding
from oracle import lyssa_oracle as lo
dev = "cuda:0"
for (n, K, N, k, cyc) in ((64, 256, 6000, 5, 1), (64, 32, 60000, 6, 2), (128, 512, 3000, 4, 1), (33, 100, 2000, 3, 1)):
    X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=1))).to(dev)
    D = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=2)).to(dev)
    enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
    codes = enc.encode_sparse(X, D)
    D2 = D.clone()
    approx_ksvd(X, D2, codes, n_cycles=cyc, verbose=False)
    torch.cuda.synchronize()
    print("ok sweep", n, K, N, k, float(D2.abs().sum()))
    if n <= 64:
        codes = enc.encode_sparse(X, D); D3 = D.clone()
        ksvd(X, D3, codes, n_cycles=1, verbose=False)
        torch.cuda.synchronize()
        print("ok exact", n, K, N, k, float(D3.abs().sum()))
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(3000, 64, seed=3))).to(dev)
D = torch.from_numpy(lo.synthetic_dictionary(1024, 64, seed=4)).to(dev)
a = engine.bomp_encode(X, D, 5, screen=True, dense=True)
b = engine.omp_encode(X, D, 5); c = engine.omp_encode(X[:, :500], D, None, tol=1.0)
Z = soft_thresholding(torch.rand((300, 1000), device=dev) - 0.5, n_nonzero_coefs=7)
torch.cuda.synchronize()
print("ok screen/omp/topk", int(a[0].idx.sum()), int(b.idx.sum()), int(c.nsel.sum()), float(Z.abs().sum()))
Xd = torch.from_numpy(np.ascontiguousarray(lo.synthetic_descriptors(3 * 1024, 128, seed=5))).to(dev)
Dd = torch.from_numpy(np.ascontiguousarray(lo.norm_cols(np.abs(np.random.default_rng(6).standard_normal((128, 512)))).astype(np.float32))).to(dev)
Do, A, B = online_dict_learn(Xd, 512, sparse_coder=sparse_encoder("bomp", {"n_nonzero_coefs": 5}, verbose=False), batch_size=1024,
                             D_init=Dd.clone(), beta=0.9, n_epochs=1)
torch.cuda.synchronize()
print("ok odl", float(Do.abs().sum()), float(A.abs().sum()))
