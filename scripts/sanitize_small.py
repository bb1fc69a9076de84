"""Small end-to-end run of every kernel family for compute-sanitizer (racecheck / memcheck / synccheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200.dict_learning import approx_ksvd, online_dict_learn
from oracle import lyssa_oracle as lo
dev = "cuda:0"
for (n, K, N, k) in ((64, 1024, 3000, 5), (64, 256, 1000, 10), (128, 2048, 500, 5), (33, 300, 400, 7)):
    X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=1))).to(dev)
    Dh = lo.synthetic_dictionary(K, n, seed=2); Dh[:, 7] = Dh[:, 3]
    D = torch.from_numpy(Dh).to(dev)
    enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
    Z = enc.encode(X, D)
    codes = enc.encode_sparse(X, D)
    D2 = D.clone()
    approx_ksvd(X, D2, codes, n_cycles=2, verbose=False)
    online_dict_learn(X, K, sparse_coder=enc, batch_size=N // 3, D_init=D.clone(), beta=0.5, n_epochs=1)
    torch.cuda.synchronize()
    print("ok", n, K, N, k, float(Z.abs().sum()), float(D2.abs().sum()))
