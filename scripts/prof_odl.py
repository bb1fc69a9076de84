"""ncu driver: a few ODL minibatches at the cfg4 shape (encode -> statistics -> dictionary update)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200.dict_learning import online_dict_learn
from oracle import lyssa_oracle as lo
dev = torch.device("cuda", 0)
n, K, k, b, n_mb = 128, 2048, 5, 4096, 3
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_descriptors(b * n_mb, n, seed=0).T)).to(dev).t()
D0 = torch.from_numpy(np.ascontiguousarray(lo.norm_cols(np.abs(np.random.default_rng(1).standard_normal((n, K)))).astype(np.float32))).to(dev)
enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
online_dict_learn(X, K, sparse_coder=enc, batch_size=b, D_init=D0.clone(), beta=0.9, n_epochs=1)
torch.cuda.synchronize()
print("done")
