"""Bring-up: where the time of an ODL minibatch goes (host wall clock vs device time), cfg4 shape."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200.dict_learning import online_dict_learn
from oracle import lyssa_oracle as lo
dev = torch.device("cuda", 0)
n, K, k, b, n_mb = 128, 2048, 5, 4096, 24
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_descriptors(b * n_mb, n, seed=0).T)).to(dev).t()
rng = np.random.default_rng(1)
D0 = torch.from_numpy(np.ascontiguousarray(lo.norm_cols(np.abs(rng.standard_normal((n, K)))).astype(np.float32))).to(dev)
enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
for rep in range(3):
    D = D0.clone(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    online_dict_learn(X, K, sparse_coder=enc, batch_size=b, D_init=D, beta=0.9, n_epochs=1)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("online_dict_learn: host %.3f ms/minibatch, with device drain %.3f ms/minibatch" % ((t1 - t0) * 1e3 / n_mb, (t2 - t0) * 1e3 / n_mb))
D = D0.clone(); A = torch.zeros((K, K), device=dev); B = torch.zeros((n, K), device=dev)
Xb = X[:, :b]
def wall(name, fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("  %-28s host %.3f ms  device-bound %.3f ms" % (name, (t1 - t0) * 1e3 / reps, (t2 - t0) * 1e3 / reps))
    return out
wall("gram", lambda: engine.gram(D))
codes = wall("encode_sparse (incl. gram)", lambda: enc.encode_sparse(Xb, D))
wall("odl_accumulate_", lambda: engine.odl_accumulate_(Xb, codes, 0.9, A, B))
wall("odl_update_dict_", lambda: engine.odl_update_dict_(D, A, B))
