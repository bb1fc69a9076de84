"""Secondary measurements: one approximate-K-SVD iteration at BASELINE cfg3 (2M 8x8 patches, K=1024, k=10)
split into its stages, and one ODL minibatch at cfg4 (n=128, K=2048, b=4096, k=5).  CUDA events, warm-up 1,
median of R.  Prints one JSON line per workload."""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.sparse_coding import sparse_encoder
from oracle import lyssa_oracle as lo

ap = argparse.ArgumentParser()
ap.add_argument("--signals", type=int, default=2 * 1000 * 1000)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--skip-odl", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


# ---- cfg3: one K-SVD iteration
n, K, k, N = 64, 1024, 10, a.signals
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=0).T)).to(dev).t()
D = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=1)).to(dev)
enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
stages = {"encode": [], "residual": [], "csr": [], "sweep": [], "error": []}
for it in range(a.iters + 1):
    t_enc, codes = timed(lambda: enc.encode_sparse(X, D), 1)
    t_res, (R, _) = timed(lambda: engine.residual(X, D, codes, want_residual=True, want_error=False), 1)
    t_csr, (rowptr, entries) = timed(lambda: engine.build_atom_csr(codes), 1)
    t_swp, flags = timed(lambda: engine.approx_ksvd_sweep(R, D, codes, rowptr, entries, n_cycles=1), 1)
    t_err, (_, err) = timed(lambda: engine.residual(X, D, codes, want_residual=False, want_error=True), 1)
    if it > 0:
        for key, v in zip(stages, (t_enc, t_res, t_csr, t_swp, t_err)):
            stages[key].append(v)
    last_err = float(err.item()); n_unused = int(flags.sum().item())
med = {key: float(np.median(v)) for key, v in stages.items()}
total = sum(med.values())
bytes_iter = N * ((4 * n + 8 * k) + 8 * n + 8 * n * k)
print(json.dumps({"workload": "approx K-SVD iteration, cfg3: %d patches, K=1024, k=10, 1 GPU" % N, "ms_per_iter": total,
                  "stages_ms": med, "us_per_atom_step": med["sweep"] * 1e3 / K, "algorithmic_GB_per_iter": bytes_iter / 1e9,
                  "achieved_GBps": bytes_iter / 1e9 / (total / 1e3), "error": last_err, "unused_atoms": n_unused}))
del X, R, codes, entries
torch.cuda.empty_cache()

# ---- cfg4: ODL minibatch
if not a.skip_odl:
    n, K, k, b = 128, 2048, 5, 4096
    Xh = np.ascontiguousarray(lo.synthetic_descriptors(b * 8, n, seed=0).T)
    X = torch.from_numpy(Xh).to(dev).t()
    rng = np.random.default_rng(1)
    D = torch.from_numpy(np.ascontiguousarray(lo.norm_cols(np.abs(rng.standard_normal((n, K)))).astype(np.float32))).to(dev)
    A = torch.zeros((K, K), device=dev); B = torch.zeros((n, K), device=dev)
    enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
    st = {"encode": [], "accumulate": [], "update": []}
    for it in range(9):
        Xb = X[:, (it % 8) * b:(it % 8 + 1) * b]
        t1, codes = timed(lambda: enc.encode_sparse(Xb, D), 1)
        t2, _ = timed(lambda: engine.odl_accumulate_(Xb, codes, 0.9, A, B), 1)
        t3, _ = timed(lambda: engine.odl_update_dict_(D, A, B), 1)
        if it > 0:
            st["encode"].append(t1); st["accumulate"].append(t2); st["update"].append(t3)
    med = {key: float(np.median(v)) for key, v in st.items()}
    print(json.dumps({"workload": "ODL minibatch, cfg4: n=128, K=2048, b=4096, k=5, 1 GPU", "ms_per_minibatch": sum(med.values()),
                      "stages_ms": med, "minibatches_per_s": 1e3 / sum(med.values())}))
