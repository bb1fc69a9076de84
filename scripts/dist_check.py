"""Multi-GPU check (launch with torchrun, one rank per GPU): the sharded paths against the
single-GPU result on identical data.
  1. encode: every rank's shard == the same columns of a full encode (no collective)
  2. approximate K-SVD sweep with the in-kernel peer all-reduce == single-GPU sweep
  3. ODL sufficient statistics all-reduced over ranks == single-GPU statistics
Prints one JSON line from rank 0 and exits non-zero on mismatch."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import engine
from lyssandra_b200.distributed import DistContext, PeerExchange
from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200.dict_learning import approx_ksvd, ksvd_dict_learn, online_dict_learn
from oracle import lyssa_oracle as lo

ctx = DistContext.init_from_env()
dev = torch.device("cuda", torch.cuda.current_device())
n, K, k, N = 64, 1024, 10, 400000
Xh = np.ascontiguousarray(lo.synthetic_patches(N, n, seed=0).T)           # same data on every rank
Dh = lo.synthetic_dictionary(K, n, seed=1)
Xfull = torch.from_numpy(Xh).to(dev).t()
lo_, hi_ = ctx.shard(N)
Xloc = Xfull[:, lo_:hi_]
enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
report = {"world": ctx.world}

# 1. encode shard invariance
D = torch.from_numpy(Dh).to(dev)
full = enc.encode_sparse(Xfull, D)
loc = enc.encode_sparse(Xloc, D)
ok1 = bool(torch.equal(loc.idx, full.idx[lo_:hi_]) and torch.equal(loc.val, full.val[lo_:hi_]))
report["encode_shard_equal"] = ok1

# 2. sweep: sharded + peer exchange vs single GPU
ex = PeerExchange(ctx)
D1 = torch.from_numpy(Dh).to(dev).clone(); c1 = enc.encode_sparse(Xfull, D1)
_, _, un1 = approx_ksvd(Xfull, D1, c1)
D2 = torch.from_numpy(Dh).to(dev).clone(); c2 = enc.encode_sparse(Xloc, D2)
torch.cuda.synchronize(); ctx.barrier()
t0 = time.perf_counter()
_, _, un2 = approx_ksvd(Xloc, D2, c2, comm=ex.handle)
torch.cuda.synchronize(); t_sweep = time.perf_counter() - t0
dD = float((D1 - D2).abs().max()); dv = float((c1.val[lo_:hi_] - c2.val).abs().max())
report.update({"sweep_max_dD": dD, "sweep_max_dval": dv, "sweep_unused_equal": un1 == un2, "sweep_s_sharded": t_sweep})
# every rank must hold the same dictionary bit for bit
Dsum = D2.clone(); ctx.allreduce_sum_(Dsum)
report["sweep_D_identical_across_ranks"] = bool(torch.equal(Dsum, D2 * ctx.world)) if ctx.world in (1, 2, 4, 8) else None
ok2 = dD < 1e-4 and dv < 1e-3 and un1 == un2

# 2b. full learner, 3 iterations, objective parity
h1, h2 = [], []
ksvd_dict_learn(Xfull, K, init_dict=torch.from_numpy(Dh).to(dev), sparse_coder=enc, max_iter=3, approx=True, verbose=False, return_codes=True, history=h1)
ksvd_dict_learn(Xloc, K, init_dict=torch.from_numpy(Dh).to(dev), sparse_coder=enc, max_iter=3, approx=True, verbose=False, return_codes=True,
                history=h2, dist=ctx, exchange=ex)
report["ksvd_errors_single"] = [h["error"] for h in h1]; report["ksvd_errors_sharded"] = [h["error"] for h in h2]
ok2b = all(abs(a["error"] - b["error"]) <= 1e-3 * a["error"] for a, b in zip(h1, h2))

# 2c. 'data' initialisation drawn over the global column space: same seed -> the sharded run starts from (and keeps) the
# dictionary the single-GPU run on the whole set gets (dict_learning/utils.py:55-70 through init_dictionary_sharded)
h3, h4 = [], []
np.random.seed(77)
Da1, _ = ksvd_dict_learn(Xfull, 256, init_dict="data", sparse_coder=enc, max_iter=2, approx=True, verbose=False, return_codes=True, history=h3)
np.random.seed(77)
Da2, _ = ksvd_dict_learn(Xloc, 256, init_dict="data", sparse_coder=enc, max_iter=2, approx=True, verbose=False, return_codes=True,
                         history=h4, dist=ctx, exchange=ex)
report["ksvd_data_init_max_dD"] = float((Da1 - Da2).abs().max())
report["ksvd_data_init_errors"] = [[h["error"] for h in h3], [h["error"] for h in h4]]
ok2c = report["ksvd_data_init_max_dD"] < 1e-4 and all(abs(a["error"] - b["error"]) <= 1e-3 * a["error"] for a, b in zip(h3, h4))

# 3. ODL statistics
n4, K4, b4 = 64, 512, 8192
Xo = Xfull[:, :b4]; Do = torch.from_numpy(lo.synthetic_dictionary(K4, n4, seed=5)).to(dev)
enc5 = sparse_encoder("bomp", {"n_nonzero_coefs": 5}, verbose=False)
Da, Aa, Ba = online_dict_learn(Xo, K4, sparse_coder=enc5, batch_size=2048, D_init=Do.clone(), beta=0.9, n_epochs=1)
l2, h2_ = ctx.shard(2048)
# sharded: every rank takes its slice of each minibatch
Xparts = torch.cat([Xo[:, m * 2048 + l2: m * 2048 + h2_] for m in range(4)], dim=1)
Db, Ab, Bb = online_dict_learn(Xparts, K4, sparse_coder=enc5, batch_size=h2_ - l2, D_init=Do.clone(), beta=0.9, n_epochs=1,
                               dist=ctx)
dA = float((Aa - Ab).abs().max() / Aa.abs().max()); dDo = float((Da - Db).abs().max())
report.update({"odl_rel_dA": dA, "odl_max_dD": dDo, "odl_bitwise_equal_to_1gpu": bool(torch.equal(Da, Db) and torch.equal(Aa, Ab) and torch.equal(Ba, Bb))})
ok3 = dA < 1e-6 and dDo < 1e-6
ex.close()
okall = torch.tensor([int(ok1 and ok2 and ok2b and ok2c and ok3)], device=dev); ctx.allreduce_sum_(okall)
if ctx.rank == 0:
    report["all_ok"] = int(okall.item()) == ctx.world
    print(json.dumps(report))
if ctx.world > 1:
    torch.distributed.destroy_process_group()
sys.exit(0 if int(okall.item()) == ctx.world else 1)
