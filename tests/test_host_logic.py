"""CPU: host-side logic of the reference-facing layer — error behaviour pinned by the
reference's own tests, column partitions, the near-tie policy helper, and the world_size-2
(gloo) sharding arithmetic used by the multi-GPU paths."""
import os
import sys

import numpy as np
import pytest
import torch

from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200 import utils as lutils
from oracle import lyssa_oracle as lo

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity  # noqa: E402


def test_invalid_encoder_raises_like_the_reference():
    # lyssa/tests/test_sparse_coding.py:10-19
    X = np.random.rand(10, 100)
    D = np.random.rand(10, 4)
    se = sparse_encoder(algorithm="se", params={"n_nonzero_coefs": 4})
    with pytest.raises(ValueError, match="Sparse optimizer not found"):
        se.encode(X, D)


def test_other_reference_coders_do_not_fall_back_to_cpu():
    for alg in ("nnomp", "group_omp", "sparse_group_omp", "somp", "lasso", "llc"):
        with pytest.raises(NotImplementedError):
            sparse_encoder(algorithm=alg, params={"n_nonzero_coefs": 4}).encode(np.zeros((4, 3)), np.eye(4))
    with pytest.raises(ValueError):
        sparse_encoder(algorithm="bomp", params={}).encode(np.zeros((4, 3)), np.eye(4))
    with pytest.raises(ValueError, match="n_nonzero_coefs.*tol"):
        sparse_encoder(algorithm="omp", params={}).encode(np.zeros((4, 3)), np.eye(4))
    # 'omp' (the reference's default algorithm) is served by the device library: without a GPU it fails loudly
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            sparse_encoder(algorithm="omp", params={"n_nonzero_coefs": 2}).encode(np.zeros((4, 3), dtype=np.float32), np.eye(4, dtype=np.float32))


def test_thresholding_coder_parameters():
    # nonzero_percentage -> floor(p * n_atoms) (sparse_coding.py:419-420), and it wins over n_nonzero_coefs
    assert sparse_encoder("thresh", {"nonzero_percentage": 0.1})._check(256) == 25
    assert sparse_encoder("thresh", {"nonzero_percentage": 0.1, "n_nonzero_coefs": 3})._check(100) == 10
    assert sparse_encoder("thresh", {"n_nonzero_coefs": 3})._check(100) == 3
    with pytest.raises(ValueError):
        sparse_encoder("thresh", {})._check(100)
    with pytest.raises(ValueError, match="eta"):
        sparse_encoder("iht", {"n_nonzero_coefs": 3})._check(100)
    with pytest.raises(ValueError, match="disagree"):
        sparse_encoder("iht", {"n_nonzero_coefs": 3, "nonzero_percentage": 0.1, "eta": 0.1, "n_iter": 2})._check(100)
    assert sparse_encoder("iht", {"n_nonzero_coefs": 3, "eta": 0.1, "n_iter": 2})._check(100) == 3


def test_public_attributes_are_mutable_like_the_reference():
    se = sparse_encoder(algorithm="bomp", params={"n_nonzero_coefs": 3}, n_jobs=2, verbose=True)
    se.mmap = True; se.verbose = False; se.params["n_nonzero_coefs"] = 5; se.n_jobs = 1
    assert (se.algorithm, se.name, se.mmap, se.verbose, se.n_jobs) == ("bomp", "sparse_coder", True, False, 1)
    assert sparse_encoder().algorithm == "omp" and sparse_encoder().params == {}


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_bomp_without_gpu_fails_loudly():
    se = sparse_encoder(algorithm="bomp", params={"n_nonzero_coefs": 2})
    with pytest.raises(RuntimeError, match="CUDA|no CPU path|status"):
        se.encode(np.zeros((4, 3), dtype=np.float32), np.eye(4, dtype=np.float32))


@pytest.mark.parametrize("N,nb", [(250, 100), (50, 100), (1000, 7)])
def test_partitions_match_the_oracle(N, nb):
    a = [(r.start, r.stop) for r in lutils.gen_even_batches(N, nb)]
    b = [(r.start, r.stop) for r in lo.gen_even_batches(N, nb)]
    assert a == b
    assert [(r.start, r.stop) for r in lutils.gen_batches(N, 64)] == [(r.start, r.stop) for r in lo.gen_batches(N, 64)]
    assert lutils.gen_batches(N) == [range(0, N)]


def test_shard_bounds_cover_all_columns():
    for N, W in ((1000, 8), (1001, 8), (7, 2), (1 << 20, 4)):
        spans = [lutils.shard_bounds(N, W, r) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == N
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))


def test_parity_policy_helper():
    idx_ref = np.array([[1, 4, 9], [2, 3, -1]]); val_ref = np.array([[1.0, -2.0, 0.5], [3.0, 1.0, 0.0]])
    idx_gpu = np.array([[9, 1, 4], [3, 2, -1]]); val_gpu = np.array([[0.5, 1.0, -2.0], [1.0, 3.0, 0.0]])
    rep = parity.check_codes(idx_gpu, val_gpu, idx_ref, val_ref)
    assert rep["compared"] == 2 and rep["coef_rel_inf"] == 0.0
    with pytest.raises(AssertionError):
        parity.check_codes(np.array([[9, 1, 5]]), val_gpu[:1], idx_ref[:1], val_ref[:1])
    ok = parity.comparable_columns(np.array([[1e-3, 1e-7], [1e-2, 1e-2]]), np.ones((2, 2)), np.array([2, 2]), 2)
    assert list(ok) == [False, True]


def _gloo_worker(rank, world, port, N, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lyssandra_b200 import distributed as ldist
    ctx = ldist.DistContext.from_env_or_group()
    lo_, hi_ = ctx.shard(N)
    # sufficient statistics summed over ranks == statistics of the whole batch
    rng = np.random.default_rng(0)
    Z = torch.from_numpy(rng.standard_normal((6, N)))
    X = torch.from_numpy(rng.standard_normal((4, N)))
    A = Z[:, lo_:hi_] @ Z[:, lo_:hi_].T
    B = X[:, lo_:hi_] @ Z[:, lo_:hi_].T
    ctx.allreduce_sum_(A); ctx.allreduce_sum_(B)
    cnt = torch.tensor([hi_ - lo_], dtype=torch.int64)
    ctx.allreduce_sum_(cnt)
    gathered = ctx.allgather_bytes(bytes([rank]) * 4)
    # the ODL minibatch exchange: every rank ends with the whole minibatch (signals + codes) in rank order
    from lyssandra_b200 import engine
    from lyssandra_b200.dict_learning.online_dict_learn import _gather_minibatch
    b_loc, k, n = 5, 3, 4
    Xb = torch.arange(n * b_loc, dtype=torch.float32).reshape(n, b_loc) + 100 * rank
    codes = engine.SparseCodes(torch.full((b_loc, k), rank, dtype=torch.int32), torch.full((b_loc, k), float(rank)),
                               torch.full((b_loc,), k, dtype=torch.int32), 7)
    Xall, call = _gather_minibatch(ctx, Xb, codes)
    want_X = torch.cat([torch.arange(n * b_loc, dtype=torch.float32).reshape(n, b_loc) + 100 * r for r in range(world)], dim=1)
    ok_mb = (tuple(Xall.shape) == (n, world * b_loc) and torch.equal(Xall, want_X) and call.n_atoms == 7
             and torch.equal(call.idx[:, 0], torch.arange(world, dtype=torch.int32).repeat_interleave(b_loc))
             and torch.equal(call.val[:, 0], torch.arange(world, dtype=torch.float32).repeat_interleave(b_loc)))
    # 'data' initialisation and unused-atom replacement over the GLOBAL column space (dict_learning/utils.py:55-70,
    # ksvd.py:199-207): rank 0's RNG draws what a single process holding the concatenated shards draws, and every
    # chosen column comes from its owner — the sharded dictionary equals the single-process one bit for bit
    from lyssandra_b200.dict_learning import utils as dlu
    rng2 = np.random.default_rng(7)
    Xall_h = rng2.standard_normal((5, 41)).astype(np.float32)
    Xall_h[:, [3, 17, 40]] = 0.0                                     # non-candidates (:55) on both shards
    bounds = [0, 19, 41]                                             # uneven shards
    Xloc = torch.from_numpy(np.ascontiguousarray(Xall_h[:, bounds[rank]:bounds[rank + 1]]))

    def normalize_(M):                                               # stands in for the device norm_cols kernel
        M /= (M.norm(dim=0, keepdim=True) + 1e-10)
        return M
    np.random.seed(123)
    Dsh, unused_sh, offs = dlu.init_dictionary_sharded(ctx, Xloc, 6, normalize_=normalize_)
    Dsh_before = Dsh.clone()
    unused_sh = dlu.replace_unused_atoms_sharded(ctx, Xloc, Dsh, [4, 1], unused_sh, offs, normalize_=normalize_)
    # the single-process run on the concatenated data, same seed, the reference's calls
    np.random.seed(123)
    cands = np.flatnonzero((Xall_h.astype(np.float64) ** 2).sum(axis=0) > 1e-6)
    chosen = cands[np.random.choice(len(cands), size=6, replace=False)]
    D1 = normalize_(torch.from_numpy(Xall_h[:, chosen].copy()))
    un1 = np.array([c for c in cands if c not in set(chosen)])
    ok_init = torch.equal(Dsh_before, D1) and list(offs) == bounds
    for slot in [4, 1]:
        pos = np.random.choice(len(un1), size=1)[0]
        D1[:, slot:slot + 1] = normalize_(torch.from_numpy(Xall_h[:, [un1[pos]]].copy()))
        un1 = np.delete(un1, pos)
    ok_repl = torch.equal(Dsh, D1) and np.array_equal(unused_sh, un1)
    ok = (torch.allclose(A, Z @ Z.T) and torch.allclose(B, X @ Z.T) and int(cnt) == N
          and gathered == [bytes([r]) * 4 for r in range(world)] and ok_mb and ok_init and ok_repl)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding_allreduce_and_minibatch_gather():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, 1001, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_product_code_never_touches_the_oracle_or_the_reference_tree():
    """the oracle is test infrastructure: nothing under lyssandra_b200/ or lyssa/ may import it, name the reference
    tree, or carry a CPU fallback for a coder (the C-ABI is the only compute path)"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for pkg in ("lyssandra_b200", "lyssa"):
        for dirpath, _, files in os.walk(os.path.join(root, pkg)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h")):
                    continue
                text = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "lyssa_oracle" in text:
                    bad.append((f, "imports the oracle"))
                if re.search(r"open\(.*/root/reference|sys\.path.*reference", text):
                    bad.append((f, "reads the reference tree"))
    assert not bad, bad


def test_thresholding_coders_without_gpu_fail_loudly():
    from lyssandra_b200 import engine
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    for alg, params in (("thresh", {"n_nonzero_coefs": 2}), ("iht", {"n_nonzero_coefs": 2, "eta": 0.1, "n_iter": 1})):
        with pytest.raises(RuntimeError):
            sparse_encoder(alg, params, verbose=False).encode(np.zeros((4, 3), dtype=np.float32), np.eye(4, dtype=np.float32))
