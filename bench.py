#!/usr/bin/env python
"""bench.py — headline benchmark of the Batch-OMP encode path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): Batch-OMP encode of 1 M synthetic 8x8 grayscale patches
(uniform pixels, mean removed), D 64x1024 unit-norm, k = 5, fp32, dense Z (K x N) written —
per GPU (weak scaling: every rank encodes its own 1 M patches, no data-path collective).
A "step" is one full encode of that batch: Gram + correlations + greedy + dense-Z write.

Own arm: inputs resident in HBM, CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  `e2e` times the same step through the host-buffer C-ABI call
(lys_bomp_encode_host) with pinned host X in and pinned dense Z out, copies inside the timed
region.  `roofline` is measured live: CUDA events around every launch of the dominant kernel
(lys_profile_*), algorithmic bytes = (4n + 4K) per patch.  `cpu_baseline` is the
NumPy oracle (bit-identical restatement of the reference) run the reference's way
(sparse_encoder('bomp', n_jobs=cores): process pool over 100 column batches, 1 BLAS thread per
worker) on a bounded contiguous sample, rank 0 at N=1 only, in a subprocess.

Reference arm (--impl reference): that same CPU path, all host cores, one bounded sample per
step; under torchrun only rank 0 works.  It never touches CUDA.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_PER_GPU = 1 << 20
N_FEATURES, N_ATOMS, K_NONZERO = 64, 1024, 5
METRIC, UNIT = "batch_omp_patches_per_sec", "patches/s"
WORKLOAD = "Batch-OMP encode, 1M synthetic 8x8 grayscale patches per GPU, D 64x1024, k=5, fp32, dense Z written"


def config(n_gpus):
    return {"workload": WORKLOAD, "n_features": N_FEATURES, "n_atoms": N_ATOMS, "n_nonzero_coefs": K_NONZERO,
            "patches_per_gpu": N_PER_GPU, "global_patches": N_PER_GPU * n_gpus,
            "parallelism": "patch-sharded x%d, no collective" % n_gpus,
            "l2_policy": "inputs larger than L2 (X 256 MiB, Z 4 GiB per GPU per step; L2 126 MB)"}


# ------------------------------------------------------------------------- reference arm
def cpu_encode_rate(sample, cores, repeats=1):
    """patches/s of the oracle's reference-regime encoder on `sample` columns."""
    from oracle import lyssa_oracle as lo
    X = lo.synthetic_patches(sample, N_FEATURES, seed=0).astype(np.float64)
    D = lo.synthetic_dictionary(N_ATOMS, N_FEATURES, seed=1).astype(np.float64)
    enc = lo.sparse_encoder("bomp", {"n_nonzero_coefs": K_NONZERO}, n_jobs=cores, verbose=False)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        Z = enc.encode(X, D)
        best = min(best, time.perf_counter() - t0)
    assert Z.shape == (N_ATOMS, sample)
    return sample / best, best


def pick_sample(cores, target_s):
    """Calibrate on 1024 columns in-process (1 core), size the sample for ~target_s seconds."""
    from oracle import lyssa_oracle as lo
    X = lo.synthetic_patches(1024, N_FEATURES, seed=0).astype(np.float64)
    D = lo.synthetic_dictionary(N_ATOMS, N_FEATURES, seed=1).astype(np.float64)
    t0 = time.perf_counter()
    lo.sparse_encoder("bomp", {"n_nonzero_coefs": K_NONZERO}, n_jobs=1, verbose=False).encode(X, D)
    rate1 = 1024 / (time.perf_counter() - t0)
    est = rate1 * max(1.0, 0.6 * cores)
    sample = int(est * target_s)
    step = 100 * max(cores, 1)
    return max(step, min(262144, sample // step * step))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    sample = args.sample or pick_sample(cores, 5.0)
    for _ in range(args.warmup if args.sample is None else 0):
        cpu_encode_rate(min(sample, 100 * cores * 4), cores)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_encode_rate(sample, cores)
        times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    desc = ("first %d of the 1M patches per step (contiguous prefix), oracle = NumPy restatement pinned "
            "bit-exact to the reference, run as sparse_encoder('bomp', n_jobs=%d)" % (sample, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                mhz, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax.append(mx)
            if t0 <= t <= t1:
                sm.append(mhz)
                for name, flag in zip(names, parts[5:9]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(p.split(",")[1]) for _, p in self.rows[-3:] if len(p.split(",")) > 2] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax) if smax else 0.0,
                "reasons": sorted(reasons), "samples": len(sm)}


def tensor_info(peaks, n, K, k, patches, kernel_ms):
    """fp16 tensor-core work of the fused kernel against the measured dense bf16/fp16 peak: EXECUTED flops (k greedy steps
    x 3 split products x 2*64*K per patch: every step re-correlates the residual) and ALGORITHMIC flops (one correlation
    D^T x per patch, 2*n*K) — the first says how busy the pipe is, only the second is credit."""
    executed = 2.0 * 64 * K * 3 * k * patches
    algorithmic = 2.0 * n * K * patches
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    sec = kernel_ms / 1e3 if kernel_ms > 0 else float("inf")
    return {"executed_tflops": executed / 1e12 / sec, "algorithmic_tflops": algorithmic / 1e12 / sec, "peak_tflops": peak_tf,
            "frac_executed": executed / 1e12 / sec / peak_tf if peak_tf else None,
            "frac_algorithmic": algorithmic / 1e12 / sec / peak_tf if peak_tf else None,
            "executed_flop_per_patch": 2.0 * 64 * K * 3 * k, "algorithmic_flop_per_patch": 2.0 * n * K,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in peaks else "fallback 1590 TFLOP/s"}


def parity_report(idx_gpu, val_gpu, n_cols):
    """cpu_baseline leg, checker role: the float64 C oracle (oracle/bomp_oracle.c) on the first n_cols patches of the
    SAME synthetic workload against the codes the timed GPU steps produced.  Supports must be identical on every column
    that is not a near-tie in the oracle's own float64 arithmetic (tests/parity.py policy); returns the counts."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    from oracle import c_oracle as co
    from oracle import lyssa_oracle as lo
    X = lo.synthetic_patches(n_cols, N_FEATURES, seed=0).astype(np.float64)
    D = lo.synthetic_dictionary(N_ATOMS, N_FEATURES, seed=1).astype(np.float64)
    idx_r, val_r, nsel, gap, vs = co.batch_omp_sparse(X, D, K_NONZERO, trace=True)
    ok = parity.comparable_columns(gap, vs, nsel, K_NONZERO)
    ig, vg = parity.sorted_codes(idx_gpu[:n_cols], val_gpu[:n_cols])
    ir, vr = parity.sorted_codes(idx_r, val_r)
    same = np.all(ig == ir, axis=1)
    both = ok & same
    scale = float(np.max(np.abs(vr[both]))) if both.any() else 1.0
    return {"columns": int(n_cols), "compared": int(ok.sum()), "support_mismatch_on_compared": int((ok & ~same).sum()),
            "excluded_near_ties": int((~ok).sum()), "mismatch_in_excluded": int((~ok & ~same).sum()),
            "coef_rel_inf": float(np.max(np.abs(vg[both] - vr[both])) / scale) if both.any() else None,
            "oracle": "float64 C restatement of batch_omp, pinned to the live reference (tests/test_oracle.py)",
            "policy": "near-tie = oracle top1/top2 gap < 1e-5 at any step, pivot < 1e-6, or fewer than k atoms"}


# ------------------------------------------------------------------- secondary metric
def ksvd_iteration_ms(dev, rank, world, iters=3):
    """K-SVD iteration time at BASELINE cfg3 (2M 8x8 patches in total, K=1024, k=10), patch-sharded
    over the ranks: Batch-OMP encode -> residual -> users-of-atom CSR -> sweep (per-atom all-reduce
    inside the kernel over peer-mapped mailboxes when world > 1) -> error (+ scalar all-reduce).
    Every timed iteration starts from the same D (a rate, not cfg3's 20-iteration trajectory).
    Returns (median ms per iteration on this rank, stage split, parity dict or None)."""
    import torch
    import torch.distributed as dist
    from lyssandra_b200 import engine
    from lyssandra_b200.distributed import DistContext, PeerExchange
    from lyssandra_b200.sparse_coding import sparse_encoder
    from oracle import lyssa_oracle as lo
    n, K, k, N = 64, 1024, 10, 2000000 // world
    X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=2000 + rank).T)).to(dev).t()
    D0 = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=1)).to(dev)
    ctx = DistContext.from_env_or_group()
    ex = PeerExchange(ctx)
    enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)

    def iteration(Xs, D, comm, timed=True):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record(); codes = enc.encode_sparse(Xs, D)
        ev[1].record(); R, _ = engine.residual(Xs, D, codes, want_residual=True, want_error=False)
        ev[2].record(); rowptr, entries = engine.build_atom_csr(codes)
        ev[3].record(); engine.approx_ksvd_sweep(R, D, codes, rowptr, entries, n_cycles=1, comm=comm)
        ev[4].record(); err = engine.frobenius2(R)     # R is kept current by the sweep
        if comm is not None:
            ctx.allreduce_sum_(err)
        ev[5].record(); torch.cuda.synchronize(dev)
        return ev, err

    totals, stages = [], []
    for it in range(iters + 1):
        D = D0.clone()
        torch.cuda.synchronize(dev); ctx.barrier()
        ev, _ = iteration(X, D, ex.handle)
        if it > 0:
            totals.append(ev[0].elapsed_time(ev[5]))
            stages.append([ev[i].elapsed_time(ev[i + 1]) for i in range(5)])

    _progress("ksvd_iteration: timed iterations done")
    # ---- parity of the sharded sweep against ONE GPU on the same signals (a 262144-patch slice: the first
    # 262144/world patches of every shard, gathered on rank 0 in rank order)
    parity = None
    if world > 1:
        m = 262144 // world
        Xs = X[:, :m]
        Dsh = D0.clone()
        torch.cuda.synchronize(dev); ctx.barrier()
        _, err_sh = iteration(Xs, Dsh, ex.handle)
        _progress("ksvd_iteration: sharded sweep of the parity slice done")
        mine = Xs.t().contiguous()                                      # (m, n)
        allX = torch.empty((world * m, n), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allX, mine)
        allD = torch.empty((world,) + tuple(Dsh.shape), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allD, Dsh.contiguous())
        if rank == 0:
            D1 = D0.clone()
            _, err_1 = iteration(allX.t(), D1, None)
            parity = {"signals": world * m,
                      "sweep_max_dD_vs_1gpu": float((D1 - Dsh).abs().max()),
                      "D_identical_across_ranks": bool(all(torch.equal(allD[r], allD[0]) for r in range(world))),
                      "objective_rel": abs(float(err_sh.item()) - float(err_1.item())) / float(err_1.item())}
        torch.cuda.synchronize(dev); ctx.barrier()
    ex.close()
    med = [float(np.median([s[i] for s in stages])) for i in range(5)]
    return float(np.median(totals)), dict(zip(["encode", "residual", "csr", "sweep", "error"], med)), parity


def _scspm_block(b, chunk, size):
    """synthetic smooth-noise images of block b (any rank can regenerate any block)"""
    rng = np.random.default_rng(5000 + b)
    base = rng.random((chunk, size + 8, size + 8), dtype=np.float32)
    return (base[:, :-8, :-8] + base[:, 4:-4, 4:-4] + base[:, 8:, 8:]) * (255.0 / 3.0)


def scspm_images_per_s(dev, rank, world, n_total=10000, size=256, chunk=250):
    """ScSPM pipeline of BASELINE cfg5: 10k synthetic 256x256 images, sharded by image over the ranks (no collective)
    -> dense SIFT (16x16 patches, grid step 6: 41x41 descriptors per image) -> Batch-OMP (D 128x1024, k=5) -> 3-level
    max-|z| pooling + l2 normalisation.  Images are resident on the device; blocks of `chunk` images per call.
    Returns (ms for this rank's shard, parity dict or None)."""
    import torch
    import torch.distributed as dist
    from lyssandra_b200.sparse_coding import sparse_encoder
    from lyssandra_b200.feature_extract import sc_spm_extractor, dsift_extractor, sc_max_pooling, l2_normalizer
    from lyssandra_b200.utils import shard_bounds
    from oracle import lyssa_oracle as lo
    n_blocks = n_total // chunk
    b_lo, b_hi = shard_bounds(n_blocks, world, rank)
    blocks = [[torch.from_numpy(np.ascontiguousarray(im)).to(dev) for im in _scspm_block(b, chunk, size)] for b in range(b_lo, b_hi)]
    D = torch.from_numpy(lo.synthetic_dictionary(1024, 128, seed=9)).to(dev)
    enc = sparse_encoder("bomp", {"n_nonzero_coefs": 5}, verbose=False)
    ex = sc_spm_extractor(feature_extractor=dsift_extractor(step_size=6, patch_size=16), levels=(1, 2, 4), sparse_coder=enc,
                          pooling_operator=sc_max_pooling(), normalizer=l2_normalizer())
    _progress("scspm: images resident")
    F = [ex.encode(blocks[0], D)]
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    F = [ex.encode(imgs, D) for imgs in blocks]
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    _progress("scspm: shard done")
    assert tuple(F[0].shape) == (21 * 1024, chunk) and all(bool(torch.isfinite(f).all()) for f in F)
    parity = None
    if world > 1:
        # the first block of the LAST rank's shard, recomputed by rank 0 alone from the same seed
        last_lo, _ = shard_bounds(n_blocks, world, world - 1)
        ref = F[0].contiguous() if rank == world - 1 else torch.empty((21 * 1024, chunk), dtype=torch.float32, device=dev)
        dist.broadcast(ref, src=world - 1)
        if rank == 0:
            mine = ex.encode([torch.from_numpy(np.ascontiguousarray(im)).to(dev) for im in _scspm_block(last_lo, chunk, size)], D)
            parity = {"images": chunk, "max_dF_vs_1gpu": float((mine - ref).abs().max()), "identical": bool(torch.equal(mine, ref))}
    return ms, parity


def odl_minibatch_ms(dev, rank, world, n=128, K=2048, k=5, b=4096, n_mb=24):
    """ODL at BASELINE cfg4 (SIFT-like 128-d descriptors, D 128x2048, minibatch 4096, beta = 0.9): the minibatch is split
    over the ranks, each rank encodes its b/world signals, the slices are all-gathered as (signal, idx, val) and the
    statistics + dictionary update run replicated in a fixed order (bit-identical D on every rank).  Returns
    (median ms per minibatch through online_dict_learn's own loop, stage split on one minibatch, parity dict or None)."""
    import torch
    import torch.distributed as dist
    from lyssandra_b200 import engine
    from lyssandra_b200.distributed import DistContext
    from lyssandra_b200.dict_learning import online_dict_learn
    from lyssandra_b200.sparse_coding import sparse_encoder
    from oracle import lyssa_oracle as lo
    ctx = DistContext.from_env_or_group()
    b_loc = b // world
    Xfull = np.ascontiguousarray(lo.synthetic_descriptors(b * n_mb, n, seed=0).T)              # (b*n_mb, n), minibatch-major
    mine = np.concatenate([Xfull[m * b + rank * b_loc: m * b + (rank + 1) * b_loc] for m in range(n_mb)], axis=0)
    X = torch.from_numpy(mine).to(dev).t()                                                  # this rank's slice of every minibatch
    rng = np.random.default_rng(1)
    D0 = torch.from_numpy(np.ascontiguousarray(lo.norm_cols(np.abs(rng.standard_normal((n, K)))).astype(np.float32))).to(dev)
    enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)

    def run(Xs, bs, dctx):
        D = D0.clone()
        torch.cuda.synchronize(dev)
        if dctx is not None:              # the single-GPU reference run is rank 0's alone: no collective in it
            dctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        D, A, B = online_dict_learn(Xs, K, sparse_coder=enc, batch_size=bs, D_init=D, beta=0.9, n_epochs=1, dist=dctx)
        e1.record(); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n_mb, D, A

    _progress("odl: data resident")
    run(X, b_loc, ctx if world > 1 else None)                                               # warm-up
    _progress("odl: warm-up epoch done")
    ms, D, A = run(X, b_loc, ctx if world > 1 else None)
    ms2, D_again, _ = run(X, b_loc, ctx if world > 1 else None)
    reproducible = bool(torch.equal(D, D_again))
    _progress("odl: timed epochs done")
    # stage split of one minibatch (this rank's slice encoded, whole minibatch accumulated)
    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev); e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1), out
    Xb = torch.from_numpy(Xfull[:b]).to(dev).t()
    Ds = D0.clone(); As = torch.zeros((K, K), device=dev); Bs = torch.zeros((n, K), device=dev)
    st = {"encode_slice": [], "accumulate": [], "update": []}
    for it in range(6):
        t1, _ = timed(lambda: enc.encode_sparse(Xb[:, :b_loc], Ds))
        codes = enc.encode_sparse(Xb, Ds)
        t2, _ = timed(lambda: engine.odl_accumulate_(Xb, codes, 0.9, As, Bs))
        t3, _ = timed(lambda: engine.odl_update_dict_(Ds, As, Bs))
        if it > 0:
            st["encode_slice"].append(t1); st["accumulate"].append(t2); st["update"].append(t3)
    stages = {key: float(np.median(v)) for key, v in st.items()}
    parity = {"bitwise_reproducible_across_runs": reproducible}
    _progress("odl: stage split done")
    if world > 1:
        allD = torch.empty((world,) + tuple(D.shape), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(allD, D.contiguous())
        if rank == 0:
            _, D1, A1 = run(torch.from_numpy(Xfull).to(dev).t(), b, None)                   # ONE GPU, whole minibatches, same order
            parity.update({"minibatches": n_mb, "max_dD_vs_1gpu": float((D1 - D).abs().max()),
                           "bitwise_equal_to_1gpu": bool(torch.equal(D1, D) and torch.equal(A1, A)),
                           "D_identical_across_ranks": bool(all(torch.equal(allD[r], allD[0]) for r in range(world)))})
        torch.cuda.synchronize(dev); ctx.barrier()
    return min(ms, ms2), stages, parity


def sibling_coders_ms(dev, n=64, K=1024, N=1 << 20, k=5, reps=3):
    """'thresh' and 'iht'  on the headline workload's shape: ms per 1M signals, codes stay sparse."""
    import torch
    from lyssandra_b200 import engine
    from oracle import lyssa_oracle as lo
    X = engine.as_device_matrix(torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=21))).to(dev), dev)
    D = engine.as_dictionary(torch.from_numpy(lo.synthetic_dictionary(K, n, seed=22)).to(dev), dev)
    out = {}
    for name, fn in (("thresh", lambda: engine.thresh_encode(X, D, k)),
                     ("thresh_dense", lambda: engine.thresh_encode(X, D, k, dense=True)),
                     ("iht_4_iterations", lambda: engine.iht_encode(X, D, k, 0.2, 4))):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        out[name] = e0.elapsed_time(e1) / reps
    return out


_T0 = time.perf_counter()


def _progress(msg):
    """phase marks on stderr (every rank): where a slow or hung run spent its time"""
    sys.stderr.write("[bench rank %s +%.1fs] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter() - _T0, msg))
    sys.stderr.flush()


# ----------------------------------------------------------------------------- own arm
def run_own(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    # CPU baseline first, in a subprocess (a fork()ing process pool must not share a CUDA context)
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                  "--warmup", "0", "--sample", str(pick_sample(os.cpu_count() or 1, 12.0))],
                                 capture_output=True, text=True, timeout=600, cwd=ROOT)
            cpu_base = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as exc:  # reported, never silently dropped
            cpu_base = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (exc,)}

    import torch
    import torch.distributed as dist
    from lyssandra_b200 import _native, engine
    from lyssandra_b200.sparse_coding import sparse_encoder
    from oracle import lyssa_oracle as lo      # synthetic data generators only (bench input, not a checker call)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.load()
    _native.check(lib.lys_device_info(local, None, None, None))
    _progress("library loaded, process group up")

    n, K, k, N = N_FEATURES, N_ATOMS, K_NONZERO, N_PER_GPU
    Xh_sm = np.ascontiguousarray(lo.synthetic_patches(N, n, seed=0 if rank == 0 else 1000 + rank).T)   # (N, n) signal-major; seed 1 is D's
    Dh = lo.synthetic_dictionary(K, n, seed=1)
    X = torch.from_numpy(Xh_sm).to(dev).t()                                        # logical (n, N), resident in HBM
    D = torch.from_numpy(Dh).to(dev)
    enc = sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
    idx = torch.empty((N, k), dtype=torch.int32, device=dev)
    val = torch.empty((N, k), dtype=torch.float32, device=dev)
    nsel = torch.empty((N,), dtype=torch.int32, device=dev)
    Zt = torch.empty((N, K), dtype=torch.float32, device=dev)
    G = torch.empty((K, K), dtype=torch.float32, device=dev)
    wsb = lib.lys_bomp_workspace_bytes(n, K, N, k)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def step():
        _native.check(lib.lys_gram(D.data_ptr(), K, n, K, G.data_ptr(), stream))
        _native.check(lib.lys_bomp_encode(X.data_ptr(), X.stride(0), X.stride(1), D.data_ptr(), K, G.data_ptr(),
                                          n, K, N, k, idx.data_ptr(), val.data_ptr(), nsel.data_ptr(),
                                          Zt.data_ptr(), 1, K, ws.data_ptr(), wsb, stream))

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    _progress("inputs resident")
    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    _progress("warm-up done")
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    lib.lys_profile_fetch(None, None, None, 1)
    lib.lys_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_begin = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    t_end = time.perf_counter()
    lib.lys_profile_enable(0)
    ms = e0.elapsed_time(e1)
    kms, kl, kname = ctypes.c_double(), ctypes.c_int64(), ctypes.c_char_p()
    _native.check(lib.lys_profile_fetch(ctypes.byref(kms), ctypes.byref(kl), ctypes.byref(kname), 1))
    # second timed region under load for the clocks record if the first was too short to sample
    if t_end - t_begin < 0.5:
        t_begin2 = time.perf_counter()
        while time.perf_counter() - t_begin2 < 0.6:
            step()
        torch.cuda.synchronize(dev)
        t_begin, t_end = t_begin2, time.perf_counter()
    sampler.stop()
    clocks = sampler.summary(t_begin, t_end)

    # sanity on the result the timed steps produced (cheap, after timing)
    assert int((nsel == k).sum()) >= N - 8 and int((idx < 0).sum()) <= 8 * k, "encode produced truncated supports on non-degenerate data"
    assert int((Zt[:4096] != 0).sum()) == 4096 * k
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            parity = parity_report(idx[:32768].cpu().numpy(), val[:32768].cpu().numpy(), 32768)
        except Exception as exc:
            parity = {"failed": repr(exc)}

    _progress("timed region done: %.3f ms per step" % (ms / args.steps))
    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    Xpin = torch.from_numpy(Xh_sm).pin_memory()
    Zpin = torch.empty((N, K), dtype=torch.float32).pin_memory()
    Dpin = torch.from_numpy(Dh).pin_memory()

    def e2e_step():
        _native.check(lib.lys_bomp_encode_host(Xpin.data_ptr(), 1, n, Dpin.data_ptr(), K, n, K, N, k,
                                               None, None, None, Zpin.data_ptr(), 1, K, local))

    e2e_steps = max(2, min(args.steps, 5))
    _progress("e2e buffers pinned")
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    assert int((Zpin[:4096] != 0).sum()) == 4096 * k
    # sparse-output variant (what the learners consume): no dense Z over PCIe
    ipin = torch.empty((N, k), dtype=torch.int32).pin_memory()
    vpin = torch.empty((N, k), dtype=torch.float32).pin_memory()
    spin = torch.empty((N,), dtype=torch.int32).pin_memory()
    _native.check(lib.lys_bomp_encode_host(Xpin.data_ptr(), 1, n, Dpin.data_ptr(), K, n, K, N, k,
                                           ipin.data_ptr(), vpin.data_ptr(), spin.data_ptr(), None, 1, K, local))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _native.check(lib.lys_bomp_encode_host(Xpin.data_ptr(), 1, n, Dpin.data_ptr(), K, n, K, N, k,
                                               ipin.data_ptr(), vpin.data_ptr(), spin.data_ptr(), None, 1, K, local))
    e2e_sparse_s = time.perf_counter() - t0

    _progress("e2e done: %.1f ms per step" % (e2e_s * 1e3 / e2e_steps))
    del Xpin, Zpin, ipin, vpin, spin, Zt
    torch.cuda.empty_cache()
    # headline numbers: max over ranks, BEFORE the secondary measurements (a stuck extra must not cost the headline line)
    t = torch.tensor([ms, e2e_s * 1e3 / e2e_steps, e2e_sparse_s * 1e3 / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms, e2e_sparse_ms = [float(v) for v in t.tolist()]

    # ---- secondary measurements (extras), each finalised (max over ranks) as soon as it is done, under a watchdog:
    # if they exceed --extras-budget seconds rank 0 prints the line with what is finished and every rank exits 0
    finished = {}                      # name -> (ms max over ranks, payload of rank 0)
    state = {"line": None, "done": False}
    lock = threading.Lock()

    def emit(aborted=None):
        with lock:
            if state["done"]:
                return
            state["done"] = True
            if rank == 0 and state["line"] is not None:
                line = state["line"](finished, aborted)
                sys.stdout.write(json.dumps(line) + "\n")
                sys.stdout.flush()

    def watchdog():
        deadline = time.perf_counter() + float(args.extras_budget)
        while time.perf_counter() < deadline:
            time.sleep(0.5)
            if state["done"]:
                return
        _progress("extras exceeded %.0f s: printing the line without the unfinished ones and exiting" % args.extras_budget)
        emit(aborted="extras exceeded the %.0f s budget; finished: %s" % (args.extras_budget, sorted(finished)))
        os._exit(0)

    def max_over_ranks(v):
        tt = torch.tensor([v if v is not None else -1.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def run_extra(name, fn, all_ranks=True):
        if args.no_extras or (not all_ranks and rank != 0):
            return
        payload, note, v = None, None, None
        try:
            v, payload = fn()
        except Exception as exc:          # secondary metric: reported, never fatal for the headline line
            note = "failed: %r" % (exc,)
        vmax = max_over_ranks(v) if all_ranks else (v if v is not None else -1.0)
        finished[name] = (vmax, payload, note)
        _progress("extras: %s done" % name)

    def _ksvd():
        v, stages, par = ksvd_iteration_ms(dev, rank, world)
        return v, (stages, par)

    def _spm():
        v, par = scspm_images_per_s(dev, rank, world)
        return v, par

    def _odl():
        v, stages, par = odl_minibatch_ms(dev, rank, world)
        return v, (stages, par)

    def _sib():
        return 0.0, sibling_coders_ms(dev)


    def build_line(fin, aborted):
        ksvd_ms_max, kp, ksvd_note = fin.get("ksvd_iteration", (-1.0, None, "not run"))
        ksvd_stages, ksvd_parity = kp if kp else (None, None)
        spm_ms_max, spm_parity, spm_note = fin.get("scspm_pipeline", (-1.0, None, "not run"))
        odl_ms_max, op, odl_note = fin.get("odl_minibatch", (-1.0, None, "not run"))
        odl_stages, odl_parity = op if op else (None, None)
        _, sib_ms, sib_note = fin.get("sibling_coders", (-1.0, None, "not run"))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bytes_per_patch = 4 * n + 4 * K
        launches_k = int(kl.value)
        achieved = (bytes_per_patch * N * args.steps / 1e9) / (kms.value / 1e3) if kms.value > 0 else 0.0
        total_patches = N * world * args.steps
        line = {
            "metric": METRIC, "value": total_patches / (ms_max / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config(world),
            "clocks": clocks,
            "e2e": {"value": N * world / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": (N * n * 4 + n * K * 4) * world, "d2h_bytes_per_step": N * k * 8 * world,
                    "api": "lys_bomp_encode_host: pinned host X -> pinned host dense Z (what sparse_encoder.encode(numpy) calls)",
                    "dense_z_bytes_written_in_host_memory_per_step": N * K * 4 * world,
                    "host_dense_gb_per_s_per_rank": N * K * 4 / 1e9 / (e2e_ms / 1e3),
                    "how": "the device returns the sparse codes (idx, val: 8k bytes per patch over PCIe); the rows of the dense (K x N) matrix the reference's contract returns (99.5 % zeros) are written into the caller's host buffer by threads of the library (zero-fill + scatter, non-temporal stores), overlapped with the device encoding the next chunks.  Round 1 sent the dense matrix over PCIe (4.29 GB per step, 53 GB/s: 1.28e7 patches/s)",
                    "bound": "host memory write bandwidth of the library's threads (all ranks of one box share the host's memory controllers, so the aggregate does not scale with N)",
                    "sparse_output_variant": {"value": N * world / (e2e_sparse_ms / 1e3), "unit": UNIT,
                                              "d2h_bytes_per_step": N * (8 * k + 4) * world,
                                              "note": "same call returning (idx,val,nsel) instead of dense Z"}},
            "gpu_launches": (lib.lys_bomp_launch_count(n, K, N, k) + 1) * args.steps,
            "roofline": {"bound": "hbm", "kernel": (kname.value or b"").decode(), "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": None,
                         "launches_timed": launches_k, "kernel_ms_per_step": kms.value / args.steps,
                         "kernel_share_of_step": kms.value / ms if ms > 0 else None,
                         "algorithmic_bytes_per_patch": bytes_per_patch, "peak_source": peak_src,
                         "step_level": {"achieved": bytes_per_patch * N * args.steps / 1e9 / (ms / 1e3),
                                        "frac": bytes_per_patch * N * args.steps / 1e9 / (ms / 1e3) / peak},
                         # the fused kernel's own ceiling is the tensor pipe: every greedy step is one
                         # fp32-faithful correlation GEMM (3 fp16 MMAs per k-step)
                         "tensor": tensor_info(peaks, n, K, k, N * args.steps, kms.value)},
            "parity": parity,
            "cpu_baseline": cpu_base,
            "extras": {"ksvd_iteration": {"workload": "approx K-SVD iteration, 2M 8x8 patches total (patch-sharded x%d), K=1024, k=10, n_cycles=1" % world,
                                          "ms_per_iter": ksvd_ms_max if ksvd_ms_max >= 0 else None, "stages_ms_rank0": ksvd_stages,
                                          "collective": "none" if world == 1 else "per-atom all-reduce of 2n+3 fixed-point words inside the sweep kernel over peer-mapped NVLink mailboxes; scalar NCCL all-reduce of the error",
                                          "timing": "median of 3 iterations, each from the same initial D (a per-iteration rate, not cfg3's 20-iteration trajectory)",
                                          "parity_vs_1gpu": ksvd_parity,
                                          "note": ksvd_note},
                       "scspm_pipeline": {"workload": "BASELINE cfg5: 10000 synthetic 256x256 images sharded by image over %d GPU(s), no collective -> dense SIFT 16x16 / step 6 (1681 descriptors per image) -> Batch-OMP D 128x1024 k=5 -> 3-level max-|z| pooling + l2, images resident on the device, 250 images per call" % world,
                                          "images_per_s": (10000.0 / (spm_ms_max / 1e3)) if spm_ms_max > 0 else None, "ms_total_max_over_ranks": spm_ms_max if spm_ms_max > 0 else None,
                                          "parity_vs_1gpu": spm_parity, "note": spm_note},
                       "odl_minibatch": {"workload": "BASELINE cfg4: ODL, minibatch 4096 SIFT-like 128-d descriptors split over %d GPU(s), D 128x2048, k=5, beta=0.9, 24 minibatches through online_dict_learn: encode slice -> all-gather (signal, idx, val) -> A,B statistics (fixed order) -> dictionary update (replicated)" % world,
                                         "ms_per_minibatch": odl_ms_max if odl_ms_max > 0 else None,
                                         "minibatches_per_s": (1e3 / odl_ms_max) if odl_ms_max > 0 else None,
                                         "stages_ms_rank0": odl_stages, "parity": odl_parity, "note": odl_note},
                       "sibling_coders": {"workload": "'thresh' / 'iht' coders, 1M synthetic patches (rank 0 only), D 64x1024, k=5, eta=0.2, sparse codes out unless noted",
                                          "ms_per_1M_signals": sib_ms, "note": sib_note}},
        }
        # DRAM traffic per launch cannot be read without a profiler: it is the dram__bytes_read.sum + dram__bytes_write.sum
        # of the committed `ncu --set full` capture of this kernel, and says so
        traffic_file = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        if os.path.isfile(traffic_file):
            try:
                tj = json.load(open(traffic_file))
                line["roofline"]["traffic"] = tj.get(line["roofline"]["kernel"])
                line["roofline"]["traffic_source"] = tj.get("_source", "ncu capture under profiles/ (not measured in this run)")
            except Exception:
                pass
        if aborted:
            line["extras"]["aborted"] = aborted
        return line

    state["line"] = build_line
    wd = threading.Thread(target=watchdog, daemon=True)
    if not args.no_extras:
        wd.start()
    run_extra("ksvd_iteration", _ksvd)
    run_extra("scspm_pipeline", _spm)
    run_extra("odl_minibatch", _odl)
    run_extra("sibling_coders", _sib, all_ranks=False)
    emit()
    _progress("line printed")
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        # The result is out; nothing after this point may keep the launch alive.  dist.barrier() with an eagerly
        # initialised NCCL communicator returns on the host before the peers have arrived, so a rank that had no
        # rank-0-only extra to run used to tear NCCL down while rank 0 was still measuring, and rank 0 then hung in its
        # own barrier (seen on the 2- and 4-GPU boxes).  The ranks now meet at an all-reduce whose result the host reads,
        # bounded by a timer, and leave without the NCCL teardown.
        threading.Timer(30.0, lambda: os._exit(0)).start()
        try:
            fin = torch.ones(1, device=dev)
            dist.all_reduce(fin)
            _progress("final rendezvous: %d of %d ranks" % (int(fin.item()), world))
        except Exception as exc:
            _progress("final rendezvous failed: %r" % (exc,))
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--sample", type=int, default=None, help="(reference arm) columns per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary K-SVD iteration timing")
    ap.add_argument("--extras-budget", type=float, default=240.0, help="seconds the secondary measurements may take before the line is printed without the unfinished ones")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
