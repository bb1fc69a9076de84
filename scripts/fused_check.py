"""Bring-up check of the fused tcgen05 Batch-OMP kernel: parity vs the C oracle on a few shapes,
then timing at the bench size.  Each shape runs in its own process under a timeout so that a
hang in one configuration cannot eat the GPU call."""
import argparse
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_one(n, K, N, k, time_it):
    import numpy as np
    import torch
    import parity
    from lyssandra_b200 import _native
    from oracle import c_oracle as co
    from oracle import lyssa_oracle as lo

    lib = _native.load()
    dev = torch.device("cuda", 0)
    Xh = lo.synthetic_patches(N, n, seed=0)
    Dh = lo.synthetic_dictionary(K, n, seed=1)
    X = torch.from_numpy(np.ascontiguousarray(Xh.T)).to(dev).t()
    D = torch.from_numpy(Dh).to(dev)
    idx = torch.full((N, k), -7, dtype=torch.int32, device=dev); val = torch.full((N, k), 7.0, device=dev)
    nsel = torch.full((N,), -7, dtype=torch.int32, device=dev)
    Zt = torch.full((N, K), 7.0, device=dev)
    G = torch.empty((K, K), device=dev)
    wsb = lib.lys_bomp_workspace_bytes(n, K, N, k)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def encode():
        _native.check(lib.lys_gram(D.data_ptr(), K, n, K, G.data_ptr(), st))
        _native.check(lib.lys_bomp_encode(X.data_ptr(), X.stride(0), X.stride(1), D.data_ptr(), K, G.data_ptr(), n, K, N, k,
                                          idx.data_ptr(), val.data_ptr(), nsel.data_ptr(), Zt.data_ptr(), 1, K,
                                          ws.data_ptr(), wsb, st))
    encode()
    torch.cuda.synchronize()
    sub = min(N, 20000)
    idx_r, val_r, nsel_r, gap, vs = co.batch_omp_sparse(Xh[:, :sub].astype(np.float64), Dh.astype(np.float64), k, trace=True)
    ok = parity.comparable_columns(gap, vs, nsel_r, k)
    ig, vg = parity.sorted_codes(idx[:sub].cpu().numpy(), val[:sub].cpu().numpy())
    ir, vr = parity.sorted_codes(idx_r, val_r)
    same = np.all(ig == ir, axis=1)
    both = ok & same
    err = float(np.max(np.abs(vg[both] - vr[both])) / np.max(np.abs(vr[both]))) if both.any() else float("nan")
    Zd = Zt[:sub].cpu().numpy()
    dense_ok = bool(np.count_nonzero(Zd) == int((val[:sub] != 0).sum().item()) and not np.any(Zd == 7.0))
    print("n=%d K=%d N=%d k=%d: comparable %d/%d, support mismatches on comparable %d (all %d), coef rel-inf %.2e, "
          "nsel==k %d, dense_ok %s" % (n, K, N, k, int(ok.sum()), sub, int((ok & ~same).sum()), int((~same).sum()), err,
                                        int((nsel[:sub] == k).sum().item()), dense_ok), flush=True)
    bad = np.flatnonzero(ok & ~same)[:3]
    for b in bad:
        print("   col %d gpu=%s ref=%s" % (b, ig[b], ir[b]))
    if time_it:
        for _ in range(2):
            encode()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            encode()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print("   timing: %.3f ms per encode of %d signals = %.3e patches/s" % (ms, N, N / ms * 1e3), flush=True)


if __name__ == "__main__":
    import threading
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="small")
    ap.add_argument("--limit", type=float, default=90.0, help="seconds per shape before the watchdog kills the process")
    a = ap.parse_args()
    sets = {
        "single": [(64, 256, 1000, 5, 0), (64, 512, 5000, 5, 0), (48, 512, 3000, 4, 0), (64, 512, 1 << 20, 5, 1)],
        "pair": [(64, 1024, 5000, 5, 0), (64, 768, 3000, 3, 0), (64, 1024, 4099, 10, 0), (64, 1024, 1 << 20, 5, 1),
                 (64, 1024, 1 << 20, 10, 1)],
    }
    import torch
    torch.cuda.init()
    print("torch ready", flush=True)
    state = {"t": time.time(), "what": "import"}

    def watchdog():
        while True:
            time.sleep(2)
            if time.time() - state["t"] > a.limit:
                print("WATCHDOG: %s exceeded %.0f s -> hang, exiting" % (state["what"], a.limit), flush=True)
                os._exit(3)
    threading.Thread(target=watchdog, daemon=True).start()
    for sh in sets[a.set]:
        state["t"] = time.time(); state["what"] = "n=%d K=%d N=%d k=%d" % sh[:4]
        t0 = time.time()
        run_one(*sh[:4], bool(sh[4]))
        print("   (%.0f s)" % (time.time() - t0), flush=True)
