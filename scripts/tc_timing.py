"""Bring-up: per-phase cycle breakdown of the fused kernel (LYS_TC_TIMING=1 must be set)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lyssandra_b200 import _native
from oracle import lyssa_oracle as lo
lib = _native.load()
dev = torch.device("cuda", 0)
n, K, k, N = 64, int(os.environ.get("TC_K", 1024)), int(os.environ.get("TC_k", 5)), 1 << 20
X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=0).T)).to(dev).t()
D = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=1)).to(dev)
idx = torch.empty((N, k), dtype=torch.int32, device=dev); val = torch.empty((N, k), device=dev)
nsel = torch.empty((N,), dtype=torch.int32, device=dev); Zt = torch.empty((N, K), device=dev)
G = torch.empty((K, K), device=dev)
wsb = lib.lys_bomp_workspace_bytes(n, K, N, k); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
_native.check(lib.lys_gram(D.data_ptr(), K, n, K, G.data_ptr(), st))
FLAGS = _native.BOMP_SCREEN if os.environ.get("TC_SCREEN") else 0
def enc(dense=True):
    _native.check(lib.lys_bomp_encode_ex(X.data_ptr(), X.stride(0), X.stride(1), D.data_ptr(), K, G.data_ptr(), n, K, N, k,
                                         idx.data_ptr(), val.data_ptr(), nsel.data_ptr(), Zt.data_ptr() if dense else None, 1, K, ws.data_ptr(), wsb, FLAGS, st))
out = (ctypes.c_ulonglong * 16)()
dbg = lib.lys_debug_tc_timing
trace = (ctypes.c_ulonglong * 8192)(); ntr = ctypes.c_uint()
for dense in (True, False):
    enc(dense); torch.cuda.synchronize(); dbg(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); enc(dense); e1.record(); torch.cuda.synchronize()
    dbg(out)
    lib.lys_debug_tc_trace(trace, ctypes.byref(ntr))
    with open(os.path.join(ROOT, "gpurun_out", "tc_trace_%s_%s.txt" % (os.environ.get("LYS_TC_SLOTS", "2"), "dense" if dense else "sparse")), "w") as fh:
        for i in range(min(int(ntr.value), 8192)):
            e = int(trace[i]); fh.write("%d %d %d\n" % (e >> 8, (e >> 4) & 15, e & 15))
    v = [int(x) for x in out]
    names = ["tile start", "zero fill", "wait acc_full", "scan", "finish", "outputs", "resolve", "update", "mma: wait a_ready", "mma: wait acc_empty", "mma: issue"]
    print("screen=%s K=%d k=%d dense=%s: %.3f ms" % (bool(FLAGS), K, k, dense, e0.elapsed_time(e1)))
    tot_e = sum(v[:8]); tot_m = sum(v[8:11])
    for i, nm in enumerate(names):
        if nm: print("   %-28s %6.2f %%  (%.0f cycles per warp-step)" % (nm, 100.0 * v[i] / (tot_e if i < 8 else tot_m), v[i] / (N / 32.0 * k)))
    if v[13]:
        print("   signal-steps %d, uncertified %d (%.2f %%), candidate pieces per uncertified %.2f" % (v[13], v[11], 100.0 * v[11] / v[13], v[12] / max(v[11], 1)))
