"""Bring-up: time dense and sparse encodes of the bench workload with the library named by LYSSA_B200_LIB."""
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
def enc(dense, flags):
    _native.check(lib.lys_bomp_encode_ex(X.data_ptr(), X.stride(0), X.stride(1), D.data_ptr(), K, G.data_ptr(), n, K, N, k,
                                         idx.data_ptr(), val.data_ptr(), nsel.data_ptr(), Zt.data_ptr() if dense else None, 1, K, ws.data_ptr(), wsb, flags, st))
for flags, name in ((0, "split3"), (_native.BOMP_SCREEN, "screen")):
    res = []
    for dense in (True, False):
        for _ in range(3): enc(dense, flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): enc(dense, flags)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    chk = int(idx.sum().item())
    print("%s %s K=%d k=%d: dense %.3f ms, sparse %.3f ms (idx checksum %d)" % (os.path.basename(os.environ.get("LYSSA_B200_LIB", "default")), name, K, k, res[0], res[1], chk), flush=True)
