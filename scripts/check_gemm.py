"""Bring-up check of the tcgen05 correlation GEMM: error of each implementation against float64."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200 import _native
from oracle import lyssa_oracle as lo
lib = _native.load(); dev = torch.device("cuda", 0)
for (K, C) in ((256, 1000), (1024, 5000), (1024, 40000)):
    Xh = np.ascontiguousarray(lo.synthetic_patches(C, 64, seed=3)); Dh = lo.synthetic_dictionary(K, 64, seed=4)
    ref = Xh.astype(np.float64).T @ Dh.astype(np.float64)
    scale = np.linalg.norm(Xh.astype(np.float64), axis=0)[:, None]
    X = torch.from_numpy(Xh).to(dev); D = torch.from_numpy(Dh).to(dev)
    for impl in (1, 2):
        out = torch.full((C, K), float("nan"), device=dev)
        rc = lib.lys_corr_gemm(X.data_ptr(), X.stride(0), X.stride(1), D.data_ptr(), K, 64, K, C, out.data_ptr(), impl, None)
        torch.cuda.synchronize()
        o = out.cpu().numpy().astype(np.float64)
        err = np.abs(o - ref) / scale
        print("K=%d C=%d impl=%d rc=%d max_rel_err=%.3e mean=%.3e nan=%d" % (K, C, impl, rc, np.nanmax(err), np.nanmean(err), int(np.isnan(o).sum())), flush=True)
    # signal-major X
    Xs = X.t().contiguous().t()
    out = torch.empty((C, K), device=dev)
    lib.lys_corr_gemm(Xs.data_ptr(), Xs.stride(0), Xs.stride(1), D.data_ptr(), K, 64, K, C, out.data_ptr(), 2, None)
    torch.cuda.synchronize()
    print("   signal-major X impl=2 max_rel_err=%.3e" % np.max(np.abs(out.cpu().numpy() - ref) / scale), flush=True)
