"""GPU: the tcgen05 (bf16x3) correlation GEMM against float64 and against the fp32 SIMT GEMM."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import lyssa_oracle as lo  # noqa: E402
from lyssandra_b200 import _native  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,K,C", [(64, 256, 1000), (64, 1024, 4099), (64, 2048, 777), (64, 512, 128), (64, 1024, 1),
                                   (128, 1024, 3001), (128, 2048, 515)])        # n = 128: two accumulated 64-feature halves
def test_tcgen05_gemm_is_fp32_faithful(n, K, C):
    lib = _native.load()
    Xh = np.ascontiguousarray(lo.synthetic_patches(C, n, seed=3)); Dh = lo.synthetic_dictionary(K, n, seed=4)
    ref = Xh.astype(np.float64).T @ Dh.astype(np.float64)
    scale = np.linalg.norm(Xh.astype(np.float64), axis=0)[:, None]         # |x| |d|, |d| = 1
    X = torch.from_numpy(Xh).to(DEV); D = torch.from_numpy(Dh).to(DEV)
    errs = {}
    for impl in (1, 2):
        for Xv in (X, X.t().contiguous().t()):                            # feature-major and signal-major
            out = torch.full((C, K), float("nan"), device=DEV)
            _native.check(lib.lys_corr_gemm(Xv.data_ptr(), Xv.stride(0), Xv.stride(1), D.data_ptr(), K, n, K, C,
                                            out.data_ptr(), impl, None))
            torch.cuda.synchronize()
            o = out.cpu().numpy().astype(np.float64)
            assert not np.isnan(o).any()
            errs[impl] = max(errs.get(impl, 0.0), float(np.max(np.abs(o - ref) / scale)))
    # tolerance stated: 4 fp32 ulps of |x||d| (fp32 FFMA chain measures ~1e-7 here)
    assert errs[1] <= 5e-7 and errs[2] <= 5e-7, errs
