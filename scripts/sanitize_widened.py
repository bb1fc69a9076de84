"""Small run of the kernels of the widened rows (thresholding coders, dense SIFT, ScSPM pooling) for compute-sanitizer."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lyssandra_b200.sparse_coding import sparse_encoder
from lyssandra_b200.feature_extract import sc_spm_extractor, dsift_extractor, sc_max_pooling, l2_normalizer
from oracle import lyssa_oracle as lo
dev = "cuda:0"
for (n, K, N, alg, params) in ((64, 1024, 1500, "thresh", {"n_nonzero_coefs": 5}),            # fused, CTA pairs, dense rows
                               (48, 512, 700, "thresh", {"n_nonzero_coefs": 10}),             # fused, single CTA
                               (64, 1024, 600, "thresh", {"nonzero_percentage": 0.1}),        # GEMM + generic selection
                               (37, 130, 500, "thresh", {"n_nonzero_coefs": 4}),              # SIMT GEMM + register selection
                               (64, 512, 900, "iht", {"n_nonzero_coefs": 8, "eta": 0.1, "n_iter": 2}),
                               (128, 2048, 300, "iht", {"n_nonzero_coefs": 4, "eta": 0.1, "n_iter": 2})):
    X = torch.from_numpy(np.ascontiguousarray(lo.synthetic_patches(N, n, seed=1))).to(dev)
    D = torch.from_numpy(lo.synthetic_dictionary(K, n, seed=2)).to(dev)
    enc = sparse_encoder(alg, dict(params), verbose=False)
    Z = enc.encode(X, D)
    codes = enc.encode_sparse(X, D)
    torch.cuda.synchronize()
    print("ok", alg, n, K, N, float(Z.abs().sum()), int(codes.idx.sum()))
imgs = [torch.from_numpy(im.astype(np.float32) * 255).to(dev) for im in lo.synthetic_images(5, seed=3, sizes=((64, 80), (48, 48), (70, 50)))]
D = torch.from_numpy(lo.synthetic_dictionary(256, 128, seed=9)).to(dev)
ex = sc_spm_extractor(feature_extractor=dsift_extractor(step_size=6, patch_size=16), levels=(1, 2, 4),
                      sparse_coder=sparse_encoder("bomp", {"n_nonzero_coefs": 5}, verbose=False),
                      pooling_operator=sc_max_pooling(), normalizer=l2_normalizer())
F = ex.encode(imgs, D)
torch.cuda.synchronize()
print("ok scspm", tuple(F.shape), float(F.abs().sum()))
