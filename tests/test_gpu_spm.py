"""GPU parity of the ScSPM pooling path (first "next" row) through the
reference-facing class: golden features from the live reference and seeded cases vs the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import lyssa_oracle as lo  # noqa: E402
from lyssandra_b200.sparse_coding import sparse_encoder  # noqa: E402
from lyssandra_b200.feature_extract import (sc_spm_extractor, sc_max_pooling, sum_pooling, average_pooling,  # noqa: E402
                                            max_pooling, l2_normalizer)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(imgs, fe, D, k, levels, op, nrm):
    enc = sparse_encoder(algorithm="bomp", params={"n_nonzero_coefs": k}, verbose=False)
    ex = sc_spm_extractor(feature_extractor=fe, levels=levels, sparse_coder=enc, pooling_operator=op,
                          normalizer=l2_normalizer() if nrm else None)
    Z = ex.encode(imgs, torch.from_numpy(np.ascontiguousarray(D, dtype=np.float32)).to(DEV))
    return Z.cpu().numpy().astype(np.float64)


def test_spm_golden(golden):
    g = golden("spm")
    imgs = [g["img%d" % i] for i in range(int(g["n_imgs"]))]
    fe = lo.grid_descriptor_extractor(step_size=int(g["step_size"]), patch_size=int(g["patch_size"]))
    levels = tuple(int(v) for v in g["levels"])
    for name, op, nrm in (("absmax_l2", sc_max_pooling(), True), ("sum", sum_pooling(), False), ("avg_l2", average_pooling(), True)):
        Z = _run(imgs, fe, g["D"], int(g["k"]), levels, op, nrm)
        ref = g["Z_" + name]
        assert Z.shape == ref.shape
        assert np.array_equal(Z != 0, ref != 0), name                      # same cells / atoms populated
        assert np.max(np.abs(Z - ref)) <= 2e-5 * np.max(np.abs(ref)), (name, np.max(np.abs(Z - ref)))


@pytest.mark.parametrize("levels,K,k", [((1, 2, 4), 256, 5), ((1, 3), 1024, 5), ((2,), 100, 2)])
def test_spm_seeded_vs_oracle(levels, K, k):
    imgs = lo.synthetic_images(7, seed=21, sizes=((64, 48), (50, 50), (37, 71)))
    fe = lo.grid_descriptor_extractor(step_size=3, patch_size=8)
    D = lo.synthetic_dictionary(K, 64, seed=22)
    enc_o = lo.sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
    Zo = lo.sc_spm_extractor(feature_extractor=fe, levels=levels, sparse_coder=enc_o, pooling_operator=lo.sc_max_pooling(),
                             normalizer=None).encode(imgs, D.astype(np.float64))
    Z = _run(imgs, fe, D, k, levels, sc_max_pooling(), False)
    assert Z.shape == (sum(l * l for l in levels) * K, len(imgs))
    # max |z| pooling only moves when a support differs: near-tie columns may change single entries
    bad = np.abs(Z - Zo) > 2e-5 * np.max(np.abs(Zo))
    assert bad.mean() < 1e-3, bad.mean()


def test_spm_errors():
    imgs = lo.synthetic_images(1, seed=1)
    fe = lo.grid_descriptor_extractor()
    D = lo.synthetic_dictionary(256, 64, seed=2)
    with pytest.raises(NotImplementedError):
        _run(imgs, fe, D, 3, (1, 2), max_pooling(), False)


# ---------------------------------------------------------------- dense SIFT producer
def test_dsift_golden(golden):
    from lyssandra_b200.feature_extract import DsiftExtractor
    g = golden("dsift")
    for gs, ps in ((6, 16), (4, 8)):
        f, p = DsiftExtractor(grid_spacing=gs, patch_size=ps).process_image(g["img"])
        ref_f, ref_p = g["feat_%d_%d" % (gs, ps)], g["pos_%d_%d" % (gs, ps)]
        assert tuple(f.shape) == ref_f.shape and np.array_equal(p.cpu().numpy(), ref_p)
        err = np.max(np.abs(f.cpu().numpy().astype(np.float64) - ref_f))
        assert err <= 2e-5, err                      # descriptors are unit-norm, entries <= 0.2 .. 0.5; float32 pipeline


def test_dsift_seeded_and_scspm_pipeline():
    """images -> dense SIFT (device) -> Batch-OMP -> pyramid pooling, vs the oracle pipeline in float64"""
    from lyssandra_b200.feature_extract import dsift_extractor
    imgs = [im * 255.0 for im in lo.synthetic_images(4, seed=31, sizes=((64, 80), (72, 72)))]
    K, k = 256, 4
    D = lo.synthetic_dictionary(K, 128, seed=32)
    enc_o = lo.sparse_encoder("bomp", {"n_nonzero_coefs": k}, verbose=False)
    Zo = lo.sc_spm_extractor(feature_extractor=lo.dsift_extractor(step_size=6, patch_size=16), levels=(1, 2, 4),
                             sparse_coder=enc_o, pooling_operator=lo.sc_max_pooling(),
                             normalizer=lo.l2_normalizer()).encode(imgs, D.astype(np.float64))
    Z = _run(imgs, dsift_extractor(step_size=6, patch_size=16), D, k, (1, 2, 4), sc_max_pooling(), True)
    assert Z.shape == Zo.shape
    bad = np.abs(Z - Zo) > 1e-4 * np.max(np.abs(Zo))
    assert bad.mean() < 5e-3, bad.mean()             # a flipped near-tie support moves a few pooled entries


@pytest.mark.parametrize("ps,gs", [(16, 6), (8, 4), (12, 5)])
def test_dsift_batched_launch_equals_per_image(ps, gs):
    """lys_dsift_batch over runs of same-sized images writes exactly what per-image lys_dsift calls write
    (runs here: 3 x (40,52), 1 x (33,41), 2 x (40,52)); ps = 16 takes the compile-time-sized kernel"""
    import torch
    from lyssandra_b200.feature_extract.dsift import DsiftExtractor
    sizes = ((40, 52), (40, 52), (40, 52), (33, 41), (40, 52), (40, 52))
    imgs = [torch.from_numpy((im * 255.0).astype(np.float32)).to("cuda:0") for im in lo.synthetic_images(6, seed=51, sizes=sizes)]
    ex = DsiftExtractor(grid_spacing=gs, patch_size=ps)
    desc, pos, counts, hw = ex.process_images(imgs)
    assert [tuple(s) for s in hw] == list(sizes) and desc.shape[0] == sum(counts)
    off = 0
    for im, c in zip(imgs, counts):
        d1, p1 = ex.process_image(im)
        assert torch.equal(desc[off:off + c], d1) and torch.equal(pos[off:off + c], p1.t())
        off += c
