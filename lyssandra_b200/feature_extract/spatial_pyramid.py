"""ScSPM feature extraction behind the reference's ``sc_spm_extractor`` (lyssa/feature_extract/
spatial_pyramid.py:34-97), B200-style: the descriptors of ALL images of the call are sparse-coded
in one Batch-OMP launch and pooled over the 1 + 4 + 16 (or any ``levels``) cells of every image
by one kernel that works on the sparse codes (``lys_spm_pool``) — the (K x patches) dense code
matrix the reference pools over (:66,:91) is never formed.

Same constructor and ``encode(imgs, dictionary)`` as the reference; returns Z of shape
(sum(levels^2) * n_atoms, n_imgs), level-major / cell-major / atom-minor (:94-96) — a transposed
view of an (n_imgs, n_features) device buffer.  ``feature_extractor`` is any object with a
``patch_size`` attribute and ``extract(img) -> (descriptors (n, P), positions (P, 2))`` (top-left
(row, col) of every patch), exactly what the reference's ``dsift_extractor`` / ``patch_extractor``
return (:15-32); the dense-SIFT producer itself is the next row of SURVEY.md section 8f.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _native as nat
from .. import engine
from .pooling import _pool_op, sc_max_pooling
from .preproc import l2_normalizer


class dsift_extractor(object):
    """lyssa/feature_extract/spatial_pyramid.py:9-20: dense SIFT descriptors (128, P) + top-left positions (P, 2)"""

    def __init__(self, step_size=None, patch_size=None):
        from .dsift import DsiftExtractor
        self.patch_size = patch_size
        self.extractor = DsiftExtractor(grid_spacing=step_size, patch_size=patch_size)

    def extract(self, img):
        dsift_patches, pos = self.extractor.process_image(img, positionNormalize=False)
        return dsift_patches.t(), pos.t()

    def extract_batch(self, imgs):
        """all images of a call at once: descriptors (128, sum P_i) as a view of one signal-major buffer,
        positions (sum P_i, 2), owner image of every patch (host int32), image sizes"""
        desc, pos, counts, sizes = self.extractor.process_images(imgs)
        owner = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
        return desc.t(), pos, torch.from_numpy(owner), sizes


def spm_pool(codes, patch_img, patch_pos, patch_size, img_hw, levels=(1, 2, 4), pooling_operator=None, normalizer=None):
    """Pool sparse codes over the spatial pyramid of every image.

    codes: engine.SparseCodes of all patches; patch_img (P,) int32 image of each patch; patch_pos
    (P, 2) float32 top-left (row, col); img_hw (n_imgs, 2) int32.  Returns F (n_imgs, cells * K)."""
    lib = nat.load()
    op = pooling_operator if pooling_operator is not None else sc_max_pooling()
    if not isinstance(op, _pool_op) or op.mode is None:
        raise NotImplementedError("pooling operator %r is not served by the fused pooling kernel "
                                  "(sc_max_pooling, sum_pooling, average_pooling are)" % (op,))
    if normalizer is not None and not isinstance(normalizer, l2_normalizer):
        raise NotImplementedError("only l2_normalizer (or None) is supported as the per-cell normalizer")
    dev = codes.idx.device
    lev = (ctypes.c_int32 * len(levels))(*[int(v) for v in levels])
    total = lib.lys_spm_total_cells(lev, len(levels))
    if total <= 0:
        raise ValueError("bad pyramid levels %r" % (levels,))
    n_imgs = int(img_hw.shape[0])
    K = codes.n_atoms
    patch_img = patch_img.to(device=dev, dtype=torch.int32).contiguous()
    patch_pos = patch_pos.to(device=dev, dtype=torch.float32).contiguous()
    img_hw = img_hw.to(device=dev, dtype=torch.int32).contiguous()
    F = torch.empty((n_imgs, total * K), dtype=torch.float32, device=dev)
    cnt = torch.empty((n_imgs * total,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.check(lib.lys_spm_pool(engine._ptr(codes.idx), engine._ptr(codes.val), codes.n_signals, codes.k, K,
                                   engine._ptr(patch_img), engine._ptr(patch_pos), float(patch_size),
                                   engine._ptr(img_hw), n_imgs, lev, len(levels), op.mode, 1 if normalizer is not None else 0,
                                   engine._ptr(F), engine._ptr(cnt), engine._stream_ptr(dev)))
    return F


class sc_spm_extractor(object):
    """lyssa/feature_extract/spatial_pyramid.py:34-97"""

    def __init__(self, feature_extractor=None, levels=(1, 2, 4), sparse_coder=None, pooling_operator=None, normalizer=None):
        self.feature_extractor = feature_extractor
        self.levels = levels
        self.sparse_coder = sparse_coder
        self.pooling_operator = pooling_operator
        self.normalizer = normalizer

    def encode(self, imgs, dictionary):
        engine._require_cuda()
        psize = self.feature_extractor.patch_size
        if hasattr(self.feature_extractor, "extract_batch"):             # device producer: no per-image tensors
            X, pos, owner, hw = self.feature_extractor.extract_batch(imgs)
            D = engine.as_dictionary(dictionary, X.device)
            codes = self.sparse_coder.encode_sparse(X, D)
            F = spm_pool(codes, owner, pos, psize, torch.tensor(hw, dtype=torch.int32),
                         self.levels, self.pooling_operator, self.normalizer)
            return F.t()
        descs, poss, owner, hw = [], [], [], []
        for i, img in enumerate(imgs):                                   # :55-63 (producer side, per image)
            desc, pos = self.feature_extractor.extract(img)
            desc = torch.as_tensor(np.asarray(desc) if not torch.is_tensor(desc) else desc)
            pos = torch.as_tensor(np.asarray(pos) if not torch.is_tensor(pos) else pos)
            descs.append(desc)
            poss.append(pos.reshape(-1, 2))
            owner.append(torch.full((desc.shape[1],), i, dtype=torch.int32))
            hw.append([int(np.shape(img)[0]), int(np.shape(img)[1])])
        D = engine.as_dictionary(dictionary, None if not (torch.is_tensor(dictionary) and dictionary.is_cuda) else dictionary.device)
        dev = D.device
        X = torch.cat([d.to(device=dev, dtype=torch.float32) for d in descs], dim=1)          # (n, all patches)
        codes = self.sparse_coder.encode_sparse(X, D)                                           # :66, one launch
        F = spm_pool(codes, torch.cat(owner), torch.cat(poss), psize, torch.tensor(hw, dtype=torch.int32),
                     self.levels, self.pooling_operator, self.normalizer)                       # :70-96
        return F.t()


def pyramid_feat_extract(imgs, extractor, D):
    """lyssa/feature_extract/spatial_pyramid.py:100-101"""
    return extractor.encode(imgs, D)
