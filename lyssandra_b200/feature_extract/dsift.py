"""Dense SIFT behind the reference's ``DsiftExtractor`` (lyssa/feature_extract/dsift.py:38-162): same
constructor and ``process_image``; the gradient / orientation maps, the bilinear spatial binning and Lowe's
normalisation run in two CUDA kernels (``lys_dsift``), descriptors and positions stay on the device.

Host-side glue computed with the reference's own formulas (a few dozen floats): the 5x5 Gaussian-derivative
kernels of ``gen_dgauss`` (:23-35) and the separable factor of the bilinear weight matrix (:57-73)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from .. import _native as nat
from .. import engine

n_angles = 8
n_bins = 4
n_samples = n_bins ** 2


def gen_dgauss(sigma):
    """gradient of the gaussian in both directions (dsift.py:23-35)"""
    fwid = int(2 * np.ceil(sigma))
    G = np.array(range(-fwid, fwid + 1)) ** 2
    G = G.reshape((G.size, 1)) + G
    G = np.exp(- G / 2.0 / sigma / sigma)
    G /= np.sum(G)
    GH, GW = np.gradient(G)
    GH *= 2.0 / np.sum(np.abs(GH))
    GW *= 2.0 / np.sum(np.abs(GW))
    return GH, GW


def bin_weights(patch_size):
    """w[bin][pixel]: the per-axis factor of the (16, ps^2) weight matrix of dsift.py:57-73"""
    sample_res = patch_size / np.double(n_bins)
    bincenter = np.array(range(1, n_bins * 2, 2)) / 2.0 / n_bins * patch_size - 0.5
    dist = np.abs(np.arange(patch_size)[None, :] - bincenter[:, None]) / sample_res
    return (1 - dist) * (dist <= 1)


class DsiftExtractor(object):
    def __init__(self, grid_spacing=None, patch_size=None, nrml_thres=1.0, sigma_edge=0.8, sift_thres=0.2):
        self.gs = grid_spacing
        self.ps = patch_size
        self.nrml_thres = nrml_thres
        self.sigma = sigma_edge
        self.sift_thres = sift_thres
        if int(2 * np.ceil(sigma_edge)) != 2:
            raise NotImplementedError("the device kernel is built for the reference's 5x5 derivative kernels (sigma_edge <= 1)")
        gh, gw = gen_dgauss(self.sigma)
        self._gh = np.ascontiguousarray(gh, dtype=np.float32)
        self._gw = np.ascontiguousarray(gw, dtype=np.float32)
        self._wt = np.ascontiguousarray(bin_weights(self.ps), dtype=np.float32)

    @staticmethod
    def _device_image(image, device=None):
        """grayscale float32 contiguous CUDA tensor of one image (:92-94)"""
        if torch.is_tensor(image):
            img = image
        else:
            img = torch.from_numpy(np.ascontiguousarray(image))
        if img.dim() == 3:
            img = img.to(torch.float64).mean(dim=2)                      # :92-94 grayscale
        if device is None:
            device = img.device if img.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if img.is_cuda and img.dtype == torch.float32 and img.device == device and img.is_contiguous():
            return img
        return img.to(device=device, dtype=torch.float32).contiguous()

    def _grid(self, lib, H, W):
        nh, nw, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        nat.check(lib.lys_dsift_grid(H, W, int(self.gs), int(self.ps), ctypes.byref(nh), ctypes.byref(nw),
                                     ctypes.byref(oh), ctypes.byref(ow)))
        return nh.value, nw.value

    def process_images(self, images, device=None):
        """Descriptors of a list of images into ONE signal-major buffer (what the ScSPM encoder consumes):
        -> desc (sum P_i, 128), pos (sum P_i, 2) top-left (row, col), counts [P_i], sizes [(H_i, W_i)].
        One pair of kernel launches per run of same-sized images, no per-image allocations or concatenation."""
        engine._require_cuda()
        lib = nat.load()
        imgs = [self._device_image(im, device) for im in images]
        if not imgs:
            raise ValueError("no images")
        dev = imgs[0].device
        sizes = [(int(t.shape[0]), int(t.shape[1])) for t in imgs]
        counts = []
        for H, W in sizes:
            nh, nw = self._grid(lib, H, W)
            counts.append(nh * nw)
        total = int(sum(counts))
        desc = torch.empty((total, n_samples * n_angles), dtype=torch.float32, device=dev)
        pos = torch.empty((total, 2), dtype=torch.float32, device=dev)
        # runs of consecutive images of one size and row stride go through ONE pair of launches
        runs, i = [], 0
        while i < len(imgs):
            j = i + 1
            while j < len(imgs) and sizes[j] == sizes[i] and imgs[j].stride(0) == imgs[i].stride(0):
                j += 1
            runs.append((i, j))
            i = j
        def ws_need(a, b):                                   # orientation maps of a whole run, capped at 2 GB
            one = lib.lys_dsift_workspace_bytes(sizes[a][0], sizes[a][1])
            return one * max(1, min(b - a, 128, (2 << 30) // max(one, 1)))
        wsb = max(ws_need(a, b) for a, b in runs)
        ws = engine.workspace(dev, wsb, tag="dsift")
        dptr, pptr, wptr, wbytes = desc.data_ptr(), pos.data_ptr(), engine._ptr(ws), ws.numel()
        gs, ps, nt, st = int(self.gs), int(self.ps), float(self.nrml_thres), float(self.sift_thres)
        gh, gw, wt = self._gh.ctypes.data, self._gw.ctypes.data, self._wt.ctypes.data
        off = 0
        with torch.cuda.device(dev):
            stream = engine._stream_ptr(dev)
            for a, b in runs:
                H, W = sizes[a]
                ptrs = (ctypes.c_void_p * (b - a))(*[imgs[q].data_ptr() for q in range(a, b)])
                nat.check(lib.lys_dsift_batch(ptrs, b - a, imgs[a].stride(0), H, W, gs, ps, nt, st, gh, gw, wt,
                                              dptr + off * 512, pptr + off * 8, wptr, wbytes, stream))
                off += sum(counts[a:b])
        return desc, pos, counts, sizes

    def process_image(self, image, positionNormalize=False, device=None):
        """-> (feat_arr (P, 128) CUDA tensor, positions (2, P) CUDA tensor), dsift.py:75-118"""
        engine._require_cuda()
        lib = nat.load()
        img = self._device_image(image, device)
        device = img.device
        H, W = int(img.shape[0]), int(img.shape[1])
        nh, nw = self._grid(lib, H, W)
        P = nh * nw
        desc = torch.empty((P, n_samples * n_angles), dtype=torch.float32, device=device)
        pos = torch.empty((P, 2), dtype=torch.float32, device=device)
        wsb = lib.lys_dsift_workspace_bytes(H, W)
        ws = engine.workspace(device, wsb, tag="dsift")
        with torch.cuda.device(device):
            nat.check(lib.lys_dsift(engine._ptr(img), img.stride(0), H, W, int(self.gs), int(self.ps),
                                    float(self.nrml_thres), float(self.sift_thres),
                                    self._gh.ctypes.data, self._gw.ctypes.data, self._wt.ctypes.data,
                                    engine._ptr(desc), engine._ptr(pos), engine._ptr(ws), ws.numel(),
                                    engine._stream_ptr(device)))
        positions = pos.t()
        if positionNormalize:
            positions = positions / torch.tensor([[float(H)], [float(W)]], device=device)
        return desc, positions
