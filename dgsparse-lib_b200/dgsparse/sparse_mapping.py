"""Kernel-map construction for sparse convolution — the step before `torch.ops.dgsparse_spconv.spconv`.

The reference has a CUDA `sparse_mapping` (src/cuda/sparse_mapping.cu:20-161) that is never registered as an op and
whose output is a dense input-major table; this module exposes the well-defined product spconv consumes
(dgs_kmap_downsample / dgs_kmap_build of include/dgsparse_b200.h):

    km = build_kernel_map(in_coords, kernel_size=3, stride=1)            # submanifold layer
    out = torch.ops.dgsparse_spconv.spconv(feats, W, km.kpos, km.qkpos, km.in_map, km.out_map,
                                           km.out_nnz, km.sum_nnz, km.separate_mid, True)

in_coords: int32 [n, 4] = (batch, x, y, z) on a CUDA device.  Deterministic: pairs are grouped by kernel tap
(tap = (kx*ks + ky)*ks + kz) and ordered by output index inside a tap.
"""
from dataclasses import dataclass

import torch

from ._lib import check, lib, ptr, require_cuda, stream_of


@dataclass
class KernelMap:
    out_coords: torch.Tensor   # int32 [out_nnz, 4]
    in_map: torch.Tensor       # int32 [pairs]
    out_map: torch.Tensor      # int32 [pairs]
    knnz: torch.Tensor         # int32 [k_vol]
    kpos: torch.Tensor         # int32 [k_vol + 1]
    qkpos: torch.Tensor        # int32 [k_vol + 1], counts rounded up to 128
    out_nnz: int
    sum_nnz: int               # qkpos[-1]
    separate_mid: bool


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(int(x) for x in v)


def downsample_coords(in_coords: torch.Tensor, stride) -> torch.Tensor:
    """Sorted unique of (batch, floor(x/sx), floor(y/sy), floor(z/sz)) — coordsDownsample + sort + unique,
    src/cuda/sparse_mapping.cu:68-97 (coordinates in OUTPUT resolution)."""
    require_cuda(in_coords)
    if in_coords.dtype != torch.int32 or in_coords.dim() != 2 or in_coords.size(1) != 4:
        raise TypeError("in_coords must be int32 [n, 4] = (batch, x, y, z)")
    c = in_coords.contiguous()
    n = c.size(0)
    sx, sy, sz = _triple(stride)
    with torch.cuda.device(c.device):
        out = torch.empty((n, 4), dtype=torch.int32, device=c.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=c.device)
        ws = torch.empty(lib.dgs_kmap_workspace_bytes(n, n, 1), dtype=torch.uint8, device=c.device)
        check(lib.dgs_kmap_downsample(n, ptr(c), sx, sy, sz, ptr(out), ptr(cnt), ptr(ws), ws.numel(), stream_of(c)),
              "dgs_kmap_downsample")
        return out[: int(cnt.item())]


def build_kernel_map(in_coords: torch.Tensor, kernel_size=3, stride=1, separate_mid: bool = False, q: int = 128) -> KernelMap:
    """stride 1: submanifold layer (out_coords = in_coords, centred taps); stride > 1: down-sampling layer
    (out_coords = downsample_coords(in_coords, stride), taps out*stride + [0, kernel_size))."""
    require_cuda(in_coords)
    ks, st = _triple(kernel_size), _triple(stride)
    c = in_coords.contiguous()
    if c.dtype != torch.int32 or c.dim() != 2 or c.size(1) != 4:
        raise TypeError("in_coords must be int32 [n, 4] = (batch, x, y, z)")
    sub = st == (1, 1, 1)
    if separate_mid and not sub:
        raise ValueError("separate_mid needs a submanifold layer (stride 1)")
    out_coords = c if sub else downsample_coords(c, st)
    n_in, n_out = c.size(0), out_coords.size(0)
    k_vol = ks[0] * ks[1] * ks[2]
    with torch.cuda.device(c.device):
        dev = c.device
        imap = torch.empty(k_vol * n_out, dtype=torch.int32, device=dev)
        omap = torch.empty(k_vol * n_out, dtype=torch.int32, device=dev)
        knnz = torch.zeros(k_vol, dtype=torch.int32, device=dev)
        kpos = torch.zeros(k_vol + 1, dtype=torch.int32, device=dev)
        qkpos = torch.zeros(k_vol + 1, dtype=torch.int32, device=dev)
        ws = torch.empty(lib.dgs_kmap_workspace_bytes(n_in, n_out, k_vol), dtype=torch.uint8, device=dev)
        check(lib.dgs_kmap_build(n_in, ptr(c), n_out, ptr(out_coords), ks[0], ks[1], ks[2], st[0], st[1], st[2], q,
                                 int(separate_mid), ptr(imap), ptr(omap), ptr(knnz), ptr(kpos), ptr(qkpos), ptr(ws),
                                 ws.numel(), stream_of(c)), "dgs_kmap_build")
        ends = torch.stack([kpos[-1], qkpos[-1]]).cpu()
    pairs, sum_nnz = int(ends[0]), int(ends[1])
    return KernelMap(out_coords, imap[:pairs], omap[:pairs], knnz, kpos, qkpos, n_out, sum_nnz, separate_mid)
