"""Kernel-map construction for sparse convolution — the step before `torch.ops.dgsparse_spconv.spconv`.

The reference has a CUDA `sparse_mapping` (src/cuda/sparse_mapping.cu:20-161) that is never registered as an op and
whose output is a dense input-major table; this module exposes the well-defined product spconv consumes
(dgs_kmap_downsample / dgs_kmap_build of include/dgsparse_b200.h):

    km = build_kernel_map(in_coords, kernel_size=3, stride=1)            # submanifold layer
    km = build_kernel_map(in_coords, kernel_size=3, stride=2, padding=1) # general strided layer (expand branch)
    out = torch.ops.dgsparse_spconv.spconv(feats, W, km.kpos, km.qkpos, km.in_map, km.out_map,
                                           km.out_nnz, km.sum_nnz, km.separate_mid, True)

in_coords: int32 [n, 4] = (batch, x, y, z) on a CUDA device.  Deterministic: pairs are grouped by kernel tap
(tap = (kx*ks + ky)*ks + kz) and ordered by output index inside a tap.
"""
import ctypes
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


_ws = {}


def _workspace(nbytes, device):
    # scratch of the build (hash table, hit table, scan temporaries), reused per (device, stream): calls on one stream are ordered
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def _check_coords(c: torch.Tensor, ranges: bool = True):
    """The kernels pack a voxel into one 64-bit key, 16 bits per component: reject what would alias another voxel.
    ranges=False skips the value check (four reductions and a device->host read): build_kernel_map's kernels test every
    coordinate themselves and report through the pair counts it reads back anyway."""
    if c.dtype != torch.int32 or c.dim() != 2 or c.size(1) != 4:
        raise TypeError("in_coords must be int32 [n, 4] = (batch, x, y, z)")
    if c.size(0) == 0 or not ranges:
        return
    xyz, b = c[:, 1:], c[:, 0]
    lo, hi, bmin, bmax = (int(v) for v in torch.stack([xyz.amin(), xyz.amax(), b.amin(), b.amax()]).tolist())
    if lo <= -32768 or hi >= 32768 or bmin < 0 or bmax > 65535:
        raise ValueError(f"voxel coordinates must lie in (-32768, 32768) and batch indices in [0, 65535] "
                         f"(got coordinate range [{lo}, {hi}], batch range [{bmin}, {bmax}])")


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(int(x) for x in v)


def downsample_coords(in_coords: torch.Tensor, stride) -> torch.Tensor:
    """Sorted unique of (batch, floor(x/sx), floor(y/sy), floor(z/sz)) — coordsDownsample + sort + unique,
    src/cuda/sparse_mapping.cu:68-97 (coordinates in OUTPUT resolution)."""
    require_cuda(in_coords)
    _check_coords(in_coords)
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


def expand_coords(in_coords: torch.Tensor, kernel_size, stride, padding=0, min_coord=None, max_coord=None) -> torch.Tensor:
    """Output voxels of a general strided layer — the coordsDownsampleExpand branch, src/cuda/sparse_mapping.cu:98-137:
    sorted unique of (in - off(tap) + padding) / stride over every (input, tap) whose division is exact and whose result
    lies in [min_coord, max_coord] (output resolution; None = unbounded)."""
    require_cuda(in_coords)
    _check_coords(in_coords)
    c = in_coords.contiguous()
    n = c.size(0)
    ks, st, pd = _triple(kernel_size), _triple(stride), _triple(padding)
    k_vol = ks[0] * ks[1] * ks[2]
    lo = (ctypes.c_int * 3)(*(_triple(min_coord) if min_coord is not None else (-(1 << 30),) * 3))
    hi = (ctypes.c_int * 3)(*(_triple(max_coord) if max_coord is not None else ((1 << 30),) * 3))
    with torch.cuda.device(c.device):
        out = torch.empty((max(n * k_vol, 1), 4), dtype=torch.int32, device=c.device)
        cnt = torch.zeros(1, dtype=torch.int32, device=c.device)
        ws = torch.empty(lib.dgs_kmap_expand_workspace_bytes(n, k_vol), dtype=torch.uint8, device=c.device)
        check(lib.dgs_kmap_downsample_expand(n, ptr(c), ks[0], ks[1], ks[2], st[0], st[1], st[2], pd[0], pd[1], pd[2],
                                             ctypes.cast(lo, ctypes.c_void_p), ctypes.cast(hi, ctypes.c_void_p), ptr(out),
                                             ptr(cnt), ptr(ws), ws.numel(), stream_of(c)), "dgs_kmap_downsample_expand")
        return out[: int(cnt.item())].clone()


def build_kernel_map(in_coords: torch.Tensor, kernel_size=3, stride=1, separate_mid: bool = False, q: int = 128,
                     padding=None, min_coord=None, max_coord=None) -> KernelMap:
    """Three kinds of layer, chosen as sparse_mapping does (src/cuda/sparse_mapping.cu:60-137):
      * stride 1, padding None: submanifold layer — out_coords = in_coords, centred taps (the `separate_mid` branch);
      * padding None or 0 and every stride equal to 1 or the kernel size: plain down-sampling — out_coords =
        downsample_coords(in_coords, stride), taps out*stride + off(tap);
      * anything else (e.g. kernel 3, stride 2, padding 1): out_coords = expand_coords(...), the set of voxels whose
        receptive field out*stride - padding + off(tap) holds an input, optionally clipped to [min_coord, max_coord]."""
    require_cuda(in_coords)
    ks, st = _triple(kernel_size), _triple(stride)
    pd = _triple(padding) if padding is not None else (0, 0, 0)
    c = in_coords.contiguous()
    sub = st == (1, 1, 1) and padding is None
    _check_coords(c, ranges=not sub)     # (the down-sampling / expansion kernels pack keys before anything is tested)
    if separate_mid and not sub:
        raise ValueError("separate_mid needs a submanifold layer (stride 1, padding None)")
    plain = pd == (0, 0, 0) and min_coord is None and max_coord is None and all(s in (1, k) for s, k in zip(st, ks))
    if sub:
        out_coords = c
    elif plain:
        out_coords = downsample_coords(c, st)
    else:
        out_coords = expand_coords(c, ks, st, pd, min_coord, max_coord)
    n_in, n_out = c.size(0), out_coords.size(0)
    k_vol = ks[0] * ks[1] * ks[2]
    with torch.cuda.device(c.device):
        dev = c.device
        maps = torch.empty(2, k_vol * n_out, dtype=torch.int32, device=dev)      # one allocation each for the maps and
        imap, omap = maps[0], maps[1]
        counts = torch.zeros(3 * k_vol + 2, dtype=torch.int32, device=dev)       # ... the three small count arrays
        knnz, kpos, qkpos = counts[:k_vol], counts[k_vol:2 * k_vol + 1], counts[2 * k_vol + 1:]
        ws = _workspace(lib.dgs_kmap_workspace_bytes(n_in, n_out, k_vol), dev)
        check(lib.dgs_kmap_build_ex(n_in, ptr(c), n_out, ptr(out_coords), ks[0], ks[1], ks[2], st[0], st[1], st[2],
                                    pd[0], pd[1], pd[2], int(sub), q, int(separate_mid), ptr(imap), ptr(omap), ptr(knnz),
                                    ptr(kpos), ptr(qkpos), ptr(ws), ws.numel(), stream_of(c)), "dgs_kmap_build_ex")
        ends = counts[2 * k_vol::k_vol + 1].cpu()      # kpos[-1], qkpos[-1]: the one device->host read of the build
    pairs, sum_nnz = int(ends[0]), int(ends[1])
    if pairs < 0:   # the kernels found a coordinate outside the 16-bit key range (dgs_kmap_build_ex poisons kpos with -1)
        raise ValueError("kernel map: a coordinate does not fit the 16-bit-per-component voxel key")
    return KernelMap(out_coords, imap[:pairs], omap[:pairs], knnz, kpos, qkpos, n_out, sum_nnz, separate_mid)
