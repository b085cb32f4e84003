"""Sparse 3-D convolution — the reference's (unbuilt) `torch.ops.dgsparse_spconv.spconv` boundary
(src/spconv.cpp:18-74) over the B200 C ABI (dgs_spconv_fwd / dgs_spconv_bwd, include/dgsparse_b200.h).

    spconv(in_feats, kernel, kpos, qkpos, in_map, out_map, out_nnz, sum_nnz, separate_mid, arch80) -> out_feats

Same argument order and meaning as the reference.  `arch80=True` selects the tensor-core path (tcgen05
kind::tf32, fp32 accumulate in TMEM; the reference's wmma tf32 kernels), `arch80=False` the exact fp32 FMA
kernel (the reference's `_fgms_fusion_fp32*`).  `set_precision("bf16" | "fp16")` switches the tensor path of fp32
inputs to 16-bit operands.  Half inputs take tcgen05 kind::f16 with fp16 operands (the reference's fp16 wmma kernels,
include/cuda/spconv.cuh:1408-1552: same 10-bit mantissa), bfloat16 inputs with bf16 operands; accumulation is fp32 in
TMEM either way.  Gradients flow to in_feats and kernel (src/spconv.cpp:43-62).

Fixes: out_feats is zero-initialised; `separate_mid` needs no cuBLAS; half inputs are returned in the input
dtype instead of being written as half into an fp32 buffer.
"""
import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_of

_PRECISION = {"fp32": _lib.SPCONV_FP32, "tf32": _lib.SPCONV_TF32, "bf16": _lib.SPCONV_BF16, "fp16": _lib.SPCONV_FP16}
_tensor_precision = _lib.SPCONV_TF32
_ws = {}


def set_precision(name):
    """Operand precision of the tensor-core path used when arch80=True: "tf32" (default), "bf16" or "fp16"."""
    global _tensor_precision
    if name not in ("tf32", "bf16", "fp16"):
        raise ValueError("precision must be 'tf32', 'bf16' or 'fp16'")
    _tensor_precision = _PRECISION[name]


def quantize_kpos(knnz, q=128):
    """kpos and qkpos from the per-offset pair counts (kpos_quantized, test/test_spconv.py:5-14).

    Returns (kpos int32[k_vol+1], qkpos int32[k_vol+1], sum_nnz) on knnz's device."""
    knnz = knnz.to(torch.int64).reshape(-1)
    zero = knnz.new_zeros(1)
    kpos = torch.cat([zero, torch.cumsum(knnz, 0)])
    qkpos = torch.cat([zero, torch.cumsum((knnz + q - 1) // q * q, 0)])
    return kpos.to(torch.int32), qkpos.to(torch.int32), int(qkpos[-1].item())


def _workspace(nbytes, device):
    # one scratch per (device, stream): the kernels of a call read the prepared weights from it, so two streams running
    # spconv concurrently must not share it (reuse on ONE stream is ordered by the stream)
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def _i32c(t, name):
    if t.dtype != torch.int32:
        raise TypeError(f"{name} must be int32 (got {t.dtype})")
    return t.contiguous()


def _prep(in_feats, kernel, kpos, qkpos, in_map, out_map):
    require_cuda(in_feats, kernel, kpos, qkpos, in_map, out_map)
    if in_feats.dim() != 2 or kernel.dim() != 3:
        raise ValueError("in_feats must be [in_nnz, c_in] and kernel [k_vol, c_in, c_out]")
    if in_feats.size(1) != kernel.size(1):
        raise ValueError("Input feature size and kernel size mismatch")  # src/cuda/spconv_cuda.cu:25-27
    return _i32c(kpos, "kpos"), _i32c(qkpos, "qkpos"), _i32c(in_map, "in_map"), _i32c(out_map, "out_map")


def _precision_of(in_feats, arch80):
    if in_feats.dtype == torch.float16:
        return _lib.SPCONV_FP16
    if in_feats.dtype == torch.bfloat16:
        return _lib.SPCONV_BF16
    return _tensor_precision if arch80 else _lib.SPCONV_FP32


def spconv_fwd_fused(in_feats, kernel, kpos, qkpos, in_map, out_map, out_nnz, sum_nnz, separate_mid, arch80,
                     precision=None):
    """spconv_fwd_fused, src/cuda/spconv_cuda.cu:18-187."""
    kpos, qkpos, in_map, out_map = _prep(in_feats, kernel, kpos, qkpos, in_map, out_map)
    prec = _precision_of(in_feats, arch80) if precision is None else _PRECISION[precision]
    x = in_feats.to(torch.float32).contiguous()
    w = kernel.to(torch.float32).contiguous()
    k_vol, c_in, c_out = w.shape
    with torch.cuda.device(x.device):
        out = torch.empty((out_nnz, c_out), dtype=torch.float32, device=x.device)
        ws = _workspace(lib.dgs_spconv_workspace_bytes(max(x.size(0), out_nnz), k_vol, c_in, c_out, prec), x.device)
        check(lib.dgs_spconv_fwd(x.size(0), out_nnz, k_vol, c_in, c_out, ptr(kpos), ptr(qkpos), ptr(in_map), ptr(out_map),
                                 int(sum_nnz), ptr(x), ptr(w), ptr(out), int(bool(separate_mid)), prec, ptr(ws),
                                 ws.numel(), stream_of(x)), "dgs_spconv_fwd")
    return out if in_feats.dtype == torch.float32 else out.to(in_feats.dtype)


def spconv_bwd_fused(out_feats_grad, in_feats, kernel, kpos, qkpos, in_map, out_map, sum_nnz, separate_mid, arch80,
                     need_in=True, need_kernel=True, precision=None):
    """spconv_bwd_fused, src/cuda/spconv_cuda.cu:189-253 -> (in_feats_grad, kernel_grad)."""
    kpos, qkpos, in_map, out_map = _prep(in_feats, kernel, kpos, qkpos, in_map, out_map)
    prec = _precision_of(in_feats, arch80) if precision is None else _PRECISION[precision]
    x = in_feats.to(torch.float32).contiguous()
    w = kernel.to(torch.float32).contiguous()
    g = out_feats_grad.to(torch.float32).contiguous()
    k_vol, c_in, c_out = w.shape
    with torch.cuda.device(x.device):
        gin = torch.empty_like(x) if need_in else None
        gk = torch.empty_like(w) if need_kernel else None
        ws = _workspace(lib.dgs_spconv_workspace_bytes(max(x.size(0), g.size(0)), k_vol, c_in, c_out, prec), x.device)
        check(lib.dgs_spconv_bwd(x.size(0), g.size(0), k_vol, c_in, c_out, ptr(kpos), ptr(qkpos), ptr(in_map),
                                 ptr(out_map), int(sum_nnz), ptr(g), ptr(x), ptr(w), ptr(gin), ptr(gk),
                                 int(bool(separate_mid)), prec, ptr(ws), ws.numel(), stream_of(x)), "dgs_spconv_bwd")
    if gin is not None and in_feats.dtype != torch.float32:
        gin = gin.to(in_feats.dtype)
    if gk is not None and kernel.dtype != torch.float32:
        gk = gk.to(kernel.dtype)
    return gin, gk


class SpConv(torch.autograd.Function):  # src/spconv.cpp:26-64
    @staticmethod
    def forward(ctx, in_feats, kernel, kpos, qkpos, in_map, out_map, out_nnz, sum_nnz, separate_mid, arch80):
        out = spconv_fwd_fused(in_feats, kernel, kpos, qkpos, in_map, out_map, out_nnz, sum_nnz, separate_mid, arch80)
        ctx.sum_nnz, ctx.separate_mid, ctx.arch80 = sum_nnz, separate_mid, arch80
        ctx.save_for_backward(in_feats, kernel, kpos, qkpos, in_map, out_map)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        in_feats, kernel, kpos, qkpos, in_map, out_map = ctx.saved_tensors
        gin, gk = spconv_bwd_fused(grad_out, in_feats, kernel, kpos, qkpos, in_map, out_map, ctx.sum_nnz,
                                   ctx.separate_mid, ctx.arch80, need_in=ctx.needs_input_grad[0],
                                   need_kernel=ctx.needs_input_grad[1])
        return gin, gk, None, None, None, None, None, None, None, None


def spconv(in_feats, kernel, kpos, qkpos, in_map, out_map, out_nnz, sum_nnz, separate_mid, arch80):
    return SpConv.apply(in_feats, kernel, kpos, qkpos, in_map, out_map, out_nnz, sum_nnz, separate_mid, arch80)


_library = torch.library.Library("dgsparse_spconv", "DEF")   # TORCH_LIBRARY(dgsparse_spconv, m), src/spconv.cpp:74
_library.define("spconv(Tensor in_feats, Tensor kernel, Tensor kpos, Tensor qkpos, Tensor in_map, Tensor out_map, "
                "int out_nnz, int sum_nnz, bool separate_mid, bool arch80) -> Tensor")
_library.impl("spconv", spconv, "CompositeImplicitAutograd")
