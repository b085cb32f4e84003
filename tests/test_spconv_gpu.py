"""GPU parity: sparse-convolution gather-GEMM-scatter (tcgen05 tf32 / bf16 and exact fp32 paths) vs the oracle.

Oracle: oracle_spconv (C restatement of cpu_compute, test/test_spconv.py:17-53) and an fp64 numpy restatement
of the same loop.  Tolerances (stated per precision, relative to sum_p |in| . |W| of each output element, which
bounds the rounding error of a dot product rigorously):
  fp32  1e-5   (exact FMA; only the accumulation order differs: atomics)
  tf32  1.5e-3 (operands rounded to 10 mantissa bits: 2 * 2^-11 per product; the reference's own (disabled)
               check used rtol 1e-2, test/test_spconv.py:157-158)
  fp16  1.5e-3 (operands rounded to 10 mantissa bits like tf32; inputs here are far inside the fp16 range)
  bf16  1.2e-2 (operands rounded to 7 mantissa bits: 2 * 2^-8 per product)
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"fp32": 1e-5, "tf32": 1.5e-3, "fp16": 1.5e-3, "bf16": 1.2e-2}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ref64(kpos, imap, omap, x, w, out_nnz, mid=None):
    """fp64 restatement of cpu_compute; also returns the |.| bound used for the tolerance."""
    x64, w64 = x.astype(np.float64), w.astype(np.float64)
    out = np.zeros((out_nnz, w.shape[2]))
    bound = np.zeros_like(out)
    for k in range(w.shape[0]):
        s, e = int(kpos[k]), int(kpos[k + 1])
        if e > s:
            np.add.at(out, omap[s:e], x64[imap[s:e]] @ w64[k])
            np.add.at(bound, omap[s:e], np.abs(x64[imap[s:e]]) @ np.abs(w64[k]))
    if mid is not None:
        out += x64 @ w64[mid]
        bound += np.abs(x64) @ np.abs(w64[mid])
    return out, bound


def make_maps(rng, in_nnz, out_nnz, k_vol, density, drop_mid=False):
    """Random kernel maps: per offset a random subset of outputs, each fed by a random input row."""
    imap, omap, knnz = [], [], []
    for k in range(k_vol):
        n = 0 if (drop_mid and k == k_vol // 2) else int(rng.integers(0, int(density * out_nnz) + 1))
        o = rng.choice(out_nnz, size=n, replace=False).astype(np.int32)
        i = rng.integers(0, in_nnz, size=n).astype(np.int32)
        imap.append(i); omap.append(np.sort(o)); knnz.append(n)
    return np.concatenate(imap), np.concatenate(omap), np.array(knnz, np.int64)


def run_fwd(x, w, knnz, imap, omap, out_nnz, precision, separate_mid=False):
    import dgsparse.spconv as S
    kpos, qkpos, sum_nnz = S.quantize_kpos(torch.from_numpy(knnz).cuda())
    out = S.spconv_fwd_fused(dev(x), dev(w), kpos, qkpos, dev(imap), dev(omap), out_nnz, sum_nnz, separate_mid,
                             precision != "fp32", precision=precision)
    torch.cuda.synchronize()
    return out.cpu().numpy(), kpos.cpu().numpy()


def check(got, want, bound, precision, what):
    err = np.abs(got.astype(np.float64) - want)
    lim = TOL[precision] * bound + 1e-30
    bad = err > lim
    assert not bad.any(), f"{what} [{precision}]: {int(bad.sum())}/{bad.size} outside tolerance, " \
                          f"max err/bound {float((err / (bound + 1e-30)).max()):.3e}"


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
@pytest.mark.parametrize("c_in,c_out", [(64, 64), (4, 64), (32, 96), (128, 32), (96, 256), (16, 16), (200, 72)])
def test_forward_random_maps(precision, c_in, c_out):
    rng = np.random.default_rng(c_in * 1000 + c_out)
    in_nnz, out_nnz, k_vol = 3000, 2500, 27
    imap, omap, knnz = make_maps(rng, in_nnz, out_nnz, k_vol, 0.3)
    x = rng.uniform(-1, 1, (in_nnz, c_in)).astype(np.float32)
    w = rng.uniform(-1, 1, (k_vol, c_in, c_out)).astype(np.float32)
    got, kpos = run_fwd(x, w, knnz, imap, omap, out_nnz, precision)
    want, bound = ref64(kpos, imap, omap, x, w, out_nnz)
    check(got, want, bound, precision, f"fwd {c_in}x{c_out}")


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_forward_matches_c_oracle_small(oracle, precision):
    """Against oracle_spconv (the C restatement of cpu_compute) itself, on a case it finishes instantly."""
    rng = np.random.default_rng(5)
    in_nnz, out_nnz, k_vol, c_in, c_out = 700, 650, 27, 64, 64
    imap, omap, knnz = make_maps(rng, in_nnz, out_nnz, k_vol, 0.5)
    x = rng.uniform(0, 1, (in_nnz, c_in)).astype(np.float32)
    w = rng.uniform(0, 1, (k_vol, c_in, c_out)).astype(np.float32)
    got, kpos = run_fwd(x, w, knnz, imap, omap, out_nnz, precision)
    want = oracle.spconv(kpos, imap, omap, x, w, out_nnz)
    rtol = 1e-5 if precision == "fp32" else 1.5e-3      # all-positive data: bound == |want|
    assert np.allclose(got, want, rtol=rtol, atol=1e-6)


@pytest.mark.parametrize("separate", [False, True])
@pytest.mark.parametrize("idx", [0, 1])
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16", "fp16"])
def test_fixture_minkunet_layers(precision, idx, separate):
    """The reference's own kernel maps (example/data/sample-data/fp32/minkunet-semantickitti/*.pth re-saved as
    tests/golden/spconv_fp32_*.npz), driven exactly as test/test_spconv.py:100-147 does (random feats/weights)."""
    g = np.load(os.path.join(GOLDEN, f"spconv_fp32_{idx}.npz"))
    in_nnz, out_nnz, k_vol, c_in, c_out = (int(g[k]) for k in ("in_nnz", "out_nnz", "k_vol", "c_in", "c_out"))
    knnz, imap, omap = g["knnz"].astype(np.int64), g["imap"], g["omap"]
    rng = np.random.default_rng(idx)
    x = rng.uniform(0, 1, (in_nnz, c_in)).astype(np.float32)
    w = rng.uniform(0, 1, (k_vol, c_in, c_out)).astype(np.float32)
    mid = k_vol // 2
    if separate:   # submanifold layer run the reference's separate_mid way: centre offset out of the maps
        assert in_nnz == out_nnz
        s, e = int(g["kpos"][mid]), int(g["kpos"][mid + 1])
        assert np.array_equal(imap[s:e], omap[s:e])          # the centre offset is the identity map
        imap, omap = np.delete(imap, np.s_[s:e]), np.delete(omap, np.s_[s:e])
        knnz = knnz.copy(); knnz[mid] = 0
    separate_mid = separate
    got, kpos = run_fwd(x, w, knnz, imap, omap, out_nnz, precision, separate_mid=separate_mid)
    want, bound = ref64(kpos, imap, omap, x, w, out_nnz, mid=mid if separate_mid else None)
    check(got, want, bound, precision, f"fixture {idx}")


def test_empty_offsets_and_ragged_tail():
    """Offsets with zero pairs, one with a single pair, one exactly 128, one 129: the qkpos tile table edge cases."""
    rng = np.random.default_rng(9)
    in_nnz = out_nnz = 400
    knnz = np.array([0, 1, 128, 129, 0, 0, 255, 0, 3], np.int64)
    imap = rng.integers(0, in_nnz, int(knnz.sum())).astype(np.int32)
    omap = np.concatenate([np.sort(rng.choice(out_nnz, int(n), replace=False)) for n in knnz]).astype(np.int32)
    x = rng.uniform(-1, 1, (in_nnz, 64)).astype(np.float32)
    w = rng.uniform(-1, 1, (9, 64, 64)).astype(np.float32)
    for precision in ("fp32", "tf32"):
        got, kpos = run_fwd(x, w, knnz, imap, omap, out_nnz, precision)
        want, bound = ref64(kpos, imap, omap, x, w, out_nnz)
        check(got, want, bound, precision, "ragged")
        untouched = np.setdiff1d(np.arange(out_nnz), omap)
        assert not got[untouched].any()          # rows no pair maps to are exactly zero (output is zero-initialised)


@pytest.mark.parametrize("arch80", [False, True])
def test_torch_op_autograd(arch80):
    """torch.ops.dgsparse_spconv.spconv forward + backward (in_feats and kernel gradients) vs fp64 autograd."""
    import dgsparse.spconv as S
    rng = np.random.default_rng(11)
    in_nnz, out_nnz, k_vol, c_in, c_out = 900, 800, 27, 32, 64
    imap, omap, knnz = make_maps(rng, in_nnz, out_nnz, k_vol, 0.4)
    kpos, qkpos, sum_nnz = S.quantize_kpos(torch.from_numpy(knnz).cuda())
    x = torch.tensor(rng.uniform(-1, 1, (in_nnz, c_in)), dtype=torch.float32, device="cuda", requires_grad=True)
    w = torch.tensor(rng.uniform(-1, 1, (k_vol, c_in, c_out)), dtype=torch.float32, device="cuda", requires_grad=True)
    gout = torch.tensor(rng.uniform(-1, 1, (out_nnz, c_out)), dtype=torch.float32, device="cuda")
    out = torch.ops.dgsparse_spconv.spconv(x, w, kpos, qkpos, dev(imap), dev(omap), out_nnz, sum_nnz, False, arch80)
    out.backward(gout)
    # fp64 reference through plain torch ops
    x64 = x.detach().double().requires_grad_()
    w64 = w.detach().double().requires_grad_()
    o64 = torch.zeros(out_nnz, c_out, dtype=torch.float64, device="cuda")
    kp = kpos.cpu().numpy()
    im, om = dev(imap).long(), dev(omap).long()
    for k in range(k_vol):
        s, e = int(kp[k]), int(kp[k + 1])
        if e > s:
            o64 = o64.index_add(0, om[s:e], x64[im[s:e]] @ w64[k])
    o64.backward(gout.double())
    tol = 1.5e-3 if arch80 else 1e-5
    scale_o = float(o64.abs().max()); scale_x = float(x64.grad.abs().max()); scale_w = float(w64.grad.abs().max())
    assert (out.double() - o64).abs().max().item() <= tol * scale_o * 4
    assert (x.grad.double() - x64.grad).abs().max().item() <= tol * scale_x * 4
    assert (w.grad.double() - w64.grad).abs().max().item() <= tol * scale_w * 4     # tf32 tensor path / fp32 FMA


@pytest.mark.parametrize("c_in,c_out", [(64, 64), (32, 96), (128, 128), (4, 64), (200, 72), (64, 256)])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_kernel_gradient(precision, c_in, c_out):
    """dW[k] = sum_p in[imap[p]]^T (x) gout[omap[p]]: tensor-core (tf32, MN-major operands) and fp32 FMA kernels vs fp64."""
    import dgsparse.spconv as S
    rng = np.random.default_rng(c_in * 7 + c_out)
    in_nnz, out_nnz, k_vol = 2000, 1800, 27
    imap, omap, knnz = make_maps(rng, in_nnz, out_nnz, k_vol, 0.3)
    kpos, qkpos, sum_nnz = S.quantize_kpos(torch.from_numpy(knnz).cuda())
    x = rng.uniform(-1, 1, (in_nnz, c_in)).astype(np.float32)
    g = rng.uniform(-1, 1, (out_nnz, c_out)).astype(np.float32)
    w = np.zeros((k_vol, c_in, c_out), np.float32)
    _, gk = S.spconv_bwd_fused(dev(g), dev(x), dev(w), kpos, qkpos, dev(imap), dev(omap), sum_nnz, False, True,
                               need_in=False, precision=precision)
    torch.cuda.synchronize()
    kp = kpos.cpu().numpy()
    want = np.zeros((k_vol, c_in, c_out))
    bound = np.zeros_like(want)
    for k in range(k_vol):
        s, e = int(kp[k]), int(kp[k + 1])
        if e > s:
            want[k] = x[imap[s:e]].astype(np.float64).T @ g[omap[s:e]].astype(np.float64)
            bound[k] = np.abs(x[imap[s:e]]).astype(np.float64).T @ np.abs(g[omap[s:e]]).astype(np.float64)
    check(gk.cpu().numpy(), want, bound, precision, f"dW {c_in}x{c_out}")


def test_errors():
    import dgsparse.spconv as S
    x = torch.zeros(10, 8, device="cuda"); w = torch.zeros(3, 4, 8, device="cuda")
    z = torch.zeros(4, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        S.spconv_fwd_fused(x, w, z, z, z, z, 10, 0, False, True)          # c_in mismatch, as the reference throws
    with pytest.raises(RuntimeError):
        S.spconv_fwd_fused(x.cpu(), torch.zeros(3, 8, 8), z, z, z, z, 10, 0, False, True)   # no CPU path


@pytest.mark.parametrize("dtype,prec", [(torch.float16, "fp16"), (torch.bfloat16, "bf16")])
def test_half_inputs_keep_their_operand_format(dtype, prec):
    """Half tensors through the op boundary (test/test_spconv.py:100-147 with precision = 'fp16'): fp16 inputs are multiplied
    as fp16 operands (tcgen05 kind::f16, like the reference's fp16 wmma kernels), bf16 inputs as bf16; fp32 accumulation; the
    result comes back in the input dtype.  The already-rounded inputs make the products exact, so the only error left is the
    accumulation order and the final rounding to the 16-bit output."""
    import dgsparse  # noqa: F401  (registers torch.ops.dgsparse_spconv.spconv)
    import dgsparse.spconv as S
    rng = np.random.default_rng(11)
    in_nnz, out_nnz, k_vol, c_in, c_out = 2000, 1800, 27, 64, 96
    imap, omap, knnz = make_maps(rng, in_nnz, out_nnz, k_vol, 0.3)
    x = torch.tensor(rng.uniform(-1, 1, (in_nnz, c_in)), dtype=torch.float32, device="cuda").to(dtype)
    w = torch.tensor(rng.uniform(-1, 1, (k_vol, c_in, c_out)), dtype=torch.float32, device="cuda").to(dtype)
    kpos, qkpos, sum_nnz = S.quantize_kpos(torch.from_numpy(knnz).cuda())
    out = torch.ops.dgsparse_spconv.spconv(x, w, kpos, qkpos, dev(imap), dev(omap), out_nnz, sum_nnz, False, True)
    assert out.dtype == dtype
    want, bound = ref64(kpos.cpu().numpy(), imap, omap, x.float().cpu().numpy(), w.float().cpu().numpy(), out_nnz)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7          # rounding of the OUTPUT to 16 bits
    err = np.abs(out.float().cpu().numpy().astype(np.float64) - want)
    assert (err <= 1e-5 * bound + ulp * np.abs(want) + 1e-30).all(), float((err / (bound + 1e-30)).max())
