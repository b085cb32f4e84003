"""GPU parity against the REFERENCE'S OWN sparse-convolution CUDA: kernel-map construction (sparse_mapping,
src/cuda/sparse_mapping.cu:20-161) and the fused gather-GEMM-scatter (spconv_fwd_fused / spconv_bwd_fused,
src/cuda/spconv_cuda.cu:18-253), compiled UNMODIFIED for sm_100a where they lie (oracle/build_ref_spconv.sh ->
oracle/_ref/_ref_spconv.so, a no-algorithm pybind shim; the reference never builds or registers either).  This pins rows
a11 / f1 / f4 of SURVEY.md §8 to the reference itself instead of to restatements.

Kernel maps are integer work: pair SETS per tap, knnz, kpos, qkpos and the output coordinates must be identical (the
reference's order inside a tap is decided by atomics, ours is sorted by output index; the reference's table is
input-major `map[k][in] = out`).  Compared where the reference is self-consistent: the submanifold branch
(`_queryhash_subm`), the stride-1 non-submanifold branch (`_queryhash_sp`, padding) and the general strided branch
(`coordsDownsampleExpand` + `_queryhash_sp`); its plain down-sampling branch emits output voxels in INPUT resolution and
then multiplies them by the stride again (sparse_mapping.cuh:296-323 vs :171-173), so there only the voxel set is compared.

spconv: the reference allocates its output with torch::empty and accumulates into it with atomics; the
harness hands it a pre-zeroed allocator block.  fp32 SIMT path: 1e-5 of sum |terms|; tf32 paths: each side within the
tf32 bound of the fp64 oracle and of each other.
"""
import os

import numpy as np
import pytest
import torch

from test_spconv_gpu import TOL, check, ref64

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def REF(oracle):
    m = oracle.ref_spconv_module()
    if m is None:
        pytest.skip("oracle/_ref/_ref_spconv.so not built (oracle/build_ref_spconv.sh)")
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def random_coords(rng, n, batch, extent):
    c = np.stack([rng.integers(0, batch, n), rng.integers(0, extent, n), rng.integers(0, extent, n), rng.integers(0, extent, n)], 1)
    return np.unique(c, axis=0).astype(np.int32)


def reference_map(REF, in_c, ks, l_stride, pad, lo, hi, separate_mid):
    """Drive sparse_mapping the way its signature asks (src/cuda/sparse_mapping.cu:20-29): caller-allocated map (-1),
    kernel_nnz (0), kernel_pos / kernel_kpos (0).  Returns (out_coords, [pair set per tap], knnz, kpos, qkpos)."""
    n, k_vol = in_c.shape[0], ks ** 3
    d_in = dev(in_c)
    m = torch.full((k_vol * n,), -1, dtype=torch.int32, device="cuda")
    knnz = torch.zeros(k_vol, dtype=torch.int32, device="cuda")
    kpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
    qkpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
    i3 = lambda v: torch.tensor([v, v, v] if isinstance(v, int) else list(v), dtype=torch.int32, device="cuda")
    out_c = REF.sparse_mapping(d_in, int(in_c[:, 0].max()) + 1, ks, ks, ks, k_vol, 4, 4, l_stride, l_stride, l_stride, 1, 1, 1,
                               i3(pad), i3(lo), i3(hi), m, knnz, kpos, qkpos, separate_mid)
    torch.cuda.synchronize()
    table = m.cpu().numpy().reshape(k_vol, n)
    pairs = []
    for k in range(k_vol):
        ins = np.nonzero(table[k] >= 0)[0]
        pairs.append(set(zip(ins.tolist(), table[k][ins].tolist())))
    return out_c.cpu().numpy(), pairs, knnz.cpu().numpy(), kpos.cpu().numpy(), qkpos.cpu().numpy()


def our_pairs(km):
    kpos, imap, omap = km.kpos.cpu().numpy(), km.in_map.cpu().numpy(), km.out_map.cpu().numpy()
    return [set(zip(imap[kpos[k]:kpos[k + 1]].tolist(), omap[kpos[k]:kpos[k + 1]].tolist())) for k in range(kpos.size - 1)]


@pytest.mark.parametrize("ks", [3, 5])
def test_kmap_submanifold_same_as_reference(REF, ks):
    """separate_mid = True is the reference's submanifold branch: out = in, centred taps, centre tap skipped."""
    from dgsparse.sparse_mapping import build_kernel_map
    in_c = random_coords(np.random.default_rng(ks), 6000, 2, 40)
    out_c, pairs, knnz, kpos, qkpos = reference_map(REF, in_c, ks, 1, 0, 0, 1000, True)
    km = build_kernel_map(dev(in_c), ks, 1, separate_mid=True)
    assert np.array_equal(out_c, in_c) and np.array_equal(km.out_coords.cpu().numpy(), in_c)
    assert np.array_equal(km.knnz.cpu().numpy(), knnz)
    assert np.array_equal(km.kpos.cpu().numpy(), kpos) and np.array_equal(km.qkpos.cpu().numpy(), qkpos)
    assert our_pairs(km) == pairs
    assert knnz[ks ** 3 // 2] == 0 and knnz.sum() > 0


@pytest.mark.parametrize("ks,stride,pad", [(3, 2, 1), (3, 2, 0), (5, 3, 2), (3, 1, 1), (2, 1, 0)])
def test_kmap_strided_same_as_reference(REF, ks, stride, pad):
    """General strided layers (coordsDownsampleExpand) and the stride-1 `_queryhash_sp` branch with padding."""
    from dgsparse.sparse_mapping import build_kernel_map
    in_c = random_coords(np.random.default_rng(10 * ks + stride), 5000, 2, 36)
    lo, hi = 0, 15 if stride > 1 else 1000
    out_c, pairs, knnz, kpos, qkpos = reference_map(REF, in_c, ks, stride, pad, lo, hi, False)
    if stride == 1:
        # reference: l_stride == 1 takes coordsDownsample with stride 1 (out = sorted unique in) and then queries with
        # `_queryhash_sp`; ours: the same output voxels handed to the strided query of dgs_kmap_build_ex
        import dgsparse.sparse_mapping as SM
        km = _build_sp(SM, dev(in_c), dev(np.unique(in_c, axis=0)), ks, stride, pad)
    else:
        km = build_kernel_map(dev(in_c), ks, stride, padding=pad, min_coord=lo, max_coord=hi)
    assert np.array_equal(km.out_coords.cpu().numpy(), out_c)
    assert np.array_equal(km.knnz.cpu().numpy(), knnz)
    assert np.array_equal(km.kpos.cpu().numpy(), kpos) and np.array_equal(km.qkpos.cpu().numpy(), qkpos)
    assert our_pairs(km) == pairs
    assert knnz.sum() > 0


def _build_sp(SM, c, out_coords, ks, stride, pad):
    """dgs_kmap_build_ex with subm = 0 on caller-supplied output voxels (the C ABI the Python wrapper sits on)."""
    from dgsparse._lib import check as ck, lib, ptr, stream_of
    n_in, n_out, k_vol = c.size(0), out_coords.size(0), ks ** 3
    imap = torch.empty(k_vol * n_out, dtype=torch.int32, device="cuda")
    omap = torch.empty_like(imap)
    knnz = torch.zeros(k_vol, dtype=torch.int32, device="cuda")
    kpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
    qkpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.dgs_kmap_workspace_bytes(n_in, n_out, k_vol), dtype=torch.uint8, device="cuda")
    ck(lib.dgs_kmap_build_ex(n_in, ptr(c), n_out, ptr(out_coords), ks, ks, ks, stride, stride, stride, pad, pad, pad, 0, 128, 0,
                             ptr(imap), ptr(omap), ptr(knnz), ptr(kpos), ptr(qkpos), ptr(ws), ws.numel(), stream_of(c)),
       "dgs_kmap_build_ex")
    pairs = int(kpos[-1].item())
    return SM.KernelMap(out_coords, imap[:pairs], omap[:pairs], knnz, kpos, qkpos, n_out, int(qkpos[-1].item()), False)


def test_kmap_plain_downsample_voxels_same_as_reference(REF):
    """ks = stride = 2: only the voxel SET is comparable (the reference keeps input resolution: x / s * s)."""
    from dgsparse.sparse_mapping import downsample_coords
    in_c = random_coords(np.random.default_rng(3), 8000, 3, 50)
    out_c, _, _, _, _ = reference_map(REF, in_c, 2, 2, 0, 0, 1000, False)
    ours = downsample_coords(dev(in_c), 2).cpu().numpy()
    ours[:, 1:] *= 2
    assert np.array_equal(ours, out_c)


def ref_forward(REF, x, w, kpos, qkpos, imap, omap, out_nnz, sum_nnz, separate_mid, arch80):
    """The reference writes into torch::empty (src/cuda/spconv_cuda.cu:30-32): hand it a zeroed allocator block."""
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    z = torch.zeros((out_nnz, w.shape[2]), dtype=torch.float32, device="cuda")
    del z
    out = REF.spconv_fwd_fused(x, w, kpos, qkpos, imap, omap, out_nnz, sum_nnz, separate_mid, arch80)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("arch80", [False, True])
@pytest.mark.parametrize("idx", [0, 1])
def test_forward_same_as_reference_on_minkunet_fixture(REF, idx, arch80):
    """The reference's own MinkUNet kernel maps (c_in = 4 and 64), driven as test/test_spconv.py:100-147 drives them."""
    import dgsparse.spconv as S
    g = np.load(os.path.join(GOLDEN, f"spconv_fp32_{idx}.npz"))
    in_nnz, out_nnz, k_vol, c_in, c_out = (int(g[k]) for k in ("in_nnz", "out_nnz", "k_vol", "c_in", "c_out"))
    rng = np.random.default_rng(40 + idx)
    x = rng.uniform(-1, 1, (in_nnz, c_in)).astype(np.float32)
    w = rng.uniform(-1, 1, (k_vol, c_in, c_out)).astype(np.float32)
    kpos, qkpos, sum_nnz = S.quantize_kpos(dev(g["knnz"]))
    dx, dw, di, do = dev(x), dev(w), dev(g["imap"]), dev(g["omap"])
    theirs = ref_forward(REF, dx, dw, kpos, qkpos, di, do, out_nnz, sum_nnz, False, arch80).cpu().numpy()
    ours = S.spconv_fwd_fused(dx, dw, kpos, qkpos, di, do, out_nnz, sum_nnz, False, arch80).cpu().numpy()
    want, bound = ref64(g["kpos"], g["imap"], g["omap"], x, w, out_nnz)
    # the reference's small-channel kernel (c_in <= 16 AND c_out <= 16) is plain fp32 whatever arch80 says (spconv_cuda.cu:136-144)
    prec_ref = "tf32" if (arch80 and not (c_in <= 16 and c_out <= 16)) else "fp32"
    prec = "tf32" if arch80 else "fp32"
    check(theirs, want, bound, prec_ref, f"reference fixture {idx}")
    check(ours, want, bound, prec, f"ours fixture {idx}")
    err = np.abs(ours.astype(np.float64) - theirs)
    assert (err <= 2 * TOL[prec] * bound + 1e-30).all(), float((err / (bound + 1e-30)).max())


def test_forward_matches_cpu_compute_golden():
    """Our CUDA (exact fp32 path) against golden OUTPUTS of the reference's cpu_compute (tests/golden/make_fixtures.py):
    all taps in the maps, and the separate_mid form (centre tap out of the maps, added from the identity)."""
    import dgsparse.spconv as S
    g = np.load(os.path.join(GOLDEN, "spconv_cpu_compute.npz"))
    out_size = int(g["out_size"])
    x, w = g["feats"], g["W"]
    assert x.shape[0] == out_size
    kp_full = np.concatenate([[0], np.cumsum(g["knnz"])])
    _, bound = ref64(kp_full, g["imap"], g["omap"], x, w, out_size)
    for knnz, imap, omap, sep, want in ((g["knnz"], g["imap"], g["omap"], False, g["out_full"]),
                                        (g["knnz_mid"], g["imap_mid"], g["omap_mid"], True, g["out_separate_mid"])):
        kpos, qkpos, sum_nnz = S.quantize_kpos(dev(knnz))
        got = S.spconv_fwd_fused(dev(x), dev(w), kpos, qkpos, dev(imap), dev(omap), out_size, sum_nnz, sep, False).cpu().numpy()
        assert (np.abs(got.astype(np.float64) - want) <= 1e-5 * bound + 1e-30).all(), sep


@pytest.mark.parametrize("arch80", [True])
def test_backward_same_as_reference(REF, arch80):
    """spconv_bwd_fused (src/cuda/spconv_cuda.cu:189-253): dX and dW, tf32 tensor-core kernels on both sides."""
    import dgsparse.spconv as S
    g = np.load(os.path.join(GOLDEN, "spconv_fp32_1.npz"))
    in_nnz, out_nnz, k_vol, c_in, c_out = (int(g[k]) for k in ("in_nnz", "out_nnz", "k_vol", "c_in", "c_out"))
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, (in_nnz, c_in)).astype(np.float32)
    w = rng.uniform(-1, 1, (k_vol, c_in, c_out)).astype(np.float32)
    go = rng.uniform(-1, 1, (out_nnz, c_out)).astype(np.float32)
    kpos, qkpos, sum_nnz = S.quantize_kpos(dev(g["knnz"]))
    dx, dw, dg, di, do = dev(x), dev(w), dev(go), dev(g["imap"]), dev(g["omap"])
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    z1, z2 = torch.zeros((in_nnz, c_in), device="cuda"), torch.zeros((k_vol, c_in, c_out), device="cuda")
    del z1, z2
    t_gin, t_gw = REF.spconv_bwd_fused(dg, dx, dw, kpos, qkpos, di, do, sum_nnz, False, arch80)
    torch.cuda.synchronize()
    o_gin, o_gw = S.spconv_bwd_fused(dg, dx, dw, kpos, qkpos, di, do, sum_nnz, False, arch80)
    torch.cuda.synchronize()
    # fp64 anchors: dX = sum_k scatter(in_map <- out_grad[out_map] @ W[k]^T); dW[k] = in[imap]^T @ out_grad[omap]
    wt = np.ascontiguousarray(np.transpose(w, (0, 2, 1)))
    want_in, bound_in = ref64(g["kpos"], g["omap"], g["imap"], go, wt, in_nnz)
    check(o_gin.cpu().numpy(), want_in, bound_in, "tf32", "ours dX")
    kp = g["kpos"]
    want_w = np.zeros((k_vol, c_in, c_out))
    bound_w = np.zeros_like(want_w)
    for k in range(k_vol):
        s, e = int(kp[k]), int(kp[k + 1])
        a, b = x[g["imap"][s:e]].astype(np.float64), go[g["omap"][s:e]].astype(np.float64)
        want_w[k] = a.T @ b
        bound_w[k] = np.abs(a).T @ np.abs(b)
    check(o_gw.cpu().numpy(), want_w, bound_w, "tf32", "ours dW")
    # The reference's backward has no test of its own (test/test_spconv.py never calls it).  Measured here on B200, built
    # unmodified: its dX (_fgms_fusion_tf32_W_transpose, include/cuda/spconv.cuh:1876+) does NOT agree with fp64 on this
    # 64 -> 64 layer (about half of the elements off by O(1) of their bound).  Where the reference is right ours must agree
    # with it; where it is wrong the finding is printed (pytest -s) and recorded in DESIGN.md §8, not mirrored.
    report = {}
    for name, theirs, ours_t, want, bound in (("dX", t_gin, o_gin, want_in, bound_in), ("dW", t_gw, o_gw, want_w, bound_w)):
        err = np.abs(theirs.cpu().numpy().astype(np.float64) - want)
        frac_bad = float((err > TOL["tf32"] * bound + 1e-30).mean())
        report[name] = frac_bad
        if frac_bad == 0.0:
            d = np.abs(ours_t.cpu().numpy().astype(np.float64) - theirs.cpu().numpy())
            assert (d <= 2 * TOL["tf32"] * bound + 1e-30).all(), name
    print("reference backward, fraction of elements outside the tf32 bound of the fp64 oracle:", report)
