"""GPU: kernel-map construction (dgs_kmap_*) against a plain-Python / numpy restatement of the reference's query rules
(_queryhash_subm / _queryhash_sp with padding 0, include/cuda/sparse_mapping.cuh:68-229; coordsDownsample + sort + unique,
src/cuda/sparse_mapping.cu:68-97).  Integer work: every output must be bit-exact.  The reference has no test or fixture
for this step; the comparison with the reference's OWN sparse_mapping CUDA (compiled unmodified) lives in
tests/test_vs_reference_spconv_gpu.py, and the end-to-end check below runs the maps through spconv against a dense convolution."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def random_coords(rng, n, batch, extent):
    c = np.stack([rng.integers(0, batch, n), rng.integers(0, extent, n), rng.integers(0, extent, n), rng.integers(0, extent, n)], 1)
    return np.unique(c, axis=0).astype(np.int32)          # voxel coordinates are unique


def tap_off(k, ks):
    """Tap decode of the strided kernels, include/cuda/sparse_mapping.cuh:195-200 and :365-370."""
    return k - (ks - 1) // 2 + (0 if (ks % 2 == 0 or ks == 1) else 1)


def ref_kernel_map(in_c, ks, st, skip_mid=False, pad=None, lo=None, hi=None):
    """Restatement with a dict: returns (out_coords, imap, omap, knnz)."""
    table = {tuple(c): i for i, c in reversed(list(enumerate(in_c.tolist())))}
    sub = st == (1, 1, 1) and pad is None
    pd = pad or (0, 0, 0)
    plain = pd == (0, 0, 0) and lo is None and hi is None and all(s in (1, k) for s, k in zip(st, ks))
    if sub:
        out_c = in_c
    elif not plain:
        # coordsDownsampleExpand, include/cuda/sparse_mapping.cuh:326-401 (C division: exact multiples only)
        cand = set()
        for b, x, y, z in in_c.tolist():
            for kx in range(ks[0]):
                cx = x - tap_off(kx, ks[0]) + pd[0]
                if cx % st[0]:
                    continue
                for ky in range(ks[1]):
                    cy = y - tap_off(ky, ks[1]) + pd[1]
                    if cy % st[1]:
                        continue
                    for kz in range(ks[2]):
                        cz = z - tap_off(kz, ks[2]) + pd[2]
                        if cz % st[2]:
                            continue
                        o = (cx // st[0], cy // st[1], cz // st[2])
                        if lo is not None and any(o[d] < lo[d] for d in range(3)):
                            continue
                        if hi is not None and any(o[d] > hi[d] for d in range(3)):
                            continue
                        cand.add((b,) + o)
        out_c = np.array(sorted(cand), dtype=np.int32).reshape(-1, 4)
    else:
        d = in_c.copy()
        d[:, 1] //= st[0]; d[:, 2] //= st[1]; d[:, 3] //= st[2]
        out_c = np.unique(d, axis=0)                        # lexicographic (batch, x, y, z)
    k_vol = ks[0] * ks[1] * ks[2]
    mid = k_vol // 2 if k_vol % 2 == 1 else 0
    imap, omap, knnz = [], [], []
    for k in range(k_vol):
        kx, ky, kz = k // (ks[2] * ks[1]), (k // ks[2]) % ks[1], k % ks[2]
        cnt = 0
        if not (skip_mid and k == mid):
            for o, (b, x, y, z) in enumerate(out_c.tolist()):
                if sub:
                    key = (b, x + kx - (ks[0] - 1) // 2, y + ky - (ks[1] - 1) // 2, z + kz - (ks[2] - 1) // 2)
                else:
                    key = (b, x * st[0] - pd[0] + tap_off(kx, ks[0]), y * st[1] - pd[1] + tap_off(ky, ks[1]),
                           z * st[2] - pd[2] + tap_off(kz, ks[2]))
                i = table.get(key)
                if i is not None:
                    imap.append(i); omap.append(o); cnt += 1
        knnz.append(cnt)
    return out_c, np.array(imap, np.int32), np.array(omap, np.int32), np.array(knnz, np.int32)


@pytest.mark.parametrize("ks,st,skip_mid", [(3, 1, False), (3, 1, True), (2, 2, False), (3, 2, False), ((1, 3, 3), 1, False),
                                            (5, 1, False)])
def test_kernel_map_bit_exact(ks, st, skip_mid):
    from dgsparse.sparse_mapping import build_kernel_map, _triple
    rng = np.random.default_rng(3)
    in_c = random_coords(rng, 6000, 2, 24)
    km = build_kernel_map(torch.from_numpy(in_c).cuda(), ks, st, separate_mid=skip_mid)
    out_c, imap, omap, knnz = ref_kernel_map(in_c, _triple(ks), _triple(st), skip_mid)
    assert np.array_equal(km.out_coords.cpu().numpy(), out_c)
    assert np.array_equal(km.knnz.cpu().numpy(), knnz)
    kpos = np.concatenate([[0], np.cumsum(knnz)]).astype(np.int32)
    assert np.array_equal(km.kpos.cpu().numpy(), kpos)
    q = np.concatenate([[0], np.cumsum((knnz + 127) // 128 * 128)]).astype(np.int32)
    assert np.array_equal(km.qkpos.cpu().numpy(), q) and km.sum_nnz == int(q[-1])
    assert np.array_equal(km.in_map.cpu().numpy(), imap) and np.array_equal(km.out_map.cpu().numpy(), omap)


@pytest.mark.parametrize("ks,st,pad,lo,hi", [(3, 2, 1, None, None), (3, 2, 1, 0, 11), (3, 2, 0, None, None), (2, 2, 1, None, None),
                                             ((3, 3, 2), (2, 1, 2), (1, 1, 0), None, None), (5, 2, 2, (0, 0, 0), (9, 10, 11)),
                                             (4, 3, 1, None, None), (3, 1, 1, None, None)])
def test_expand_branch_bit_exact(ks, st, pad, lo, hi):
    """General strided layers (stride neither 1 nor the kernel size, padding, bounds): output voxels and pair lists."""
    from dgsparse.sparse_mapping import build_kernel_map, _triple
    rng = np.random.default_rng(11)
    in_c = random_coords(rng, 3000, 2, 24)
    in_c[:, 1:] -= 3                                       # a few negative coordinates: exact-multiple rule below zero
    in_c = np.unique(in_c, axis=0).astype(np.int32)
    lo3, hi3 = (_triple(lo) if lo is not None else None), (_triple(hi) if hi is not None else None)
    km = build_kernel_map(torch.from_numpy(in_c).cuda(), ks, st, padding=pad, min_coord=lo, max_coord=hi)
    out_c, imap, omap, knnz = ref_kernel_map(in_c, _triple(ks), _triple(st), False, _triple(pad), lo3, hi3)
    assert np.array_equal(km.out_coords.cpu().numpy(), out_c)
    assert np.array_equal(km.knnz.cpu().numpy(), knnz)
    assert np.array_equal(km.in_map.cpu().numpy(), imap) and np.array_equal(km.out_map.cpu().numpy(), omap)
    q = np.concatenate([[0], np.cumsum((knnz + 127) // 128 * 128)]).astype(np.int32)
    assert np.array_equal(km.qkpos.cpu().numpy(), q) and km.sum_nnz == int(q[-1])
    # every input reaches at least one output unless the bounds cut it off
    if lo is None and pad != 0:
        assert np.unique(km.in_map.cpu().numpy()).size == in_c.shape[0]


def test_expand_empty_input():
    from dgsparse.sparse_mapping import build_kernel_map
    none = torch.zeros((0, 4), dtype=torch.int32, device="cuda")
    km = build_kernel_map(none, 3, 2, padding=1)
    assert km.out_nnz == 0 and km.in_map.numel() == 0 and km.sum_nnz == 0


def test_edge_cases():
    from dgsparse.sparse_mapping import build_kernel_map, downsample_coords
    one = torch.tensor([[0, 5, 5, 5]], dtype=torch.int32, device="cuda")
    km = build_kernel_map(one, 3, 1)
    assert km.knnz.cpu().tolist() == [0] * 13 + [1] + [0] * 13 and km.in_map.cpu().tolist() == [0]
    neg = torch.tensor([[0, -3, 0, 7], [0, -4, 1, 6], [1, -3, 0, 7]], dtype=torch.int32, device="cuda")
    assert downsample_coords(neg, 2).cpu().tolist() == [[0, -2, 0, 3], [1, -2, 0, 3]]     # floor division, batch kept apart
    with pytest.raises(TypeError):
        build_kernel_map(one.long(), 3, 1)
    with pytest.raises(RuntimeError):
        build_kernel_map(one.cpu(), 3, 1)


def test_maps_drive_spconv_like_a_dense_convolution():
    """kernel map -> spconv == dense conv3d evaluated at the occupied voxels (submanifold and stride-2 layers)."""
    import torch.nn.functional as F
    from dgsparse.sparse_mapping import build_kernel_map
    rng = np.random.default_rng(7)
    E, c_in, c_out = 12, 8, 16
    in_c = random_coords(rng, 500, 1, E)
    n = in_c.shape[0]
    feats = torch.tensor(rng.uniform(-1, 1, (n, c_in)), dtype=torch.float32, device="cuda")
    dense = torch.zeros(1, c_in, E, E, E, device="cuda")
    ic = torch.from_numpy(in_c).cuda().long()
    dense[0, :, ic[:, 1], ic[:, 2], ic[:, 3]] = feats.T
    for ks, st, pad in ((3, 1, None), (2, 2, None), (3, 2, 1), (3, 1, 1), (2, 1, 0)):
        W = torch.tensor(rng.uniform(-1, 1, (ks ** 3, c_in, c_out)), dtype=torch.float32, device="cuda")
        if pad is not None:
            # general layer (expand branch): the dense convolution's full output grid bounds the voxel set, and every
            # grid point the maps leave out must be exactly zero in the dense result
            n_o = (E + 2 * pad - ks) // st + 1
            km = build_kernel_map(torch.from_numpy(in_c).cuda(), ks, st, padding=pad, min_coord=0, max_coord=n_o - 1)
            out = torch.ops.dgsparse_spconv.spconv(feats, W, km.kpos, km.qkpos, km.in_map, km.out_map, km.out_nnz,
                                                   km.sum_nnz, False, False)
            w5 = W.reshape(ks, ks, ks, c_in, c_out).permute(4, 3, 0, 1, 2).contiguous()
            ref = F.conv3d(dense.double(), w5.double(), stride=st, padding=pad)
            oc = km.out_coords.long()
            want = ref[0][:, oc[:, 1], oc[:, 2], oc[:, 3]].T
            assert torch.allclose(out.double(), want, rtol=1e-5, atol=1e-5), (ks, st, pad)
            covered = torch.zeros_like(ref[0][0], dtype=torch.bool)
            covered[oc[:, 1], oc[:, 2], oc[:, 3]] = True
            assert float(ref[0][:, ~covered].abs().max() if (~covered).any() else 0.0) < 1e-12   # cuDNN fp64 rounding noise
            continue
        km = build_kernel_map(torch.from_numpy(in_c).cuda(), ks, st)
        out = torch.ops.dgsparse_spconv.spconv(feats, W, km.kpos, km.qkpos, km.in_map, km.out_map, km.out_nnz, km.sum_nnz,
                                               False, False)                       # exact fp32 path
        w5 = W.reshape(ks, ks, ks, c_in, c_out).permute(4, 3, 0, 1, 2).contiguous()   # [c_out, c_in, kx, ky, kz]
        ref = F.conv3d(dense.double(), w5.double(), stride=st, padding=(ks - 1) // 2 if st == 1 else 0)
        oc = km.out_coords.long()
        want = ref[0][:, oc[:, 1], oc[:, 2], oc[:, 3]].T
        assert torch.allclose(out.double(), want, rtol=1e-5, atol=1e-5), (ks, st)


def test_out_of_range_coordinates_are_rejected():
    """a voxel key holds 16 bits per component; coordinates outside (-32768, 32768) or batch indices above 65535
    used to alias other voxels silently.  The Python face raises, the C ABI poisons knnz / kpos / qkpos with -1."""
    from dgsparse.sparse_mapping import build_kernel_map, downsample_coords
    from dgsparse._lib import check, lib, ptr, stream_of
    good = torch.tensor([[0, 1, 2, 3], [0, 2, 2, 3], [1, 5, 5, 5]], dtype=torch.int32, device="cuda")
    for bad_row in ([0, 40000, 0, 0], [0, 0, -32768, 0], [70000, 1, 1, 1], [-1, 1, 1, 1]):
        c = torch.cat([good, torch.tensor([bad_row], dtype=torch.int32, device="cuda")])
        with pytest.raises(ValueError):
            build_kernel_map(c, 3, 1)
        with pytest.raises(ValueError):
            downsample_coords(c, 2)
        n, k_vol = c.size(0), 27
        imap = torch.empty(k_vol * n, dtype=torch.int32, device="cuda")
        omap = torch.empty_like(imap)
        knnz = torch.zeros(k_vol, dtype=torch.int32, device="cuda")
        kpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
        qkpos = torch.zeros(k_vol + 1, dtype=torch.int32, device="cuda")
        ws = torch.empty(lib.dgs_kmap_workspace_bytes(n, n, k_vol), dtype=torch.uint8, device="cuda")
        check(lib.dgs_kmap_build_ex(n, ptr(c), n, ptr(c), 3, 3, 3, 1, 1, 1, 0, 0, 0, 1, 128, 0, ptr(imap), ptr(omap), ptr(knnz),
                                    ptr(kpos), ptr(qkpos), ptr(ws), ws.numel(), stream_of(c)), "dgs_kmap_build_ex")
        assert int(kpos[-1]) == -1 and int(qkpos[-1]) == -1 and bool((knnz == -1).all())
    km = build_kernel_map(good, 3, 1)          # the in-range part alone is fine
    assert km.sum_nnz >= 0 and int(km.kpos[-1]) >= good.size(0)
