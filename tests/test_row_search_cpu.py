"""Host-side check of the G-ary row search the SDDMM kernels run at every chunk start (`row_of_nnz_group<G>`,
dgsparse-lib_b200/csrc/sddmm.cu): a literal restatement of its rounds (G evenly spaced probes, first probe beyond p, shrink) against
numpy.searchsorted, i.e. against the contract of the reference's findRow / binary_search_segment_number
(src/util/cuda_util.cuh:53-89, 167-174): the row r with rowptr[r] <= p < rowptr[r+1], empty rows skipped.  The CUDA code
itself is covered by the -m gpu parity tests; this pins the arithmetic (probe positions, range updates, termination)."""
import numpy as np
import pytest


def row_of_nnz_group(rowptr, M, p, G):
    LG = {32: 5, 16: 4, 8: 3, 4: 2}[G]
    lo, hi, loads = 0, M, 0
    while hi - lo >= G:
        span = hi - lo
        beyond = [rowptr[lo + ((span * (gl + 1)) >> LG)] > p for gl in range(G)]      # one load per lane, one ballot
        assert beyond[-1], "the last probe is hi itself and rowptr[hi] > p is the loop invariant"
        f = beyond.index(True)
        q_prev = lo + ((span * f) >> LG)
        new_hi = lo + ((span * (f + 1)) >> LG)
        lo = lo if f == 0 else q_prev + 1
        assert new_hi - lo < span, "every round must shrink the range"
        hi = new_hi
        loads += 1
    beyond = [(lo + gl >= hi) or rowptr[lo + gl] > p for gl in range(G)]
    return lo + beyond.index(True) - 1, loads + 1


@pytest.mark.parametrize("G", [4, 8, 16, 32])
def test_group_row_search_matches_searchsorted(G):
    rng = np.random.default_rng(G)
    for trial in range(400):
        M = int(rng.integers(1, 40)) if trial % 3 == 0 else int(rng.integers(1, 6000))
        deg = rng.integers(0, 4, M) * (rng.random(M) < 0.6)          # many empty rows
        if trial % 5 == 0:
            deg[rng.integers(0, M)] += 700                           # a hub row
        rowptr = np.zeros(M + 1, np.int64)
        rowptr[1:] = np.cumsum(deg)
        nnz = int(rowptr[-1])
        if nnz == 0:
            continue
        for p in [0, nnz - 1] + [int(x) for x in rng.integers(0, nnz, 6)]:
            got, _ = row_of_nnz_group(rowptr, M, p, G)
            assert got == int(np.searchsorted(rowptr, p, side="right")) - 1, (M, G, p)


def test_group_row_search_depth():
    """What it is for: 4 dependent loads for a full warp on an arxiv-sized row pointer where a bisection needs 18."""
    rng = np.random.default_rng(0)
    M = 169343
    rowptr = np.zeros(M + 1, np.int64)
    rowptr[1:] = np.cumsum(rng.integers(0, 14, M))
    depth = {G: max(row_of_nnz_group(rowptr, M, int(p), G)[1] for p in rng.integers(0, rowptr[-1], 50)) for G in (8, 16, 32)}
    assert depth[32] <= 4 and depth[16] <= 5 and depth[8] <= 6, depth
