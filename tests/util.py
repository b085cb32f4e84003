"""Shared helpers for the parity tests: seeded inputs and tolerances.

Tolerance (BASELINE north_star "within 1e-5 relative"), per ELEMENT, against the fp64 oracle when one is
given (the serial fp32 oracle itself carries ~deg * 2^-24 relative error on long rows), else against the fp32 oracle:

    |ours - anchor| <= 1e-5 * |anchor| + 1e-6 * T

  * T = 1 by default: the fixed absolute term of SURVEY §8c (inputs of magnitude O(1));
  * T = `absref`, the element's own sum of |terms| (the same reduction over |val|, |B| in fp64), wherever inputs are
    signed or rows are long: cancellation makes |anchor| arbitrarily small while the rounding noise of a sum scales with
    the magnitude of its terms.  For non-negative inputs absref == |anchor| and the bound is 1.1e-5 relative;
  * arg index E, max / min outputs and every integer output: exact (asserted with array_equal at the call sites).
The absolute term never scales with the largest output of the matrix.
"""
import numpy as np

RTOL = 1e-5
ATOL_UNIT = 1e-6


def assert_close_f32(ours, ref, ref64=None, what="", absref=None, scale=None):
    ours = np.asarray(ours, np.float64)
    ref = np.asarray(ref, np.float64)
    assert ours.shape == ref.shape, (what, ours.shape, ref.shape)
    anchor = ref if ref64 is None else np.asarray(ref64, np.float64)
    if absref is not None:
        T = np.asarray(absref, np.float64)
        assert T.shape == anchor.shape, (what, T.shape, anchor.shape)
    else:
        T = 1.0 if scale is None else float(scale)
    err = np.abs(ours - anchor)
    bound = RTOL * np.abs(anchor) + ATOL_UNIT * T
    bad = err > bound
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.size} outside tolerance, max err {err.max():.3e}, " \
                          f"worst at {np.unravel_index(np.argmax(err - bound), err.shape)}"


def spmm_absref(oracle, rowptr, col, val, B, reduce="sum", compute="mul"):
    """Per-element sum of |terms| of a (generalized) SpMM in fp64: |b - a| <= |a| + |b|, |b / a| = |b| / |a|."""
    av = None if val is None else np.abs(val)
    comp = "add" if compute == "sub" else compute
    return oracle.spmm_f64(rowptr, col, av, np.abs(B), "mean" if reduce == "mean" else "sum", comp)


def sddmm_absref(oracle, rowptr, col, D1, D2, mean=False):
    return oracle.sddmm_csr(rowptr, col, np.abs(D1), np.abs(D2), mean, f64=True)


def seeded(graphs, n, seed, lo=0.0, hi=1.0):
    return graphs.uniform(n, seed, lo, hi)
