"""Shared helpers for the parity tests: seeded inputs and tolerances.

Tolerance (BASELINE north_star "within 1e-5 relative"):
  * vs the serial-order fp32 oracle:  |ours - oracle| <= 1e-5 * |oracle| + ATOL, where ATOL scales with
    the magnitude of the terms (fp32 summation-order noise; the GPU folds a row in segments);
  * vs the fp64 oracle: max relative error <= 1e-5 on rows of any length;
  * arg index E and every integer output: exact.
"""
import numpy as np

RTOL = 1e-5


def assert_close_f32(ours, ref, ref64=None, what="", scale=None):
    ours = np.asarray(ours, np.float64)
    ref = np.asarray(ref, np.float64)
    assert ours.shape == ref.shape, (what, ours.shape, ref.shape)
    anchor = ref if ref64 is None else np.asarray(ref64, np.float64)
    if scale is None:
        scale = max(1.0, float(np.abs(anchor).max()) if anchor.size else 1.0)
    # the serial fp32 oracle itself carries ~deg * 2^-24 relative error on long rows: judge against
    # the fp64 anchor when given, with the stated 1e-5 relative bound plus 1e-6 of the output scale
    err = np.abs(ours - anchor)
    bound = RTOL * np.abs(anchor) + 1e-6 * scale
    bad = err > bound
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.size} outside tolerance, max err {err.max():.3e}, " \
                          f"worst at {np.unravel_index(np.argmax(err - bound), err.shape)}"


def seeded(graphs, n, seed, lo=0.0, hi=1.0):
    return graphs.uniform(n, seed, lo, hi)
