"""GPU drop-in proof of the dgsparse.h C ABI: the REFERENCE'S OWN self-checking example drivers — example/ge-spmm/spmm.cu
(every gespmmCsrSpMM algorithm, :171-216) and example/sddmm/sddmm.cu (sddmm_cuda_csr, :167-195) — compiled UNMODIFIED
where they lie under /root/reference and linked against OUR libdgsparse_b200.so instead of the reference's libgespmm.a /
libsddmm.a (oracle/Makefile target `refdrivers`; the binaries travel in oracle/_ref/).  Each driver compares every output
element with spmm_reference_host / sddmm_reference_host (example/util/sp_util.hpp:62-131) and prints its "Report" line
only when that check passed.  Skipped when the binaries were not built (reference tree absent at build time)."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def write_mtx(path, rowptr, col, shape):
    """The fixture CSR as a MatrixMarket `coordinate pattern general` file (what read_mtx_file, sp_util.hpp:171-248, reads)."""
    rows = np.repeat(np.arange(shape[0]), np.diff(rowptr))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate pattern general\n")
        f.write(f"{shape[0]} {shape[1]} {col.size}\n")
        np.savetxt(f, np.stack([rows + 1, col.astype(np.int64) + 1], 1), fmt="%d")


def run_driver(name, mtx, width):
    exe = os.path.join(REF_DIR, name)
    if not os.path.exists(exe):
        pytest.skip(f"oracle/_ref/{name} not built (make -C oracle refdrivers needs /root/reference)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = "/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([exe, mtx, str(width)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.fixture(scope="module")
def gnutella_mtx(tmp_path_factory, graphs):
    rowptr, col, shape = graphs.load_fixture("p2p-Gnutella31")
    p = str(tmp_path_factory.mktemp("mtx") / "p2p-Gnutella31.mtx")
    write_mtx(p, rowptr, col, shape)
    return p, shape, int(col.size)


def report_times(out, tag):
    """{label: ms} from the drivers' 'Report' blocks: '[tag...] Report: ...\n Time X (ms), Throughput Y (gflops).'"""
    t = {}
    lines = out.splitlines()
    for i, l in enumerate(lines):
        if l.startswith(tag) and "Report" in l and i + 1 < len(lines):
            m = re.search(r"Time ([0-9.eE+-]+) \(ms\)", lines[i + 1])
            if m:
                t[l.split(" Report")[0]] = float(m.group(1))
    return t


@pytest.mark.parametrize("N", [32])       # configs[0] of BASELINE.json (feat = 32); one width keeps the run short: a driver run is ~45 s
def test_reference_spmm_example_passes_on_our_library(gnutella_mtx, N):
    mtx, shape, nnz = gnutella_mtx
    out = run_driver("spmm_example.out", mtx, N)
    assert f"Finish reading matrix {shape[0]} rows, {shape[1]} columns, {nnz} nnz" in out
    assert "Wrong result" not in out
    # six algorithms (example/ge-spmm/spmm.cu:172-175), each reported only after its element-wise check passed
    reports = [l for l in out.splitlines() if l.startswith("[GE-SpMM][Alg:")]
    assert len(reports) == 6, out[-3000:]
    # The driver's own timing loop (100 calls from C++, GpuTimer) is the reference's PUBLISHED measurement
    # (example/README.md:47-60: p2p-Gnutella31, N = 32).  Ours must beat the cuSPARSE call the driver times first ...
    ours = report_times(out, "[GE-SpMM][Alg:")
    cusparse = report_times(out, "[Cusparse]")
    assert len(ours) == 6 and len(cusparse) == 1, out[-3000:]
    t_cusparse = next(iter(cusparse.values()))
    assert max(ours.values()) < t_cusparse, (ours, t_cusparse)
    # ... and the reference's own library under the SAME driver on the same GPU (oracle/_ref/spmm_example_ref.out), for every
    # algorithm value (ours serves all six with one kernel).  The hard assertion carries 25 % of slack — these are ~10 us calls
    # and a timing assertion must not be what stops a parity run; the measured pair is printed (pytest -s) and recorded in
    # profiles/r02_reference_drivers.txt
    if os.path.exists(os.path.join(REF_DIR, "spmm_example_ref.out")):
        theirs = report_times(run_driver("spmm_example_ref.out", mtx, N), "[GE-SpMM][Alg:")
        print("driver-reported ms per call, ours vs the reference's library:", {k: (ours[k], theirs.get(k)) for k in ours})
        best_theirs = min(theirs.values())
        assert min(ours.values()) <= 1.25 * best_theirs, (ours, theirs)


@pytest.mark.parametrize("K", [64])
def test_reference_sddmm_example_passes_on_our_library(gnutella_mtx, K):
    mtx, shape, nnz = gnutella_mtx
    out = run_driver("sddmm_example.out", mtx, K)
    assert "Wrong result" not in out
    assert any(l.startswith("[SDDMM] Report") for l in out.splitlines()), out[-3000:]
    ours = report_times(out, "[SDDMM]")
    cusparse = report_times(out, "[cuSPARSE]")
    if ours and cusparse:
        assert next(iter(ours.values())) < next(iter(cusparse.values())), (ours, cusparse)
    if os.path.exists(os.path.join(REF_DIR, "sddmm_example_ref.out")):
        theirs = report_times(run_driver("sddmm_example_ref.out", mtx, K), "[SDDMM]")
        print("driver-reported ms per call, ours vs the reference's library:", ours, theirs)
