#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload reddit64|products128|arxiv256]

Workload (N=1): config 2 — SpMM sum on a reddit-like CSR (232 965 x 232 965, 114 615 892 nnz, synthetic,
SURVEY.md §8d), feat = 64, fp32, edge values present.  A step = one pass of the hot path (the row-segment
SpMM kernel + its fix-up kernel) over the whole matrix.  Metric: GFLOP/s = 2*nnz*N / t (the reference's
convention, example/ge-spmm/spmm.cu:213-215), with the achieved HBM GB/s (algorithmic bytes / t) beside it.

N>1 (torchrun, one rank per GPU): the feature axis is sharded — every rank holds the replicated CSR and a
64-column panel of B (total feat = 64*N; N=8 is config 5), computes its panel of C and makes it visible on
every rank (fused NVLS-multicast or peer-store epilogue over NVLink + a one-int NCCL barrier, or an NCCL allgather
when neither mapping is available).  Per-GPU work is fixed -> "weak".  value = 2*nnz*64*N / max-over-ranks time.

--impl reference: the reference's own CPU implementation of the path (spmm_reference_host from
oracle/_ref when it was built from /root/reference, else the oracle port) on the host cores, rank 0 only.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "dgsparse-lib_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

FEAT_PER_GPU = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reddit64", choices=["reddit64", "products128", "arxiv256"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic graph (debug only; default 1.0)")
    ap.add_argument("--mode", default=None, choices=[None, "mcast", "peer", "nccl"], help="N>1 exchange (default: mcast, falling back to peer, then nccl)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the configs 3 / 4 block appended to the default N=1 run")
    ap.add_argument("--no-legs", action="store_true", help="N>1: skip the compute-only / comm-only / strong-scaling legs")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload):
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu capture (profiles/traffic.json) — only when
    that capture was taken on the kernel sources this run uses (tools/csrc_hash.py); a capture older than the code reads
    as null instead of going stale silently.  -> (bytes or None, note)."""
    from tools import csrc_hash
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        ent = json.load(open(p))[workload]
    except Exception:
        return None, "no ncu capture recorded for this workload"
    now = csrc_hash.family_hash(ent.get("family", "spmm"))
    if ent.get("csrc_sha") != now:
        return None, f"capture {ent.get('source')} predates the current kernel sources ({ent.get('csrc_sha')} != {now})"
    return ent["traffic_bytes"], f"{ent.get('source')} ({ent.get('kernel')})"


def algorithmic_bytes_spmm(M, nnz, N, k_touched, has_value, with_arg=False):
    """rowptr + col + val + each referenced B row once + C (+ E for max/min)."""
    return 4 * (M + 1) + 4 * nnz + (4 * nnz if has_value else 0) + 4 * k_touched * N + 4 * M * N + (4 * M * N if with_arg else 0)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload_string(wl, M, nnz, N, world, has_value):
    """config.workload — the SAME string from both arms (ours and --impl reference) for the same workload."""
    return (f"{wl['op']} on {wl['name']}, M=K={M}, nnz={nnz}, feat={N} per GPU ({N * world} total), fp32, "
            f"edge values {'present' if has_value else 'absent'}")


def make_workload(name, scale):
    from tools import graphs
    graphs.build()
    if name == "reddit64":
        rowptr, col = graphs.reddit_like(scale)
        return dict(name="reddit-like synthetic CSR (lognormal degrees, uniform columns, seed 20240001)", rowptr=rowptr,
                    col=col, N=64, op="spmm_sum", has_value=True)
    if name == "products128":
        rowptr, col = graphs.products_like(scale)
        return dict(name="ogbn-products-like synthetic CSR (power-law degrees, seed 20240002)", rowptr=rowptr, col=col,
                    N=128, op="gspmm_u_mul_e_max", has_value=True)
    rowptr, col = graphs.arxiv_like(scale)
    return dict(name="ogbn-arxiv-like synthetic CSR (seed 20240003)", rowptr=rowptr, col=col, N=256, op="sddmm_csr",
                has_value=False)


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, steps, warmup, budget_s=12.0):
    """Times the reference's CPU path on a bounded sample.  -> (gflops, info dict)."""
    from oracle import oracle
    from tools import graphs
    oracle.build()
    rowptr, col, N = wl["rowptr"], wl["col"], wl["N"]
    M = rowptr.size - 1
    cores = len(os.sched_getaffinity(0))
    val = graphs.uniform(col.size, 1)
    B = graphs.uniform(M * N, 2).reshape(M, N)
    have_ref = oracle.ref_lib() is not None
    if have_ref:
        kind = "reference"

        def run(r0, r1):
            oracle.ref_spmm_host_threads(rowptr, col, val, B, threads=cores, rows=(r0, r1))
    else:
        kind = "port"
        oracle.set_num_threads(cores)

        def run(r0, r1):
            oracle.spmm(rowptr[r0:r1 + 1] - rowptr[r0], col[rowptr[r0]:rowptr[r1]], val[rowptr[r0]:rowptr[r1]], B)
    # probe on ~1/128 of the rows, then size the sample to the budget
    probe_rows = max(64, M // 128)
    run(0, probe_rows)
    t0 = time.perf_counter(); run(0, probe_rows); tp = time.perf_counter() - t0
    rate = int(rowptr[probe_rows]) / max(tp, 1e-6)                      # nnz / s
    total_steps = max(1, steps + warmup)
    sample_nnz = min(int(col.size), int(rate * budget_s / total_steps))
    r1 = int(np.searchsorted(rowptr, sample_nnz, side="right") - 1)
    r1 = min(max(r1, probe_rows), M)
    nnz_s = int(rowptr[r1])
    for _ in range(warmup):
        run(0, r1)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); run(0, r1); ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    gflops = 2.0 * nnz_s * N / t / 1e9
    info = {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": f"rows [0,{r1}) of the same CSR ({nnz_s} nnz = {100.0 * nnz_s / col.size:.1f}% of the workload), "
                      f"full B [{M},{N}], {steps} timed passes, row blocks spread over {cores} host threads",
            "ms_per_sample": t * 1e3}
    # what the reference's own tests call as their CPU oracle (test/test_spmm.py:60-61): torch.sparse.mm(csr_cpu, X_cpu),
    # on the same row sample, with torch's thread count stated
    try:
        import torch
        rp_s = torch.from_numpy(rowptr[:r1 + 1].astype(np.int64))
        A = torch.sparse_csr_tensor(rp_s, torch.from_numpy(col[:nnz_s].astype(np.int64)), torch.from_numpy(val[:nnz_s]),
                                    size=(r1, M))
        X = torch.from_numpy(B)
        torch.sparse.mm(A, X)
        t0 = time.perf_counter()
        for _ in range(max(1, steps)):
            torch.sparse.mm(A, X)
        tt = (time.perf_counter() - t0) / max(1, steps)
        info["torch_sparse_mm"] = {"value": 2.0 * nnz_s * N / tt / 1e9, "unit": "GFLOP/s", "threads": torch.get_num_threads(),
                                   "ms_per_sample": tt * 1e3, "what": "torch.sparse.mm(csr_cpu, X_cpu) on the same row sample"}
    except Exception as ex:  # informational only
        info["torch_sparse_mm"] = {"error": repr(ex)}
    return gflops, info


def reference_cuda_run(wl, M, N, nnz, rp, cc, vv, B, D1, D2, our_step, steps, warmup, flop):
    """Times the reference's CUDA C ABI (compiled unmodified for sm_100a into oracle/_ref/libref_cuda.so)
    on the same device buffers, and compares its output with ours.  Test/bench infrastructure only."""
    import torch
    from oracle import oracle
    R = oracle.ref_cuda_lib()
    if R is None:
        return {"unavailable": "oracle/_ref/libref_cuda.so was not built (reference tree absent at build time)"}
    dev = rp.device
    if wl["op"] == "sddmm_csr":
        ref_out = torch.zeros(nnz, device=dev)
        our_out = torch.empty(1, nnz, device=dev)
        import dgsparse._lib as L

        def ref_step():
            R.sddmm_cuda_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), D2.data_ptr(), ref_out.data_ptr())
        L.check(L.lib.dgs_sddmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), N, D2.data_ptr(), N,
                                    None, 0, our_out.data_ptr(), torch.cuda.current_stream().cuda_stream), "sddmm")
        what = "sddmm_cuda_csr (src/sddmm/sddmm.cu:25-41)"
    elif wl["op"] == "spmm_sum":
        ref_out = torch.zeros(M, N, device=dev)
        import dgsparse._lib as L
        import dgsparse._kernels as K
        our_out = K.spmm(rp, cc, vv, B, L.SUM, L.MUL)

        def ref_step():
            R.spmm_cuda(M, N, rp.data_ptr(), cc.data_ptr(), vv.data_ptr(), B.data_ptr(), ref_out.data_ptr())
        what = "spmm_cuda -> gespmmCsrSpMM (src/ge-spmm/gespmm.cc:29-123)"
    else:
        # gspmm-fp: the reference's own pybind module, built unmodified by oracle/build_ref_gspmm.sh
        G = oracle.ref_gspmm_module()
        if G is None:
            return {"unavailable": "oracle/_ref/spmm.so (reference gspmm-fp module) was not built"}
        import dgsparse._lib as L
        import dgsparse._kernels as K
        red = "MAX" if "max" in wl["op"] else "MEAN" if "mean" in wl["op"] else "SUM"
        our_out = K.spmm(rp, cc, vv, B, getattr(L, red), L.MUL)
        holder = {}

        def ref_step():
            holder["out"] = G.GSpMM_u_e(rp, cc, vv, B, getattr(G.REDUCEOP, red), G.COMPUTEOP.MUL)
        ref_step()
        ref_out = holder["out"]
        what = f"GSpMM_u_e(..., {red}, MUL) (src/gspmm-fp/gspmm.cu:406-473; allocates its output every call)"
    for _ in range(warmup):
        ref_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ref_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ref_out = locals().get("holder", {}).get("out", ref_out)
    a, b = our_out.reshape(-1), ref_out.reshape(-1)
    err = float(((a - b).abs() / b.abs().clamp_min(1e-6)).max().item())
    return {"kernel": what, "ms_per_step": ms, "value": flop / (ms * 1e-3) / 1e9, "unit": "GFLOP/s",
            "max_rel_diff_ours_vs_reference_cuda": err,
            "note": "device-resident, default stream, same buffers and step count as `value`"}


def sampled_rows_ok(rowptr, rp, cc, vv, B, C, reduce_name, n_rows=64):
    """fp64 recompute of a sample of rows (the 8 longest included) on the device: parity evidence inside the bench run.
    sum / mean: |C - ref| <= 1e-5 |ref| + 1e-6 * sum|terms|;  max: exact (fp32 products)."""
    import torch
    M = rowptr.size - 1
    rows = np.unique(np.concatenate([np.linspace(0, M - 1, n_rows).astype(np.int64), np.argsort(np.diff(rowptr))[-8:]]))
    for r in rows:
        s, e = int(rowptr[r]), int(rowptr[r + 1])
        if e == s:
            if not bool((C[r] == 0).all()):
                return False
            continue
        t32 = B[cc[s:e].long()]
        if vv is not None:
            t32 = t32 * vv[s:e][:, None]
        if reduce_name == "max":
            if not torch.equal(C[r], t32.max(0).values):
                return False
            continue
        t = t32.double()
        want, mag = t.sum(0), t.abs().sum(0)
        if reduce_name == "mean":
            want, mag = want / (e - s), mag / (e - s)
        if not bool(((C[r].double() - want).abs() <= 1e-5 * want.abs() + 1e-6 * mag).all()):
            return False
    return True


def secondary_run(name, args, dev, hbm_peak):
    """BASELINE configs 3 and 4 inside the driver-run record: products-like feat 128 (gspmm u_mul_e_max) and arxiv-like
    K 256 (SDDMM), each device-timed like the headline (CUDA events, warm-up, inputs larger than the L2), with the
    per-launch kernel time, the roofline fraction, a parity check and the reference's own CUDA on the same buffers."""
    import torch
    import dgsparse._kernels as K
    import dgsparse._lib as L
    wl = make_workload(name, 1.0)
    rowptr, col, N = wl["rowptr"], wl["col"], wl["N"]
    M, nnz = rowptr.size - 1, int(col.size)
    k_touched = int(np.unique(col).size)
    rp, cc = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
    gen = torch.Generator(dev).manual_seed(4321)
    steps, warm = 10, 3
    if wl["op"] == "sddmm_csr":
        vv = None
        D1, D2 = torch.rand(M, N, device=dev, generator=gen), torch.rand(M, N, device=dev, generator=gen)
        out = torch.empty(1, nnz, device=dev)

        def step():
            L.check(L.lib.dgs_sddmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), N, D2.data_ptr(), N,
                                        None, 0, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "sddmm")
        alg_bytes = 4 * (M + 1) + 4 * nnz + 4 * M * N + 4 * k_touched * N + 4 * nnz
        main_id, B = 3, None
    else:
        vv = torch.rand(nnz, device=dev, generator=gen) + 0.5
        B = torch.rand(M, N, device=dev, generator=gen)
        D1 = D2 = None
        out = torch.empty(M, N, device=dev)

        def step():
            K.spmm(rp, cc, vv, B, L.MAX, L.MUL, out=out)
        alg_bytes = algorithmic_bytes_spmm(M, nnz, N, k_touched, True)
        main_id = 1
    flop = 2.0 * nnz * N
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    L.lib.dgs_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ids, ms = (ctypes.c_int * 4096)(), (ctypes.c_float * 4096)()
    nrec = L.lib.dgs_profile_collect(4096, ids, ms)
    L.lib.dgs_profile_enable(0)
    # a step may be several launches of the kernel family (column-slab passes: one per slab, plus the partition, id 7)
    kern = [ms[i] for i in range(nrec) if ids[i] in (main_id, 7)]
    step_ms = e0.elapsed_time(e1) / steps
    k_avg = sum(kern) / steps if kern else step_ms
    launches = nrec // steps if nrec else None
    if wl["op"] == "sddmm_csr":
        e = np.unique(np.concatenate([np.linspace(0, nnz - 1, 256).astype(np.int64)]))
        rows = np.searchsorted(rowptr, e, side="right") - 1
        want = (D1[torch.from_numpy(rows).to(dev)].double() * D2[cc[torch.from_numpy(e).to(dev)].long()].double()).sum(1)
        got = out[0][torch.from_numpy(e).to(dev)].double()
        ok = bool(((got - want).abs() <= 1e-5 * want.abs()).all())     # inputs >= 0: purely relative
    else:
        ok = sampled_rows_ok(rowptr, rp, cc, vv, B, out, "max")
    traffic, tnote = ncu_traffic(name)
    res = {"workload": f"{wl['op']} on {wl['name']}, M=K={M}, nnz={nnz}, feat={N}, fp32", "steps": steps, "warmup": warm,
           "ms_per_step": step_ms, "value": flop / (step_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "parity_ok": ok,
           "profiled_scopes_per_step": launches,
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (k_avg * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (k_avg * 1e-3) / 1e9 / hbm_peak, "traffic": traffic, "traffic_source": tnote,
                        "kernel_ms_avg": k_avg, "algorithmic_bytes_per_launch": alg_bytes,
                        "note": "kernel_ms_avg = all launches of the SpMM / SDDMM kernel family in one step (a column-slab SpMM "
                                "is one partition + one kernel per slab); achieved = algorithmic bytes of the step / that"}}
    if not args.no_ref_cuda:
        res["reference_cuda"] = reference_cuda_run(wl, M, N, nnz, rp, cc, vv, B, D1, D2, step, steps, warm, flop)
    return res


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)

    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = make_workload(args.workload, args.scale)
        gflops, info = cpu_reference_run(wl, args.steps, args.warmup, budget_s=60.0)
        M, nnz = wl["rowptr"].size - 1, wl["col"].size
        line = {"impl": "reference", "metric": "spmm_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": n_gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["ms_per_sample"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(wl, M, nnz, wl["N"], n_gpus, wl["has_value"]),
                           "parallelism": "CPU reference (spmm_reference_host, example/util/sp_util.hpp:62-84) on the host cores, "
                                          "rank 0 only, a bounded row sample of the workload per step (cpu_baseline.sample)"},
                "cpu_baseline": info,
                "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import dgsparse  # noqa: F401
    import dgsparse._lib as L
    import dgsparse._kernels as K
    from dgsparse.distributed import ColumnShardedSpMM
    from tools import graphs

    wl = make_workload(args.workload, args.scale)
    rowptr, col, N = wl["rowptr"], wl["col"], wl["N"]
    M, nnz = rowptr.size - 1, int(col.size)
    k_touched = int(np.unique(col).size) if nnz < 5e8 else M
    val = graphs.uniform(nnz, 1) if wl["has_value"] else None
    hbm_peak, peak_src = peaks()
    has_value = val is not None

    rp, cc = torch.from_numpy(rowptr).to(dev), torch.from_numpy(col).to(dev)
    vv = torch.from_numpy(val).to(dev) if val is not None else None
    gen = torch.Generator(dev).manual_seed(1234 + rank)

    extra = {}
    if wl["op"] == "sddmm_csr":
        D1 = torch.rand(M, N, device=dev, generator=gen)
        D2 = torch.rand(M, N, device=dev, generator=gen)
        out = torch.empty(1, nnz, device=dev)

        def step():
            L.check(L.lib.dgs_sddmm_csr(M, N, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), N, D2.data_ptr(), N,
                                        None, 0, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "sddmm")
        flop = 2.0 * nnz * N
        alg_bytes = 4 * (M + 1) + 4 * nnz + 4 * M * N + 4 * k_touched * N + 4 * nnz
        launches_per_step, main_id = 1, 3
        mode = "replicas"
    else:
        reduce = L.MAX if "max" in wl["op"] else L.SUM
        B = torch.rand(M, N, device=dev, generator=gen)
        sh = ColumnShardedSpMM(rp, cc, vv, N, reduce=reduce, compute=L.MUL, mode=args.mode)
        mode = sh.mode

        def step():
            sh(B)
        flop = 2.0 * nnz * N * world
        alg_bytes = algorithmic_bytes_spmm(M, nnz, N, k_touched, vv is not None)
        launches_per_step, main_id = 2, 1
        if world > 1:
            extra["comm_bytes_in_per_rank"] = 4 * M * N * (world - 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- warm-up, then EXACTLY K timed steps ----------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.lib.dgs_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    ids = (ctypes.c_int * (64 * args.steps + 8))()
    ms = (ctypes.c_float * (64 * args.steps + 8))()
    nrec = L.lib.dgs_profile_collect(len(ids), ids, ms)
    L.lib.dgs_profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    kern_ms = [ms[i] for i in range(nrec) if ids[i] in (main_id, 7)]     # 7 = the column-slab partition (products-like operands)
    fix_ms = [ms[i] for i in range(nrec) if ids[i] == 2]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = flop / (ms_per_step * 1e-3) / 1e9

    # --- e2e: the same op through the HOST-pointer C-ABI call, H2D + D2H inside the timed region -------
    e2e = None
    if not args.no_e2e and wl["op"] != "sddmm_csr":
        hp = lambda a: torch.from_numpy(a).pin_memory()
        h_rp, h_cc = hp(rowptr), hp(col)
        h_val = hp(val) if val is not None else None
        h_B = B.cpu().pin_memory()
        h_C = torch.empty(M, N, dtype=torch.float32).pin_memory()
        red = L.MAX if "max" in wl["op"] else L.SUM

        if world == 1:
            def e2e_step():
                L.check(L.lib.dgs_spmm_csr_host(M, M, N, nnz, h_rp.data_ptr(), h_cc.data_ptr(),
                                                h_val.data_ptr() if h_val is not None else None, h_B.data_ptr(),
                                                h_C.data_ptr(), None, red, L.MUL), "dgs_spmm_csr_host")
            h2d = 4 * (M + 1) + 4 * nnz + (4 * nnz if val is not None else 0) + 4 * M * N
            d2h = 4 * M * N
            api = "dgs_spmm_csr_host (include/dgsparse_b200.h): pinned host CSR + B in, C out, every step"
        else:
            # host operands on every rank; the CSR crosses PCIe once per box (1/world per rank) and is replicated by an
            # NCCL all-gather over NVLink; every rank gets its own C panel back on the host
            from dgsparse.distributed import HostColumnShardedSpMM
            hop = HostColumnShardedSpMM(M, nnz, N, val is not None, dev, reduce=red, compute=L.MUL, mode=args.mode)

            def e2e_step():
                hop(h_rp, h_cc, h_val, h_B, h_C)
            tot = torch.tensor([hop.h2d_bytes + 4 * M * N, hop.d2h_bytes], device=dev, dtype=torch.float64)
            dist.all_reduce(tot)
            h2d, d2h = int(tot[0].item()), int(tot[1].item())
            api = ("dgsparse.distributed.HostColumnShardedSpMM: pinned host CSR slice (1/world per rank) + B panel in, "
                   "NCCL all-gather of col/val over NVLink, fused multicast / peer-store SpMM, own C panel out; bytes are whole-job totals")
        resident = None
        if world == 1:
            # the GNN use of the same host path: the CSR is uploaded ONCE (dgs_csr_upload, outside the timed region — it is not
            # a per-step input when A is fixed), every step moves B in and C out (dgs_spmm_csr_resident_host).  Reported
            # beside `e2e`, which keeps re-sending the whole CSR every step.
            hnd = ctypes.c_void_p()
            L.check(L.lib.dgs_csr_upload(M, M, nnz, h_rp.data_ptr(), h_cc.data_ptr(), h_val.data_ptr() if h_val is not None else None,
                                         ctypes.byref(hnd)), "dgs_csr_upload")
            for _ in range(2):
                L.check(L.lib.dgs_spmm_csr_resident_host(hnd, N, h_B.data_ptr(), h_C.data_ptr(), None, red, L.MUL), "resident")
            rsteps = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(rsteps):
                L.check(L.lib.dgs_spmm_csr_resident_host(hnd, N, h_B.data_ptr(), h_C.data_ptr(), None, red, L.MUL), "resident")
            tr = (time.perf_counter() - t0) / rsteps
            L.lib.dgs_csr_free(hnd)
            resident = {"value": flop / tr / 1e9, "unit": "GFLOP/s", "ms_per_step": tr * 1e3, "steps": rsteps,
                        "h2d_bytes_per_step": 4 * M * N, "d2h_bytes_per_step": 4 * M * N,
                        "api": "dgs_csr_upload once (untimed: A is not a per-step input when the graph is fixed) + "
                               "dgs_spmm_csr_resident_host per step: pinned B in, C out"}
        for _ in range(2):
            e2e_step()
        barrier()
        ksteps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / ksteps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": flop / float(te.item()) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(te.item()) * 1e3, "steps": ksteps, "api": api}
        if resident is not None:
            extra["e2e_resident_csr"] = resident

    # --- the reference's own CUDA kernels (oracle/_ref/libref_cuda.so, built unmodified for sm_100a) on the
    #     same device buffers, same timing loop: the "vs reference CUDA" comparison of SURVEY.md §8d.
    ref_cuda = None
    if world == 1 and not args.no_ref_cuda:
        ref_cuda = reference_cuda_run(wl, M, N, nnz, rp, cc, vv, locals().get("B"), locals().get("D1"),
                                      locals().get("D2"), step, args.steps, max(args.warmup, 3), flop)

    # --- parity inside the bench run (every N): the C this rank now holds against a local single-GPU recompute ---------
    parity = None
    if wl["op"] != "sddmm_csr":
        from dgsparse.distributed import panels_to_row_major
        red_name = "max" if "max" in wl["op"] else "sum"
        C_all = sh(B)
        if sh.mode == "nccl":
            C_all = panels_to_row_major(C_all)
        torch.cuda.synchronize()
        ok, what = True, []
        if world > 1:
            # every rank regenerates every rank's B panel (same seeds, same device RNG) and runs the single-GPU kernel on
            # it: the fused multicast / peer-store exchange must have delivered all N panels bit-identically
            for r in range(world):
                Br = torch.rand(M, N, device=dev, generator=torch.Generator(dev).manual_seed(1234 + r))
                ref_panel = K.spmm(rp, cc, vv, Br, reduce, L.MUL)
                ok = ok and bool(torch.equal(C_all[:, r * N:(r + 1) * N], ref_panel))
                del Br, ref_panel
            what.append(f"all {world} column panels of C[M,{N * world}] on every rank bit-identical to a local single-GPU recompute")
        ok = ok and sampled_rows_ok(rowptr, rp, cc, vv, B, C_all[:, rank * N:(rank + 1) * N], red_name)
        what.append("72 sampled rows (8 longest included) of this rank's panel against an fp64 recompute, 1e-5 relative")
        okt = torch.tensor([1 if ok else 0], device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        parity = {"parity_ok": bool(okt.item()), "checked": "; ".join(what)}

    # --- N>1: where the step time goes and a strong-scaling leg (feat 512 in total) ----------
    legs = None
    if wl["op"] != "sddmm_csr" and not args.no_legs:
        def timed(fn, n=5, warm=2):
            for _ in range(warm):
                fn()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            barrier()
            tt = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        legs = {}
        if world > 1:
            C_loc = torch.empty(M, N, device=dev)
            panels = torch.empty(world, M, N, device=dev)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            legs["compute_only_ms"] = timed(lambda: K.spmm(rp, cc, vv, B, reduce, L.MUL, out=C_loc))
            legs["comm_only_ms"] = timed(lambda: dist.all_gather_into_tensor(panels, C_loc))
            legs["barrier_only_ms"] = timed(lambda: sh.barrier(), n=20)
            legs["nccl_allreduce_1int_ms"] = timed(lambda: dist.all_reduce(flag), n=20)
            legs["note"] = ("max over ranks; compute_only = the single-GPU kernel + fix-up with local stores, comm_only = one "
                            "ncclAllGather of the [M, 64] panels (the unfused baseline's exchange), barrier_only = what closes a fused step "
                            "(dgs_mcast_barrier: one multimem.red + spin, in mcast mode; the one-int NCCL all-reduce it replaced is timed "
                            "beside it); end to end = ms_per_step")
            del C_loc, panels
        if world == 1:
            # a cold-L2 figure beside the steady-state one.  Each step is preceded by a 512 MB write (4x the L2),
            # timed per step with its own event pair so that the flush itself stays outside the measurement.
            flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
            ts = []
            for _ in range(5):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); step(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            legs["cold_l2_ms_per_step"] = sorted(ts)[len(ts) // 2]
            legs["cold_l2_note"] = "median of 5 steps, each after a 512 MB write that evicts the L2 (B and the first CSR bytes come from HBM)"
            del flush
        if 512 % world == 0:
            n_s = 512 // world
            Bs = torch.rand(M, n_s, device=dev, generator=torch.Generator(dev).manual_seed(99 + rank))
            shs = ColumnShardedSpMM(rp, cc, vv, n_s, reduce=reduce, compute=L.MUL, mode=args.mode)
            t_s = timed(lambda: shs(Bs), n=5 if world > 1 else 3)
            legs["strong_leg"] = {"feat_total": 512, "feat_per_gpu": n_s, "ms_per_step": t_s, "exchange": shs.mode,
                                  "value": 2.0 * nnz * 512 / (t_s * 1e-3) / 1e9, "unit": "GFLOP/s",
                                  "note": "SURVEY 8d config 5 as STRONG scaling: total feature width fixed at 512"}
            shs.close()
            del shs, Bs
        torch.cuda.empty_cache()

    # --- configs 3 and 4 in the same record (N=1 default run only) ------------------------------------------------------
    secondary = None
    if world == 1 and args.workload == "reddit64" and args.scale == 1.0 and not args.no_secondary:
        keep = (M, nnz, N)
        sh.close()
        rp = cc = vv = B = sh = C_all = None      # release the headline workload's device memory (closures see the cells)
        torch.cuda.empty_cache()
        secondary = {}
        for name in ("products128", "arxiv256"):
            try:
                secondary[name] = secondary_run(name, args, dev, hbm_peak)
            except Exception as ex:
                secondary[name] = {"error": repr(ex)}
            torch.cuda.empty_cache()
        M, nnz, N = keep

    if rank == 0:
        traffic, traffic_note = ncu_traffic(args.workload)
        k_avg = sum(kern_ms) / args.steps if kern_ms else ms_per_step     # all launches of the kernel family in one step
        achieved = alg_bytes / (k_avg * 1e-3) / 1e9
        line = {
            "metric": "spmm_gflops" if wl["op"] != "sddmm_csr" else "sddmm_gflops",
            "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(wl, M, nnz, N, world, has_value),
                       "parallelism": f"feature-axis column shard x{world}, CSR replicated, exchange={mode}",
                       "l2": "no flush: inputs per step (%.0f MB) exceed the 126 MB L2" % (alg_bytes / 1e6)},
            "achieved_hbm_gbs": alg_bytes * (world) / (ms_per_step * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak,
                         "traffic": traffic if args.scale == 1.0 else None, "traffic_source": traffic_note,
                         "peak_source": peak_src,
                         "kernel": "spmm_rowseg_kernel" if main_id == 1 else "sddmm_ring_kernel",
                         "kernel_ms_avg": k_avg, "fixup_ms_avg": (sum(fix_ms) / args.steps) if fix_ms else None,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "gather_bytes_per_launch": 4.0 * nnz * N,
                         "l2_gather_gbs": 4.0 * nnz * N / (k_avg * 1e-3) / 1e9,
                         "note": "gathered B rows (nnz*N*4 B) are served by L2 (l2_gather_gbs; ncu: lts__lts2xbar_cycles_active 83 % of "
                                 "peak on reddit@64, profiles/r02_ncu_full_reddit64.txt): the L2 slices' output, not HBM, bounds "
                                 "the kernel when B fits L2 — see DESIGN.md 4.1"},
            "gpu_launches": max(launches_per_step * args.steps, nrec),
            "clocks": clocks,
        }
        line.update(extra)
        if parity is not None:
            line.update(parity)
        if legs:
            line["legs"] = legs
        if secondary is not None:
            line["secondary"] = secondary
        if e2e is not None:
            line["e2e"] = e2e
        if ref_cuda is not None:
            line["reference_cuda"] = ref_cuda
        if world == 1 and not args.no_cpu_baseline and wl["op"] != "sddmm_csr":
            try:
                _, info = cpu_reference_run(wl, steps=2, warmup=1, budget_s=12.0)
                line["cpu_baseline"] = info
            except Exception as ex:  # the checker is optional at bench time, the product is not
                line["cpu_baseline"] = {"error": str(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
