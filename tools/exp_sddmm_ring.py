#!/usr/bin/env python
"""SDDMM ring geometry sweep: warps per CTA (-> CTAs per SM from the occupancy API) x ring stages x edges per warp, through
the library's run-time knobs (dgs_set_option: sddmm_stages, sddmm_wpc, sddmm_chunk), on the arxiv-like graph (config 4) and
the reference's two fixtures (latency regime).  One JSON line per setting with the geometry the library reports
(dgs_sddmm_last_geometry); the first line of every (graph, K) is the library's own choice.

    python tools/exp_sddmm_ring.py [--quick]
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "dgsparse-lib_b200")]
import dgsparse._lib as L  # noqa: E402
from tools import graphs  # noqa: E402
from tools.bench_vs_ref import timeit  # noqa: E402

graphs.build()
quick = "--quick" in sys.argv
d1_mode = "--d1" in sys.argv              # two vs four D1 slots per ring stage (x a few chunk lengths)
chunks_mode = "--chunks" in sys.argv      # sweep the edges per warp at the library's own CTA geometry instead of the CTA size
st = torch.cuda.current_stream().cuda_stream


def setopt(**kw):
    for k in ("sddmm_stages", "sddmm_wpc", "sddmm_chunk", "sddmm_no_ring", "sddmm_d1slots"):
        L.lib.dgs_set_option(k.encode(), int(kw.get(k, -1)))


def geometry():
    w, c, e = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.lib.dgs_sddmm_last_geometry(ctypes.byref(w), ctypes.byref(c), ctypes.byref(e))
    return {"wpc": w.value, "ctas_per_sm": c.value, "edges_per_warp": e.value}


cases = [("arxiv-like", graphs.arxiv_like(1.0), (64, 128, 256, 512), 50),
         ("ca-CondMat (example/data)", graphs.load_fixture("ca-CondMat")[:2], (64, 128, 256, 512), 200),
         ("p2p-Gnutella31 (example/data)", graphs.load_fixture("p2p-Gnutella31")[:2], (64, 128, 256, 512), 200)]
for gname, (rowptr, col), widths, reps in cases:
    M, nnz = rowptr.size - 1, int(col.size)
    rp, cc = torch.from_numpy(rowptr).cuda(), torch.from_numpy(col).cuda()
    for Kd in widths:
        D1, D2 = torch.rand(M, Kd, device="cuda"), torch.rand(M, Kd, device="cuda")
        out = torch.empty(nnz, device="cuda")
        ref = None
        run = lambda: L.lib.dgs_sddmm_csr(M, Kd, nnz, rp.data_ptr(), cc.data_ptr(), D1.data_ptr(), Kd, D2.data_ptr(), Kd, None, 0,
                                          out.data_ptr(), st)
        settings = [dict()]
        if d1_mode:
            for nd in (4, 2):
                settings.append(dict(sddmm_d1slots=nd))
                for c in (96, 144, 192, 256):
                    settings.append(dict(sddmm_d1slots=nd, sddmm_chunk=c))
        elif chunks_mode:
            for c in (32, 48, 64, 96, 128, 160, 192, 256, 320, 384, 512, 768, 1024):
                settings.append(dict(sddmm_chunk=c))
            if Kd == 64:
                settings += [dict(sddmm_no_ring=0), dict(sddmm_no_ring=1)]
        else:
            for wpc in range(1, 17):
                settings.append(dict(sddmm_wpc=wpc))
            if not quick:
                for wpc in (4, 6, 8):
                    settings.append(dict(sddmm_stages=3, sddmm_wpc=wpc))
        seen = set()
        for s in settings:
            setopt(**s)
            t = timeit(run, reps)
            geo = geometry()
            key = (s.get("sddmm_stages", 2), geo["wpc"], geo["ctas_per_sm"], geo["edges_per_warp"], s.get("sddmm_no_ring", -1), s.get("sddmm_d1slots", -1))
            if s and key in seen:      # the override did not fit and the library fell back to a geometry already timed
                continue
            seen.add(key)
            if ref is None:
                ref = out.clone()
            print(json.dumps({"op": "sddmm_csr", "graph": gname, "nnz": nnz, "K": Kd, "setting": s or "library default", **geo,
                              "resident_warps_per_sm": geo["wpc"] * geo["ctas_per_sm"], "ms": t,
                              "bit_identical_to_default": bool(torch.equal(out, ref))}), flush=True)
        setopt()
        del D1, D2, out
