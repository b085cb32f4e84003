#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [out.txt] [--traffic WORKLOAD FAMILY]
--traffic records the first kernel's DRAM read + write bytes in profiles/traffic.json together with the hash of the kernel
sources it was taken on (tools/csrc_hash.py); bench.py prints roofline.traffic only while that hash matches the code."""
import csv
import io
import json
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep = sys.argv[1]
    traffic_for = None
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        traffic_for = (sys.argv[i + 1], sys.argv[i + 2])
        del sys.argv[i:i + 3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none summary of {rep}", ""]
    for r in rows[2:]:
        out.append("kernel: " + r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                out.append(f"  {w:80s} {r[hdr.index(w)]:>20s} {units[hdr.index(w)]}")
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        u = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        out.append(f"  traffic (dram read + write): {rd} {u[0]} + {wr} {u[1]}")
        out.append("")
    if traffic_for and len(rows) > 2:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from tools import csrc_hash
        r = rows[2]
        tb = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")]) + \
            to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        db = json.load(open(path))
        db[traffic_for[0]] = {"kernel": r[hdr.index("Kernel Name")], "family": traffic_for[1], "traffic_bytes": int(tb),
                              "csrc_sha": csrc_hash.family_hash(traffic_for[1]),
                              "source": "profiles/" + os.path.basename(sys.argv[2]) if len(sys.argv) > 2 else rep}
        json.dump(db, open(path, "w"), indent=1)
        out.append(f"recorded {int(tb)} B for {traffic_for[0]} in profiles/traffic.json")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
