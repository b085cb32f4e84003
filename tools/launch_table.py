#!/usr/bin/env python
"""Prints an ncu launch list (--csv --metrics gpu__time_duration.sum[,dram__bytes_*]) as one line per launch."""
import csv
import sys


def main():
    path, last = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = {}
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
    for (i, k), m in sorted(d.items())[-last:]:
        us = m.get("gpu__time_duration.sum", 0) / 1e3
        rd, wr = m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6
        print(f"{i:5d} {us:10.1f} us  rd {rd:9.1f} MB  wr {wr:9.1f} MB  {k[:90]}")


if __name__ == "__main__":
    main()
