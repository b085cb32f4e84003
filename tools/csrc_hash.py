"""Hash of the kernel sources of one kernel family, so that a recorded ncu capture (profiles/traffic.json) can be tied to
the code it profiled: bench.py prints roofline.traffic only when the hashes agree.
    python tools/csrc_hash.py [family]"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dgsparse-lib_b200", "csrc")
FAMILIES = {
    "spmm": ["common.cuh", "spmm_rowseg.cuh", "spmm_rowpar.cuh", "spmm.cu", "spmm_inst.cu"],
    "sddmm": ["common.cuh", "sddmm.cu"],
    "csr2csc": ["common.cuh", "csr2csc.cu"],
    "spconv": ["spconv.cu"],
}


def family_hash(family):
    h = hashlib.sha1()
    for f in FAMILIES[family]:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    for fam in (sys.argv[1:] or sorted(FAMILIES)):
        print(fam, family_hash(fam))
