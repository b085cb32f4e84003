#!/usr/bin/env python
"""Builds a VARIANT of libdgsparse_b200.so with extra -D flags into tools/_build/variant_<tag>/ (not tracked; it travels to the GPU
box with the snapshot) for A/B runs against the shipped library in one process (tools/exp_ab_variant.py).

    python tools/build_variant.py t64 -DDGS_SPMM_THREADS=64
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, defs = sys.argv[1], sys.argv[2:]
    spec = importlib.util.spec_from_file_location("dgs_build", os.path.join(ROOT, "dgsparse-lib_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    out = os.path.join(ROOT, "tools", "_build", "variant_" + tag)
    b.OBJ = os.path.join(out, "obj")
    b.LIB = os.path.join(out, "libdgsparse_b200_%s.so" % tag)
    b.FLAGS = b.FLAGS + defs
    os.makedirs(b.OBJ, exist_ok=True)
    print(b.build(force=True, verbose=True))


if __name__ == "__main__":
    main()
