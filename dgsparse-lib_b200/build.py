"""Builds dgsparse-lib_b200/lib/libdgsparse_b200.so IN-TREE with nvcc for sm_100a.

    python dgsparse-lib_b200/build.py [--force] [--verbose]

Torch-free: every kernel and the C ABI live in csrc/*.cu.  The SpMM kernel is compiled once per lane-group geometry
(spmm_inst.cu with -DINST_VEC/-DINST_G) so the translation units build in parallel.

build_torch() then compiles the PyTorch operator boundary (host C++ only, no kernels) IN-TREE, the two files the
reference ships (setup.py:26-84): dgsparse/_spmm_cuda.so (csrc/torch_ops.cpp: TORCH_LIBRARY(dgsparse_spmm) + autograd over
the C ABI, loaded with torch.ops.load_library) and dgsparse/_C.so (csrc/version.cpp: pybind `cuda_version`).
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libdgsparse_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]

GEOMS = [(4, 4), (4, 8), (4, 16), (4, 32), (1, 4), (1, 8), (1, 16), (1, 32)]


def units():
    u = []
    for name in ["spmm.cu", "sddmm.cu", "csr2csc.cu", "cabi.cu", "spconv.cu", "kmap.cu", "options.cu"]:
        if os.path.exists(os.path.join(CSRC, name)):
            u.append((name, [], name.replace(".cu", ".o")))
    for v, g in GEOMS:
        u.append(("spmm_inst.cu", [f"-DINST_VEC={v}", f"-DINST_G={g}"], f"spmm_inst_v{v}_g{g}.o"))
    return u


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h", ".cu")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return max(m, os.path.getmtime(os.path.abspath(__file__)))


def _compile(job):
    src, defs, obj, verbose = job
    out = os.path.join(OBJ, obj)
    cmd = [NVCC] + ARCH + FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(out + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src} {defs}:\n{log}")
    if verbose:
        print(f"[build] {obj}")
    return out


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    newest = _deps_mtime()
    jobs, objs = [], []
    for src, defs, obj in units():
        out = os.path.join(OBJ, obj)
        objs.append(out)
        if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
            jobs.append((src, defs, obj, verbose))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(_compile, jobs))
    if jobs or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(f"[build] linked {LIB}")
    return LIB


TORCH_OPS = os.path.join(HERE, "dgsparse", "_spmm_cuda.so")
TORCH_C = os.path.join(HERE, "dgsparse", "_C.so")


def build_torch(force=False, verbose=False):
    """dgsparse/_spmm_cuda.so + dgsparse/_C.so with g++ against this interpreter's torch (needs libdgsparse_b200.so)."""
    import sysconfig

    import torch
    tdir = os.path.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    inc = [f"-I{tdir}/include", f"-I{tdir}/include/torch/csrc/api/include", "-I/usr/local/cuda/include",
           f"-I{sysconfig.get_paths()['include']}"]
    common = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-D_GLIBCXX_USE_CXX11_ABI={abi}"] + inc
    link = [f"-L{os.path.dirname(LIB)}", "-ldgsparse_b200", "-Wl,-rpath,$ORIGIN/../lib", f"-L{tdir}/lib", f"-Wl,-rpath,{tdir}/lib"]
    jobs = [(TORCH_OPS, "torch_ops.cpp", ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
                                          "-L/usr/local/cuda/lib64", "-lcudart"]),
            (TORCH_C, "version.cpp", [])]
    hdr = os.path.join(os.path.dirname(HERE), "include", "dgsparse_b200.h")
    todo = []
    for out, src, libs in jobs:
        srcp = os.path.join(CSRC, src)
        newest = max(os.path.getmtime(srcp), os.path.getmtime(hdr))
        if force or not os.path.exists(out) or os.path.getmtime(out) < newest:
            todo.append((out, common + [srcp, "-o", out] + link + libs))

    def run(job):
        out, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {out}:\n" + r.stdout + r.stderr)
        if verbose:
            print(f"[build] {out}")
        return out
    if todo:
        with cf.ThreadPoolExecutor(max_workers=2) as ex:
            list(ex.map(run, todo))
    return TORCH_OPS, TORCH_C


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_torch(force="--force" in sys.argv, verbose=True))
