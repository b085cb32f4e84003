"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dgsparse-lib_b200", "lib", "libdgsparse_b200.so")


def _declared():
    names = set()
    for h in ("dgsparse.h", "dgsparse_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"#else.*?#endif", "#endif", src, flags=re.S)   # the C-only duplicate prototype
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src):
            names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    assert os.path.exists(LIB)
    lib = ctypes.CDLL(LIB)
    decl = _declared()
    for must in ["spmm_cuda", "spmm_cuda_no_edge_value", "sddmm_cuda_coo", "sddmm_cuda_csr", "gespmmCsrSpMM",
                 "dgs_spmm_csr", "dgs_sddmm_csr", "dgs_csr2csc", "dgs_spmm_csr_host"]:
        assert must in decl, f"header parser lost {must}"
    missing = [n for n in sorted(decl) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    lib.dgs_cuda_version.restype = ctypes.c_int
    assert lib.dgs_cuda_version() // 1000 == 12
    lib.dgs_spmm_workspace_bytes.restype = ctypes.c_size_t
    lib.dgs_spmm_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int]
    assert lib.dgs_spmm_workspace_bytes(64, 1000000, 0) > 0


def test_python_binding_covers_the_headers():
    import dgsparse._lib as L
    bound = set(L.SIGNATURES) | {"gespmmCsrSpMM"}
    assert _declared() <= bound, sorted(_declared() - bound)


def test_package_surface_matches_reference():
    """dgsparse/__init__.py:46-49 __all__, the five torch ops (src/spmm.cpp:264-270), the gspmm-fp
    wrappers (example/gspmm-fp/util.py:17-110)."""
    import torch
    import dgsparse
    assert set(dgsparse.__all__) == {"spmm_sum", "spmm_max", "spmm_min", "spmm_mean", "Storage", "SparseTensor",
                                     "csr2csc"}
    for op in ["spmm_sum", "spmm_max", "spmm_min", "spmm_mean", "csr2csc"]:
        assert hasattr(torch.ops.dgsparse_spmm, op)
    for c in ["add", "sub", "mul", "div"]:
        for r in ["sum", "max", "min", "mean"]:
            assert callable(getattr(dgsparse.gspmm, f"u_{c}_e_{r}"))
    for r in ["sum", "max", "min", "mean"]:
        assert callable(getattr(dgsparse.gspmm, f"copy_u_{r}"))
    assert int(dgsparse.gspmm.REDUCEOP.MEAN) == 3 and int(dgsparse.gspmm.COMPUTEOP.DIV) == 3
    assert dgsparse._C.cuda_version() == dgsparse.cuda_version


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    import dgsparse._kernels as K
    rowptr = torch.tensor([0, 1], dtype=torch.int32)
    col = torch.tensor([0], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        K.spmm(rowptr, col, None, torch.ones(1, 4))


def test_compiled_op_library_loads_without_the_python_package():
    """The reference's boundary is a loadable op library (dgsparse/_spmm_cuda.so + torch.ops.load_library,
    dgsparse/__init__.py:16-26; TORCH_LIBRARY at src/spmm.cpp:264-270).  Ours (csrc/torch_ops.cpp) must load in a FRESH
    interpreter that never imports the dgsparse package or its Python op registration, and expose the same five ops with
    the schemas the reference's C++ signatures imply; on CPU tensors they fail loudly (no CPU path)."""
    import subprocess
    import sys
    import __graft_entry__
    __graft_entry__.build()
    so = os.path.join(ROOT, "dgsparse-lib_b200", "dgsparse", "_spmm_cuda.so")
    assert os.path.exists(so)
    code = f"""
import sys, torch
torch.ops.load_library({so!r})
assert 'dgsparse' not in sys.modules and 'dgsparse._ops' not in sys.modules
want = "(Tensor _0, Tensor _1, Tensor _2, Tensor _3, Tensor _4, Tensor _5, Tensor _6, bool _7, int _8) -> Tensor _0"
for op in ("spmm_sum", "spmm_max", "spmm_min", "spmm_mean"):
    s = str(getattr(torch.ops.dgsparse_spmm, op).default._schema)
    assert s == "dgsparse_spmm::" + op + want, s
assert str(torch.ops.dgsparse_spmm.csr2csc.default._schema) == "dgsparse_spmm::csr2csc(Tensor _0, Tensor _1, Tensor _2) -> Tensor[] _0"
i = lambda *v: torch.tensor(v, dtype=torch.int32)
try:
    torch.ops.dgsparse_spmm.spmm_sum(i(0, 1), i(0), torch.ones(1), i(0, 1), i(0), i(0), torch.ones(1, 4), True, 0)
except RuntimeError as e:
    assert "no CPU path" in str(e), e
else:
    raise SystemExit("CPU tensors did not raise")
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_package_uses_the_compiled_ops_by_default():
    import dgsparse
    assert dgsparse.ops_backend.startswith("compiled"), dgsparse.ops_backend
    assert dgsparse._C.__file__.endswith("_C.so")
