"""Generalized SpMM — mirror of the gspmm-fp pybind module `spmm` (src/gspmm-fp/gspmm.cc:27-44) and of
its 20 Python wrappers (example/gspmm-fp/util.py:17-110).

    GSpMM_u_e(rowptr, colind, edge_val, feat, REDUCEOP, COMPUTEOP)   out[i] = REDUCE_j COMPUTE(e_ij, feat[j])
    GSpMM_u(rowptr, colind, feat, REDUCEOP)                          copy_u_<reduce>
COMPUTE(a = edge, b = feat): ADD a+b, SUB b-a, MUL a*b, DIV b/a (src/gspmm-fp/gspmm.h:53-79).
"""
import enum

import torch

from . import _kernels as K


class REDUCEOP(enum.IntEnum):  # include/gspmm.h:13
    SUM = 0
    MAX = 1
    MIN = 2
    MEAN = 3


class COMPUTEOP(enum.IntEnum):  # include/gspmm.h:14
    ADD = 0
    SUB = 1
    MUL = 2
    DIV = 3


SUM, MAX, MIN, MEAN = REDUCEOP.SUM, REDUCEOP.MAX, REDUCEOP.MIN, REDUCEOP.MEAN   # pybind export_values()
ADD, SUB, MUL, DIV = COMPUTEOP.ADD, COMPUTEOP.SUB, COMPUTEOP.MUL, COMPUTEOP.DIV


def _check(t, dtype, name):  # assertTensor, src/gspmm-fp/gspmm.cc:3-7
    if not (t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise AssertionError(f"{name}: expected a contiguous CUDA tensor of {dtype}")


def GSpMM_u_e(A_rowptr, A_colind, A_csrVal, B, re_op, comp_op):
    _check(A_rowptr, torch.int32, "rowptr")
    _check(A_colind, torch.int32, "colind")
    _check(A_csrVal, torch.float32, "edge_val")
    _check(B, torch.float32, "feat")
    return K.spmm(A_rowptr, A_colind, A_csrVal.reshape(-1), B, int(re_op), int(comp_op))


def GSpMM_u(A_rowptr, A_colind, B, op):
    _check(A_rowptr, torch.int32, "rowptr")
    _check(A_colind, torch.int32, "colind")
    _check(B, torch.float32, "feat")
    return K.spmm(A_rowptr, A_colind, None, B, int(op))


def _mk_u_e(cop, rop):
    def f(rowptr, colind, edge_feature, node_feat):
        return GSpMM_u_e(rowptr, colind, edge_feature, node_feat, rop, cop)
    return f


def _mk_u(rop):
    def f(rowptr, colind, node_feat):
        return GSpMM_u(rowptr, colind, node_feat, rop)
    return f


for _c in COMPUTEOP:
    for _r in REDUCEOP:
        _n = f"u_{_c.name.lower()}_e_{_r.name.lower()}"
        globals()[_n] = _mk_u_e(_c, _r)
        globals()[_n].__name__ = _n
for _r in REDUCEOP:
    _n = f"copy_u_{_r.name.lower()}"
    globals()[_n] = _mk_u(_r)
    globals()[_n].__name__ = _n
