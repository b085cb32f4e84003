// spmm_inst.cu — compiled once per (INST_VEC, INST_G) pair (see build.py); instantiates every
// (REDUCE, COMPUTE, ARG) flavour of spmm_rowseg_kernel for that lane-group geometry and exports a
// lookup function the dispatcher in spmm.cu uses.
#include "spmm_rowseg.cuh"
#include "spmm_rowpar.cuh"

#ifndef INST_VEC
#error "compile with -DINST_VEC=<1|4> -DINST_G=<4|8|16|32>"
#endif

namespace dgs {

#define DGS_CAT_(a, b, c, d) a##b##c##d
#define DGS_CAT(a, b, c, d) DGS_CAT_(a, b, c, d)
#define DGS_LOOKUP DGS_CAT(spmm_lookup_v, INST_VEC, _g, INST_G)

template <int RED, bool ARG> static SpmmKernel by_comp(int comp) {
  switch (comp) {
  case C_MUL: return spmm_rowseg_handle<INST_VEC, INST_G, RED, C_MUL, ARG>();
  case C_COPY: return spmm_rowseg_handle<INST_VEC, INST_G, RED, C_COPY, ARG>();
  default: break;
  }
  if (comp == C_MASK) {  // max/min backward wrt dense: SUM only, no arg output
    if (RED == R_SUM && !ARG) return spmm_rowseg_handle<INST_VEC, INST_G, R_SUM, C_MASK, false>();
    return SpmmKernel();
  }
  if (ARG) return SpmmKernel();  // arg index only exists on the torch face (multiply / no value)
  switch (comp) {
  case C_ADD: return spmm_rowseg_handle<INST_VEC, INST_G, RED, C_ADD, false>();
  case C_SUB: return spmm_rowseg_handle<INST_VEC, INST_G, RED, C_SUB, false>();
  case C_DIV: return spmm_rowseg_handle<INST_VEC, INST_G, RED, C_DIV, false>();
  default: return SpmmKernel();
  }
}

SpmmKernel DGS_LOOKUP(int red, int comp, bool arg) {
  switch (red) {
  case R_SUM:
  case R_MEAN: return arg ? SpmmKernel() : by_comp<R_SUM, false>(comp);
  case R_MAX: return arg ? by_comp<R_MAX, true>(comp) : by_comp<R_MAX, false>(comp);
  case R_MIN: return arg ? by_comp<R_MIN, true>(comp) : by_comp<R_MIN, false>(comp);
  default: return SpmmKernel();
  }
}

// Row-parallel single-launch kernels (spmm_rowpar.cuh): the 16-byte geometries only, multiply / no-edge-value.
#if INST_VEC == 4 && INST_G <= 16
#define DGS_ROWPAR_LOOKUP DGS_CAT(spmm_rowpar_lookup_v, INST_VEC, _g, INST_G)
template <int RED, bool ARG> static SpmmLaunchFn rowpar_by_comp(int comp) {
  switch (comp) {
  case C_MUL: return &launch_spmm_rowpar<INST_VEC, INST_G, RED, C_MUL, ARG>;
  case C_COPY: return &launch_spmm_rowpar<INST_VEC, INST_G, RED, C_COPY, ARG>;
  default: return nullptr;
  }
}
SpmmLaunchFn DGS_ROWPAR_LOOKUP(int red, int comp, bool arg) {
  switch (red) {
  case R_SUM:
  case R_MEAN: return arg ? nullptr : rowpar_by_comp<R_SUM, false>(comp);
  case R_MAX: return arg ? rowpar_by_comp<R_MAX, true>(comp) : rowpar_by_comp<R_MAX, false>(comp);
  case R_MIN: return arg ? rowpar_by_comp<R_MIN, true>(comp) : rowpar_by_comp<R_MIN, false>(comp);
  default: return nullptr;
  }
}
#endif

}  // namespace dgs
