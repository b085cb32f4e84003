// spmm_narrow_inst.cu — compiled once per INST_G (2 or 4; see build.py): the narrow-panel SpMM kernels
// (spmm_narrow.cuh) for the reduce / compute flavours the wide-operand workloads use: sum | max | min (mean = sum + a
// division in the epilogue) x multiply | no-edge-value, with and without the arg index.  Other compute ops keep the
// 64-column path.
#include "spmm_narrow.cuh"

#ifndef INST_G
#error "compile with -DINST_G=<2|4>"
#endif

namespace dgs {

#define DGS_NCAT_(a, b) a##b
#define DGS_NCAT(a, b) DGS_NCAT_(a, b)
#define DGS_NARROW_LOOKUP DGS_NCAT(spmm_narrow_lookup_g, INST_G)

template <int RED, bool ARG> static SpmmLaunchFn narrow_by_comp(int comp) {
  switch (comp) {
  case C_MUL: return &launch_spmm_narrow<INST_G, RED, C_MUL, ARG>;
  case C_COPY: return &launch_spmm_narrow<INST_G, RED, C_COPY, ARG>;
  default: return nullptr;
  }
}

SpmmLaunchFn DGS_NARROW_LOOKUP(int red, int comp, bool arg) {
  switch (red) {
  case R_SUM:
  case R_MEAN: return arg ? nullptr : narrow_by_comp<R_SUM, false>(comp);
  case R_MAX: return arg ? narrow_by_comp<R_MAX, true>(comp) : narrow_by_comp<R_MAX, false>(comp);
  case R_MIN: return arg ? narrow_by_comp<R_MIN, true>(comp) : narrow_by_comp<R_MIN, false>(comp);
  default: return nullptr;
  }
}

}  // namespace dgs
