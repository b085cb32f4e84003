// options.h — experiment / test knobs of the library.  Read ONCE from the environment (a getenv per knob per call was a
// measurable part of a 15 us call), overridable at run time with dgs_set_option(name, value) (tests, tools).
//   name             environment variable      meaning (unset = -1 = the library decides)
//   spmm_rowpar      DGS_SPMM_ROWPAR           0 never / 1 always take the row-parallel single-launch SpMM
//   spmm_panel       DGS_SPMM_PANEL            32: 32-column panels (8-lane geometry) on matrices wider than 32
//   spmm_no_pdl      DGS_SPMM_NO_PDL           1: fix-up grid not launched as a programmatic dependent
//   spmm_segs        DGS_SPMM_SEGS             segments per resident lane group
//   spmm_chunk_cap   DGS_SPMM_CHUNK_CAP        longest segment (nonzeros)
//   sddmm_no_ring    DGS_SDDMM_NO_RING         1: register-staged SDDMM kernel for every K; 0: the ring also for K = 64 on small inputs
//   sddmm_stages     DGS_SDDMM_STAGES          2 | 3 ring stages
//   sddmm_chunk      DGS_SDDMM_CHUNK           edges per warp of the ring kernel (multiple of 32)
//   sddmm_wpc        DGS_SDDMM_WPC             warps per CTA of the ring kernel
//   spconv_ctas      DGS_SPCONV_CTAS           1 .. 3: persistent CTAs per SM of the pipelined spconv kernel (unset: as many as fit)
//   sddmm_threads    DGS_SDDMM_THREADS         256: 256-thread CTAs for the register-staged SDDMM kernel (unset: 64)
//   sddmm_d1slots    DGS_SDDMM_D1SLOTS         2 | 4: D1 row slots per ring stage
//   spmm_colmajor    DGS_SPMM_COLMAJOR         0: gespmmCsrSpMM(transpose_BC = false) takes the thread-per-element kernel instead of transposes around the row-major one
#pragma once

namespace dgs {
enum Option { OPT_SPMM_ROWPAR = 0, OPT_SPMM_PANEL, OPT_SPMM_NO_PDL, OPT_SPMM_SEGS, OPT_SPMM_CHUNK_CAP, OPT_SDDMM_NO_RING, OPT_SDDMM_STAGES, OPT_SDDMM_CHUNK, OPT_SDDMM_WPC, OPT_SPCONV_CTAS, OPT_SPMM_COLMAJOR, OPT_SDDMM_THREADS, OPT_SDDMM_D1SLOTS, OPT_COUNT };
int option(Option o);                          // -1 when unset
int set_option(const char *name, int value);   // value < 0 clears the override (back to the environment); 0 ok, -1 unknown name
}  // namespace dgs
