// options.cu — see options.h
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "options.h"

namespace dgs {
namespace {
const char *kNames[OPT_COUNT] = {"spmm_rowpar", "spmm_panel", "spmm_no_pdl", "spmm_segs", "spmm_chunk_cap", "sddmm_no_ring", "sddmm_stages",
                                 "sddmm_chunk", "sddmm_wpc", "spconv_ctas", "spmm_colmajor", "sddmm_threads", "sddmm_d1slots"};
const char *kEnv[OPT_COUNT] = {"DGS_SPMM_ROWPAR", "DGS_SPMM_PANEL", "DGS_SPMM_NO_PDL", "DGS_SPMM_SEGS", "DGS_SPMM_CHUNK_CAP", "DGS_SDDMM_NO_RING",
                               "DGS_SDDMM_STAGES", "DGS_SDDMM_CHUNK", "DGS_SDDMM_WPC", "DGS_SPCONV_CTAS", "DGS_SPMM_COLMAJOR", "DGS_SDDMM_THREADS", "DGS_SDDMM_D1SLOTS"};
std::atomic<int> g_env[OPT_COUNT];
std::atomic<int> g_override[OPT_COUNT];
std::once_flag g_once;

void load_env() {
  for (int i = 0; i < OPT_COUNT; i++) {
    const char *e = getenv(kEnv[i]);
    g_env[i].store(e ? atoi(e) : -1);
    g_override[i].store(-1);
  }
}
}  // namespace

int option(Option o) {
  std::call_once(g_once, load_env);
  const int v = g_override[o].load(std::memory_order_relaxed);
  return v >= 0 ? v : g_env[o].load(std::memory_order_relaxed);
}

int set_option(const char *name, int value) {
  std::call_once(g_once, load_env);
  for (int i = 0; i < OPT_COUNT; i++)
    if (strcmp(name, kNames[i]) == 0) { g_override[i].store(value < 0 ? -1 : value); return 0; }
  return -1;
}
}  // namespace dgs
