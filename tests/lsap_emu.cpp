// Host emulation harness for pointtinybenchmark_b200/csrc/lsap_core.cuh (test infrastructure, compiled by
// tests/test_lsap.py with g++ -DPTB_LSAP_HOST_EMU): runs the kernel body of hungarian_v2_kernel with ONE emulated thread, so
// the bookkeeping of the GPU formulation is checked against scipy on a box without a GPU.  The transposed copy that
// lsap_prep_kernel writes on the device is built here with a plain loop.
#define PTB_LSAP_HOST_EMU 1
#include "../pointtinybenchmark_b200/csrc/lsap_core.cuh"
#include <stdlib.h>
#include <vector>

extern "C" int emu_hungarian_v2(int N, int n, const float* cost, int topk_k, const int32_t* row_idx, int64_t* out) {
  if (N <= 0 || n <= 0) return 0;
  for (int64_t e = 0; e < (int64_t)N * n; ++e)
    if (cost[e] != cost[e] || cost[e] == -INFINITY) return 2;
  std::vector<char> buf(ptb_lsap::ws_bytes(N, n) + 8);
  char* base = buf.data();
  base += (8 - (reinterpret_cast<uintptr_t>(base) & 7)) & 7;
  ptb_lsap::Ws w = ptb_lsap::ws_carve(base, N, n);
  if (n < N)
    for (int p = 0; p < N; ++p)
      for (int g = 0; g < n; ++g) w.T[(size_t)g * N + p] = cost[(size_t)p * n + g];
  ptb_lsap::Ctx cx;
  cx.bc.err = 0;
  return ptb_lsap::hungarian_v2_image(cx, cost, N, n, topk_k, w, row_idx, out);
}
