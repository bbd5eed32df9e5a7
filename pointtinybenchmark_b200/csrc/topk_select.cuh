// Exact per-image top-k by (key desc, index asc): 4-pass 8-bit radix select of the k-th key + ordered tie compaction + bitonic sort
// of <= 4096 (key, index) pairs in shared memory.  Shared by the P2P decode path (p2p.cu) and the RPN proposal path (rpn.cu); the
// kernels are `static` so each translation unit carries its own copy.
#pragma once
#include "ptb_common.cuh"

namespace ptb {

constexpr int TOPK_MAX = 4096;
constexpr int SEL_THREADS = 1024;

static __device__ __forceinline__ void bitonic_sort_u64(unsigned long long* a, int n /*pow2*/) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long x = a[lo], y = a[hi];
        if ((x > y) == up) { a[lo] = y; a[hi] = x; }
      }
    }
  }
  __syncthreads();
}

// one CTA per image: exact top-`topk` of key[b][0..Q) by (key desc, index asc); row b of the output starts at b*out_stride
static __global__ void __launch_bounds__(SEL_THREADS)
p2p_select_kernel(const float* __restrict__ key, int Q, int topk, int32_t* __restrict__ out_idx, int out_stride) {
  __shared__ unsigned long long sel[TOPK_MAX];
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_need, s_cnt_gt, s_eq_base;
  __shared__ unsigned int warp_tot[SEL_THREADS / 32];
  const int b = blockIdx.x;
  const unsigned int* kb = reinterpret_cast<const unsigned int*>(key + (size_t)b * Q);
  const int tid = threadIdx.x;
  if (tid == 0) { s_prefix = 0; s_need = (unsigned)topk; s_cnt_gt = 0; s_eq_base = 0; }
  // ---- radix select (MSB first, 8 bits per pass) of the topk-th largest key
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const unsigned int mask_hi = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < Q; i += SEL_THREADS) {
      const unsigned int v = kb[i];
      if ((v & mask_hi) == prefix) atomicAdd(&hist[(v >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int need = s_need, cum = 0;
      int d = 255;
      for (; d > 0; --d) {
        if (cum + hist[d] >= need) break;
        cum += hist[d];
      }
      s_need = need - cum;                       // how many still to take inside digit d
      s_prefix = prefix | ((unsigned int)d << shift);
    }
    __syncthreads();
  }
  const unsigned int T = s_prefix;               // exact key of the topk-th largest element
  const unsigned int need_eq = s_need;           // number of == T elements to take (lowest indices first)
  // ---- collect: all > T (any order), first need_eq of == T in index order
  for (int base = 0; base < Q; base += SEL_THREADS) {
    const int i = base + tid;
    const unsigned int v = i < Q ? kb[i] : 0u;
    const bool gt = i < Q && v > T;
    const bool eq = i < Q && v == T;
    if (gt) {
      const unsigned int slot = atomicAdd(&s_cnt_gt, 1u);
      sel[slot] = ((unsigned long long)(~v) << 32) | (unsigned int)i;
    }
    // ordered rank among == T
    const unsigned int bal = __ballot_sync(0xffffffffu, eq);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    unsigned int before = s_eq_base;
    for (int w = 0; w < wid; ++w) before += warp_tot[w];
    const unsigned int rank = before + __popc(bal & ((1u << lane) - 1u));
    if (eq && rank < need_eq) sel[(unsigned)topk - need_eq + rank] = ((unsigned long long)(~v) << 32) | (unsigned int)i;
    __syncthreads();
    if (tid == 0) {
      unsigned int t = 0;
      for (int w = 0; w < SEL_THREADS / 32; ++w) t += warp_tot[w];
      s_eq_base += t;
    }
    __syncthreads();
  }
  // (#gt == topk - need_eq by construction, so slots [0,topk-need_eq) and [topk-need_eq, topk) are all filled)
  int n2 = 1;
  while (n2 < topk) n2 <<= 1;
  for (int i = topk + tid; i < n2; i += SEL_THREADS) sel[i] = 0xFFFFFFFFFFFFFFFFull;
  bitonic_sort_u64(sel, n2);
  for (int r = tid; r < topk; r += SEL_THREADS) out_idx[(size_t)b * out_stride + r] = (int32_t)(sel[r] & 0xFFFFFFFFull);
}

}  // namespace ptb
