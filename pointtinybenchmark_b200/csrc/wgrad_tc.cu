// Weight gradient of the tower's conv3x3 (256 -> 256, stride 1, pad 1) on the 5th-gen tensor cores, fp32-accurate:
//     dW[co][ci][kh][kw] = sum_{b,h,w} dy[b][h][w][co] * x[b][h+kh-1][w+kw-1][ci]          (autograd of cpr_head.py:1033-1043's convs)
// A GEMM with M = co, N = ci and K = PIXELS: both operands are channels-last activations, i.e. their contiguous dimension is
// M / N, not K.  tcgen05 reads such "MN-major" operands directly (instruction-descriptor bits 15/16, canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units for SWIZZLE_128B, cute/atom/mma_traits_sm100.hpp), so no transpose pass:
//   * a 4-D TMA box {64 channels, 16 w, 2 h, 1 image} of fp16 lands in shared memory as 32 pixel rows of 128 B with the
//     SWIZZLE_128B XOR — exactly one MN-major atom column (64 channels) x 4 K-atoms (8 pixels each, SBO = 1024 B);
//     co 0..127 = 2 such boxes (LBO = 4096 B), ci 0..255 = 4 boxes.  The x box is fetched at the tap-shifted origin; its
//     out-of-bounds part (the conv's zero padding, partial edge tiles) is zero-filled by the TMA unit, and dy's out-of-image
//     rows are zero too, so edge tiles need no masks.
//   * fp32 accuracy as in conv_tc.cu: dy*s1 = h + l and x*s2 = h + l as fp16 pairs, h*h -> main accumulator, l*h + h*l ->
//     correction accumulator (512 TMEM columns), summed in fp32 in the epilogue.
//   * the tensor core adds into the fp32 accumulator with truncation (~0.5 ulp of the accumulator per step, see conv_tc.cu);
//     over the 1100 accumulation steps of a pixel split that is a 3e-4 error on dW (measured vs fp64 at the headline shape).
//     The main accumulator is therefore FLUSHED every WG_FLUSH pixel blocks: the epilogue warps store it as one more fp32
//     partial (store-only: a read-modify-write of the previous partial cost 25 us per flush, ncu; re-measured in round 2 with
//     L2-resident ld.cg/st.cg: +47 us on the kernel for -40 us on the reduction, no gain) and the MMA warp restarts it
//     from zero; the correction accumulator is 2^-11 smaller and runs through.  The reduction kernel adds runs and splits with
//     round-to-nearest fp32 adds in a fixed order.
//   * work: 2 co-halves x 9 taps x S pixel splits = 18*S CTAs (S = 8 -> 144 of 148 SMs), the two co halves being the two CTAs of a
//     tcgen05 cta_group::2 pair (see the kernel); a CTA streams its pixel blocks through a 6 x 32 KB mbarrier ring (4 x 48 KB in the
//     single-CTA debug mode PTB_WGRAD_PAIR=0; warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue) and writes
//     128 x 256 fp32 partials; `wgrad_reduce_kernel` adds the S partials in a fixed order, applies the (power-of-two) inverse
//     operand scales and writes OIHW.  Deterministic.
#include "tc_ptx.cuh"
#include <stdlib.h>

namespace ptb {

constexpr int WG_PX = 32;                         // pixels per K-block: box {64 ch, 16 w, 2 h}
constexpr int WG_TW = 16, WG_TH = 2;
constexpr int WG_STAGES = 4;
constexpr uint32_t WG_BOX_BYTES = WG_PX * 128;    // 4 KB: 32 pixel rows x 64 fp16 channels
constexpr uint32_t WG_A_BYTES = 2 * WG_BOX_BYTES; // 128 co
constexpr uint32_t WG_B_BYTES = 4 * WG_BOX_BYTES; // 256 ci
constexpr uint32_t WG_STAGE_BYTES = 2 * WG_A_BYTES + 2 * WG_B_BYTES;   // 48 KB
constexpr int WG_STAGES_PAIR = 6;                 // CTA-pair mode stages only half of the x tile: 6 x 32 KB
constexpr uint32_t WG_STAGE_BYTES_PAIR = 2 * WG_A_BYTES + WG_B_BYTES;  // 32 KB
constexpr uint32_t WG_RING_BYTES = WG_STAGES * WG_STAGE_BYTES;          // = WG_STAGES_PAIR * WG_STAGE_BYTES_PAIR = 192 KB
static_assert(WG_STAGES_PAIR * WG_STAGE_BYTES_PAIR == WG_RING_BYTES, "both modes share one ring size");
constexpr uint32_t WG_SMEM_BYTES = WG_RING_BYTES + 1024 + 256;
constexpr int WG_THREADS = 192;
constexpr int WG_C = 256;                         // Cout = Cin = 256
constexpr int WG_FLUSH = 128;                     // pixel blocks (256 accumulation steps) between two flushes of the main accumulator

// MN-major, SWIZZLE_128B shared-memory matrix descriptor: atoms of 64 elements (128 B) x 8 K-rows = 1024 B;
// LBO = byte distance between atoms along M/N, SBO = byte distance between atoms along K.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;                          // version = 1 (sm_100)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// kind::f16, fp16 operands, fp32 accumulate, A and B MN-major, M = 128, N = 256
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn_m128_n256() {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// Accumulation runs of a pixel split: the FIRST run of split s is shortened to WG_FLUSH * (s + 1) / splits blocks, later runs have WG_FLUSH
// blocks.  All CTAs stream at the same pace, so equal runs made all 144 of them flush their 128 KB accumulators in the same few
// microseconds (an 18.9 MB store burst at the DRAM write rate, ~7 us with every tensor pipe idle, five times per launch); staggered, the
// splits flush 1/8 of a run apart and the stores of one hide behind the MMAs of the others.
__host__ __device__ inline int wg_first_run(int split, int splits) {
  const int f = (int)(((long long)WG_FLUSH * (split + 1)) / splits);
  return f < 1 ? 1 : f;
}
__host__ __device__ inline int wg_runs(int n_my, int first) {
  if (n_my <= 0) return 0;
  return n_my <= first ? 1 : 1 + (n_my - first + WG_FLUSH - 1) / WG_FLUSH;
}

struct WgradShape {
  int B, H, W;
  int tiles_h, tiles_w, n_blocks;    // pixel blocks of 16 x 2
  int splits;
  int max_runs;                      // partial slots per split (upper bound of wg_runs())
  int taps;                          // 9: conv3x3 (pad 1), 1: conv1x1 / per-cell Linear (CPRHead's cls_out / ins_out logit map)
  int Cout;                          // rows of dW actually wanted (<= 256); rows beyond it are the TMA unit's zero fill of dy
};

// PAIR = true (round 2, default): the two co halves of a (tap, pixel split) form a CTA pair issuing ONE tcgen05.mma.cta_group::2 per
// product (M = 256 = all output channels): each CTA stages its own dy half and only HALF of the x tile (128 input channels), i.e. 32 KB
// instead of 48 KB per pixel block — the single-CTA kernel pulled its operands at the L2 -> SM ceiling (144 CTAs x 48 KB per 768 tensor
// cycles = 11.5 TB/s; ncu: tensor pipe 45 %).  Protocol as in conv_tc.cu: the leader's "full" barrier collects both CTAs' TMA bytes,
// tcgen05.commit.cta_group::2 multicasts "stage free" / "run complete", the peer's epilogue warps release TMEM with remote arrives.
template <bool PAIR>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_dyh, const __grid_constant__ CUtensorMap tm_dyl,
                const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl, WgradShape ws,
                float* __restrict__ partial /*[splits][max_runs][taps][256 co][256 ci]*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int NST = PAIR ? WG_STAGES_PAIR : WG_STAGES;
  constexpr uint32_t STB = PAIR ? WG_STAGE_BYTES_PAIR : WG_STAGE_BYTES;
  constexpr uint32_t BLO = PAIR ? WG_B_BYTES / 2 : WG_B_BYTES;      // offset of the x "lo" half behind the x "hi" half
  const uint32_t bar_base = smem_base + WG_RING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  const uint32_t tfull_bar = bar_base + 128u;
  const uint32_t tempty_bar = bar_base + 136u;
  const uint32_t tmem_slot = bar_base + 160u;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_aligned + WG_RING_BYTES + 160);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit = (split, tap, co half)
  const int unit = blockIdx.x;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;          // PAIR: cluster = (unit 2u, 2u+1) = the two co halves; rank == m_half
  const int m_half = unit & 1;
  const int tap = (unit >> 1) % ws.taps;
  const int split = unit / (2 * ws.taps);
  const int kh = ws.taps == 9 ? tap / 3 : 1, kw = ws.taps == 9 ? tap - (tap / 3) * 3 : 1;
  const int blk0 = (int)(((long long)ws.n_blocks * split) / ws.splits);
  const int blk1 = (int)(((long long)ws.n_blocks * (split + 1)) / ws.splits);
  const int n_my = blk1 - blk0;
  const int first_run = wg_first_run(split, ws.splits);
  const int n_flush = wg_runs(n_my, first_run);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, PAIR ? 8 : 4); // one arrive per epilogue warp (PAIR: of both CTAs, on the leader's barrier)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // The WHOLE warp walks the block list with warp-uniform coordinates (advanced incrementally: no division per block) and lane 0
    // issues: with the loop inside an `if (lane == 0)` the compiler cannot prove the operands uniform and wraps every UTMALDG in an
    // ELECT / R2UR waterfall — ncu's source page showed this single thread ISSUE-bound (~1100 cycles per 32 KB stage, the MMA warp
    // starving at 50 % tensor-pipe activity).
    {
      int stage = 0;
      uint32_t phase = 0;
      const int per_img = ws.tiles_h * ws.tiles_w;
      int b = blk0 / per_img;
      int th = (blk0 - b * per_img) / ws.tiles_w;
      int tw = (blk0 - b * per_img) - th * ws.tiles_w;
      for (int blk = blk0; blk < blk1; ++blk) {
        mbar_wait(empty_bar(stage), phase ^ 1u);               // all lanes wait (measured faster than lane 0 alone waiting inside the branch)
        // shfl-broadcasts: the values are uniform anyway, this is what lets ptxas SEE it (uniform registers feed UTMALDG directly)
        const int ust = __shfl_sync(0xffffffffu, stage, 0);
        const int h0 = __shfl_sync(0xffffffffu, th, 0) * WG_TH, w0 = __shfl_sync(0xffffffffu, tw, 0) * WG_TW;
        const int ub = __shfl_sync(0xffffffffu, b, 0);
        const uint32_t sA_h = smem_base + ust * STB;
        const uint32_t sA_l = sA_h + WG_A_BYTES;
        const uint32_t sB_h = sA_l + WG_A_BYTES;
        const uint32_t sB_l = sB_h + BLO;
        if (PAIR) {
          // my dy half (128 co) + MY half of the x tile (128 ci); every byte is counted on the LEADER's barrier
          const uint32_t lead_full = mapa_rank(full_bar(ust), 0u);
          if (lane == 0) {
            if (rank == 0) mbar_expect_tx(full_bar(ust), 2u * (2 * WG_A_BYTES + WG_B_BYTES));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              tma_load_4d_pair(&tm_dyh, lead_full, sA_h + i * WG_BOX_BYTES, m_half * 128 + 64 * i, w0, h0, ub);
              tma_load_4d_pair(&tm_dyl, lead_full, sA_l + i * WG_BOX_BYTES, m_half * 128 + 64 * i, w0, h0, ub);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              tma_load_4d_pair(&tm_xh, lead_full, sB_h + j * WG_BOX_BYTES, 64 * (2 * (int)rank + j), w0 + kw - 1, h0 + kh - 1, ub);
              tma_load_4d_pair(&tm_xl, lead_full, sB_l + j * WG_BOX_BYTES, 64 * (2 * (int)rank + j), w0 + kw - 1, h0 + kh - 1, ub);
            }
          }
        } else if (lane == 0) {
          mbar_expect_tx(full_bar(ust), WG_STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            tma_load_4d(&tm_dyh, full_bar(ust), sA_h + i * WG_BOX_BYTES, m_half * 128 + 64 * i, w0, h0, ub);
            tma_load_4d(&tm_dyl, full_bar(ust), sA_l + i * WG_BOX_BYTES, m_half * 128 + 64 * i, w0, h0, ub);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            tma_load_4d(&tm_xh, full_bar(ust), sB_h + j * WG_BOX_BYTES, 64 * j, w0 + kw - 1, h0 + kh - 1, ub);
            tma_load_4d(&tm_xl, full_bar(ust), sB_l + j * WG_BOX_BYTES, 64 * j, w0 + kw - 1, h0 + kh - 1, ub);
          }
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
        if (++tw == ws.tiles_w) { tw = 0; if (++th == ws.tiles_h) { th = 0; ++b; } }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0 && n_my > 0 && (!PAIR || rank == 0)) {
      const uint32_t idesc = PAIR ? ((umma_idesc_f16_mn_m128_n256() & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24))
                                  : umma_idesc_f16_mn_m128_n256();
      const uint32_t d_main = tmem_base, d_corr = tmem_base + 256u;
      int stage = 0;
      uint32_t phase = 0;
      int in_run = 0, run_len = first_run, run_idx = 0;        // position inside / length / index of the current accumulation run
      for (int it = 0; it < n_my; ++it) {
        if (in_run == 0 && it > 0) {                           // the epilogue has stored the previous run
          mbar_wait(tempty_bar, (uint32_t)((run_idx - 1) & 1));
          tc_fence_after();
        }
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sA_h = smem_base + stage * STB;
        const uint32_t sA_l = sA_h + WG_A_BYTES;
        const uint32_t sB_h = sA_l + WG_A_BYTES;
        const uint32_t sB_l = sB_h + BLO;
#pragma unroll
        for (int k = 0; k < WG_PX / 16; ++k) {                 // UMMA_K = 16 pixels = two 8-row atoms = 2048 B
          const uint64_t a_h = umma_desc_mn_sw128(sA_h + 2048u * k, WG_BOX_BYTES, 1024u);
          const uint64_t a_l = umma_desc_mn_sw128(sA_l + 2048u * k, WG_BOX_BYTES, 1024u);
          const uint64_t b_h = umma_desc_mn_sw128(sB_h + 2048u * k, WG_BOX_BYTES, 1024u);
          const uint64_t b_l = umma_desc_mn_sw128(sB_l + 2048u * k, WG_BOX_BYTES, 1024u);
          if (PAIR) {
            umma_ss_pair<true>(d_main, a_h, b_h, idesc, (in_run | k) != 0);
            umma_ss_pair<true>(d_corr, a_l, b_h, idesc, (it | k) != 0);
            umma_ss_pair<true>(d_corr, a_h, b_l, idesc, 1u);
          } else {
            umma_ss<true>(d_main, a_h, b_h, idesc, (in_run | k) != 0);
            umma_ss<true>(d_corr, a_l, b_h, idesc, (it | k) != 0);
            umma_ss<true>(d_corr, a_h, b_l, idesc, 1u);
          }
        }
        if (PAIR) umma_commit_pair(empty_bar(stage), (uint16_t)0x3);
        else umma_commit(empty_bar(stage));
        if (++stage == NST) { stage = 0; phase ^= 1u; }
        if (in_run == run_len - 1 || it == n_my - 1) {                              // run complete: hand it to the epilogue(s)
          if (PAIR) umma_commit_pair(tfull_bar, (uint16_t)0x3);
          else umma_commit(tfull_bar);
          in_run = 0; run_len = WG_FLUSH; ++run_idx;
        } else {
          ++in_run;
        }
      }
    }
  } else {
    // =============================== epilogue (warps 2..5) ===============================
    const int q = warp & 3;                                    // TMEM lane quarter this warp may access
    const int co = m_half * 128 + q * 32 + lane;
    float* out0 = partial + ((((size_t)split * ws.max_runs) * ws.taps + tap) * WG_C + co) * WG_C;       // run r: + r * taps*256*256
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int f = 0; f < n_flush; ++f) {
      mbar_wait(tfull_bar, (uint32_t)(f & 1));
      tc_fence_after();
      const bool last = f == n_flush - 1;
      float* out = out0 + (size_t)f * ws.taps * WG_C * WG_C;
#pragma unroll 1
      for (int c = 0; c < WG_C / 32; ++c) {
        uint32_t v[32], vc[32];
        tmem_ld32_nowait(t_lane + (uint32_t)(c * 32), v);
        if (last) tmem_ld32_nowait(t_lane + 256u + (uint32_t)(c * 32), vc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
          if (last) {
            o.x = __fadd_rn(o.x, __uint_as_float(vc[4 * j]));     o.y = __fadd_rn(o.y, __uint_as_float(vc[4 * j + 1]));
            o.z = __fadd_rn(o.z, __uint_as_float(vc[4 * j + 2])); o.w = __fadd_rn(o.w, __uint_as_float(vc[4 * j + 3]));
          }
          __stcg(reinterpret_cast<float4*>(out + c * 32 + 4 * j), o);      // L2-resident: all CTAs flush at once (18.9 MB burst), the reduction re-reads it
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_rank(tempty_bar, 0u));
        else mbar_arrive(tempty_bar);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (PAIR) cluster_sync_all();          // no CTA exits while the peer can still arrive on its barriers / read its operands
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// dw[co][ci][tap] = scale * sum_{split, run} partial[split][run][tap][co][ci]   (fixed order; scale = product of the inverse
// operand scales).  A split owns blocks [n*s/S, n*(s+1)/S) and has wg_runs() runs (0 for an empty split).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, WgradShape ws, float scale, const float* __restrict__ dev_scale_a,
                    const float* __restrict__ dev_scale_b, float* __restrict__ dw, int accumulate) {
  const int i = blockIdx.x * 256 + threadIdx.x;      // over [tap][co][ci]
  if (i >= ws.taps * WG_C * WG_C) return;
  const int ci = i % WG_C, co = (i / WG_C) % WG_C, tap = i / (WG_C * WG_C);
  if (co >= ws.Cout) return;
  const size_t slot = (size_t)ws.taps * WG_C * WG_C;
  float s = 0.f;
  for (int k = 0; k < ws.splits; ++k) {
    const int n_my = (int)(((long long)ws.n_blocks * (k + 1)) / ws.splits) - (int)(((long long)ws.n_blocks * k) / ws.splits);
    const int runs = wg_runs(n_my, wg_first_run(k, ws.splits));
    for (int r = 0; r < runs; ++r) s += __ldcg(partial + ((size_t)k * ws.max_runs + r) * slot + i);
  }
  float sc = scale;
  if (dev_scale_a) sc *= *dev_scale_a;
  if (dev_scale_b) sc *= *dev_scale_b;
  float* o = dw + ((size_t)co * WG_C + ci) * ws.taps + tap;
  *o = accumulate ? *o + s * sc : s * sc;
}

// column sums of a row-major [M][ld] matrix (bias gradient of the logit-map Linear): per-slice partials, then reduce_cols in slice order
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ y, long long M, int N, int ld, int rows_per_slice, float* __restrict__ part /*[slices][N]*/) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const long long m0 = (long long)blockIdx.y * rows_per_slice;
  long long m1 = m0 + rows_per_slice;
  if (m1 > M) m1 = M;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains (fixed association: deterministic)
  long long m = m0;
  for (; m + 3 < m1; m += 4) {
    a0 += y[m * ld + n]; a1 += y[(m + 1) * ld + n]; a2 += y[(m + 2) * ld + n]; a3 += y[(m + 3) * ld + n];
  }
  for (; m < m1; ++m) a0 += y[m * ld + n];
  part[(size_t)blockIdx.y * N + n] = (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ part, int slices, int N, float* __restrict__ out) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int i = 0; i < slices; ++i) s += part[(size_t)i * N + n];
  out[n] = s;
}

static int make_px_map(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C) {
  EncodeTiledFn enc = tc_get_encode();
  if (!enc) return fail("%s", "cuTensorMapEncodeTiled is unavailable (driver too old?)");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, WG_TW, WG_TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(wgrad operand) failed: %s%lld", "", (long long)r);
  return 0;
}

}  // namespace ptb

using namespace ptb;

static int wgrad_splits(int taps) {
  int s = sm_count() / (2 * taps);
  if (s < 1) s = 1;
  if (s > 96) s = 96;
  return s;
}

static int wgrad_max_runs(int n_blocks, int splits) {
  const int per_split = (n_blocks + splits - 1) / splits;
  return 1 + (per_split + WG_FLUSH - 1) / WG_FLUSH;      // upper bound of wg_runs(): a shortened first run + full runs
}

static uint64_t wgrad_ws_bytes(int B, int H, int W, int taps) {
  if (B <= 0 || H <= 0 || W <= 0 || (taps != 1 && taps != 9)) return 0;
  const int n_blocks = B * ((H + WG_TH - 1) / WG_TH) * ((W + WG_TW - 1) / WG_TW);
  const int splits = wgrad_splits(taps);
  return (uint64_t)splits * wgrad_max_runs(n_blocks, splits) * taps * WG_C * WG_C * sizeof(float);
}

extern "C" uint64_t ptb_conv3x3_wgrad_workspace(int B, int H, int W) { return wgrad_ws_bytes(B, H, W, 9); }
extern "C" uint64_t ptb_conv_tc_wgrad_workspace(int B, int H, int W, int taps) { return wgrad_ws_bytes(B, H, W, taps); }

static int wgrad_run(const void* dy_h, const void* dy_l, const void* x_h, const void* x_l, int B, int H, int W, int Cout, int Cin, int taps,
                     float scale, const float* dev_scale_dy, const float* dev_scale_x, void* workspace, float* dw, int accumulate,
                     void* stream, const char* what) {
  CUtensorMap tm_dyh, tm_dyl, tm_xh, tm_xl;
  int rc;
  if ((rc = make_px_map(&tm_dyh, dy_h, B, H, W, Cout))) return rc;
  if ((rc = make_px_map(&tm_dyl, dy_l, B, H, W, Cout))) return rc;
  if ((rc = make_px_map(&tm_xh, x_h, B, H, W, Cin))) return rc;
  if ((rc = make_px_map(&tm_xl, x_l, B, H, W, Cin))) return rc;
  WgradShape ws;
  ws.B = B; ws.H = H; ws.W = W;
  ws.tiles_h = (H + WG_TH - 1) / WG_TH;
  ws.tiles_w = (W + WG_TW - 1) / WG_TW;
  ws.n_blocks = B * ws.tiles_h * ws.tiles_w;
  ws.splits = wgrad_splits(taps);
  ws.max_runs = wgrad_max_runs(ws.n_blocks, ws.splits);
  ws.taps = taps; ws.Cout = Cout;
  // per-device function attribute: set on every call (a process may drive several devices)
  if (cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM_BYTES) != cudaSuccess ||
      cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM_BYTES) != cudaSuccess)
    return fail("%s", "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed for the wgrad kernel");
  cudaStream_t st = (cudaStream_t)stream;
  // CTAs of the upper co half have nothing to do when Cout <= 128: they still run (zero operands) to keep the unit decomposition uniform
  const char* e_pair = getenv("PTB_WGRAD_PAIR");
  if (e_pair && e_pair[0] == '0') {
    wgrad_tc_kernel<false><<<2 * taps * ws.splits, WG_THREADS, WG_SMEM_BYTES, st>>>(tm_dyh, tm_dyl, tm_xh, tm_xl, ws,
                                                                                   reinterpret_cast<float*>(workspace));
  } else {      // CTA pairs: units (2u, 2u+1) = the two co halves of one (tap, split)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * taps * ws.splits);
    cfg.blockDim = dim3(WG_THREADS);
    cfg.dynamicSmemBytes = WG_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, wgrad_tc_kernel<true>, tm_dyh, tm_dyl, tm_xh, tm_xl, ws, reinterpret_cast<float*>(workspace));
    if (e != cudaSuccess) return fail("wgrad: cluster launch failed: %s", cudaGetErrorString(e));
  }
  if ((rc = check_launch(what))) return rc;
  const int n = taps * WG_C * WG_C;
  wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float*>(workspace), ws, scale, dev_scale_dy,
                                                      dev_scale_x, dw, accumulate);
  return check_launch(what);
}

extern "C" int ptb_conv3x3_wgrad_f16x2(const void* dy_h, const void* dy_l, const void* x_h, const void* x_l, int B, int H, int W,
                                       int Cout, int Cin, float scale, const float* dev_scale_dy, const float* dev_scale_x,
                                       void* workspace, float* dw, int accumulate, void* stream) {
  PTB_REQUIRE(B > 0 && H > 0 && W > 0, "shape");
  PTB_REQUIRE(Cout == WG_C && Cin == WG_C, "the tensor-core weight gradient covers the head's 256 -> 256 convolutions");
  PTB_REQUIRE(dy_h && dy_l && x_h && x_l && workspace && dw, "NULL input");
  PTB_REQUIRE(((uintptr_t)dy_h % 16 == 0) && ((uintptr_t)dy_l % 16 == 0) && ((uintptr_t)x_h % 16 == 0) && ((uintptr_t)x_l % 16 == 0) &&
                  ((uintptr_t)workspace % 16 == 0), "16-byte alignment");
  return wgrad_run(dy_h, dy_l, x_h, x_l, B, H, W, Cout, Cin, 9, scale, dev_scale_dy, dev_scale_x, workspace, dw, accumulate, stream,
                   "ptb_conv3x3_wgrad_f16x2");
}

extern "C" int ptb_conv_tc_wgrad_f16x2(const void* dy_h, const void* dy_l, const void* x_h, const void* x_l, int B, int H, int W,
                                       int Cout, int Cin, int taps, float scale, const float* dev_scale_dy, const float* dev_scale_x,
                                       void* workspace, float* dw, int accumulate, void* stream) {
  PTB_REQUIRE(B > 0 && H > 0 && W > 0 && (taps == 1 || taps == 9), "shape");
  PTB_REQUIRE(Cin == WG_C && Cout > 0 && Cout <= WG_C && Cout % 8 == 0, "Cin must be 256, Cout a multiple of 8 up to 256");
  PTB_REQUIRE(dy_h && dy_l && x_h && x_l && workspace && dw, "NULL input");
  PTB_REQUIRE(((uintptr_t)dy_h % 16 == 0) && ((uintptr_t)dy_l % 16 == 0) && ((uintptr_t)x_h % 16 == 0) && ((uintptr_t)x_l % 16 == 0) &&
                  ((uintptr_t)workspace % 16 == 0), "16-byte alignment");
  return wgrad_run(dy_h, dy_l, x_h, x_l, B, H, W, Cout, Cin, taps, scale, dev_scale_dy, dev_scale_x, workspace, dw, accumulate, stream,
                   "ptb_conv_tc_wgrad_f16x2");
}

extern "C" uint64_t ptb_col_sum_workspace(int64_t M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return (uint64_t)148 * 4 * N * sizeof(float);
}

extern "C" int ptb_col_sum(const float* y, int64_t M, int N, int ld, float* workspace, float* out, void* stream) {
  PTB_REQUIRE(M > 0 && N > 0 && ld >= N, "shape");
  PTB_REQUIRE(y && workspace && out, "NULL input");
  int slices = 148 * 4;
  if ((int64_t)slices > M) slices = (int)M;
  const int rps = (int)((M + slices - 1) / slices);
  slices = (int)((M + rps - 1) / rps);
  cudaStream_t st = (cudaStream_t)stream;
  colsum_partial_kernel<<<dim3((N + 255) / 256, slices), 256, 0, st>>>(y, M, N, ld, rps, workspace);
  int rc = check_launch("ptb_col_sum/partial");
  if (rc) return rc;
  colsum_reduce_kernel<<<(N + 255) / 256, 256, 0, st>>>(workspace, slices, N, out);
  return check_launch("ptb_col_sum");
}
