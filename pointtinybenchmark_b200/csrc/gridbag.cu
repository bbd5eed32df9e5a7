// Grid-cell bags ("neighbour index" bags): GridPtFeatGenerator.generate with GridCirclesPtFeatGenerator.get_chosen_neighbours
// (reference cpr_head.py:296-350, 418-444).  For every GT centre the bag is
//     [ the centres of the grid cells within radius*stride of the GT, in row-major order | zero padding | the GT centre ]
// with EXACT copies of the map's cell vectors for the cells and a bilinear sample (grid_sample, align_corners=False, border)
// for the centre.  The reference builds it with one boolean-mask indexing per GT in a Python loop; here one CTA per GT:
// warp 0 walks a conservative window around the GT in row-major order and compacts the chosen cells with ballots (stable,
// so the order is the reference's), then all warps copy the cell vectors with 128-bit loads/stores.
//
// Arithmetic of the neighbour test (bit-exact target): cell centre = float(j*stride) + stride/2 (cpr_head.py:240-244),
// d = torch.norm([px-cx, py-cy]) which ATen evaluates as sqrt(fma(dy, dy, dx*dx)) (checked on 8.4e6 pairs in the build
// container, 0 mismatches; the other three candidate orders mismatch), chosen = d <= radius*stride.
#include "ptb_common.cuh"

namespace ptb {

constexpr int GB_THREADS = 128;
constexpr int GB_MAX_SLOTS = 4096;      // cells per bag held in shared memory

__device__ __forceinline__ float cell_coord(int j, float stride) { return __fadd_rn(__fmul_rn((float)j, stride), 0.5f * stride); }

__global__ void __launch_bounds__(GB_THREADS)
grid_bag_kernel(const float* __restrict__ map, int H, int W, int C, int ld, const float* __restrict__ centers,
                const int32_t* __restrict__ bag_img, float stride, float radius_px, int cap /*cell slots*/,
                float* __restrict__ out_feats, float* __restrict__ out_pts, uint8_t* __restrict__ out_valid,
                int32_t* __restrict__ out_cell, int32_t* __restrict__ overflow) {
  extern __shared__ int32_t s_cell[];   // [cap]
  __shared__ int s_count;
  const int g = blockIdx.x;
  const int b = bag_img[g];
  const float cx = centers[2 * g], cy = centers[2 * g + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Kt = cap + 1;

  if (warp == 0) {
    // conservative window: every chosen cell satisfies |px - cx| <= radius_px, i.e. j within (cx -+ radius_px)/stride - 0.5
    const int j0 = max(0, (int)floorf((cx - radius_px) / stride - 0.5f) - 1);
    const int j1 = min(W - 1, (int)floorf((cx + radius_px) / stride - 0.5f) + 2);
    const int i0 = max(0, (int)floorf((cy - radius_px) / stride - 0.5f) - 1);
    const int i1 = min(H - 1, (int)floorf((cy + radius_px) / stride - 0.5f) + 2);
    const int ww = j1 - j0 + 1, wh = i1 - i0 + 1;
    const int total = (ww > 0 && wh > 0) ? ww * wh : 0;
    int count = 0;
    for (int base = 0; base < total; base += 32) {
      const int q = base + lane;
      bool chosen = false;
      int cell = -1;
      if (q < total) {
        const int i = i0 + q / ww, j = j0 + q % ww;
        const float dx = __fsub_rn(cell_coord(j, stride), cx), dy = __fsub_rn(cell_coord(i, stride), cy);
        const float d = __fsqrt_rn(__fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
        chosen = d <= radius_px;
        cell = i * W + j;
      }
      const unsigned m = __ballot_sync(0xffffffffu, chosen);
      if (chosen) {
        const int slot = count + __popc(m & ((1u << lane) - 1u));
        if (slot < cap) s_cell[slot] = cell;
      }
      count += __popc(m);
    }
    if (lane == 0) {
      if (count > cap) { if (overflow) *overflow = 1; count = cap; }     // the reference raises here (cpr_head.py:334)
      s_count = count;
    }
  }
  __syncthreads();
  const int count = s_count;
  const size_t img_base = (size_t)b * H * W;
  // ---- per-slot scalars
  for (int s = threadIdx.x; s < Kt; s += GB_THREADS) {
    const size_t o = (size_t)g * Kt + s;
    float px = 0.f, py = 0.f, ps = 0.f;
    int cell = -1;
    bool v = false;
    if (s == Kt - 1) { px = cx; py = cy; ps = stride; cell = -2; v = true; }
    else if (s < count) {
      cell = s_cell[s];
      px = cell_coord(cell % W, stride); py = cell_coord(cell / W, stride); ps = stride; v = true;
    }
    if (out_pts) { out_pts[3 * o] = px; out_pts[3 * o + 1] = py; out_pts[3 * o + 2] = ps; }
    if (out_valid) out_valid[o] = v;
    if (out_cell) out_cell[o] = cell;
  }
  if (!out_feats) return;
  // ---- cell vectors: exact copies (or zeros for the padding)
  const int CG = C >> 2;
  float4* dst = reinterpret_cast<float4*>(out_feats + (size_t)g * Kt * C);
  const int n_vec = cap * CG;
  for (int idx = threadIdx.x; idx < n_vec; idx += GB_THREADS) {
    const int s = idx / CG, c4 = idx - s * CG;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < count) v = __ldg(reinterpret_cast<const float4*>(map + (img_base + s_cell[s]) * ld) + c4);
    __stcs(dst + idx, v);
  }
  // ---- the GT centre: bilinear sample (same FMA chain as ptb_cpr_bag_gather = ATen's CPU grid_sample)
  const Taps tp = make_taps(cx, cy, stride, H, W);
  const float4* b00 = reinterpret_cast<const float4*>(map + (img_base + tp.o00) * ld);
  const float4* b01 = reinterpret_cast<const float4*>(map + (img_base + tp.o01) * ld);
  const float4* b10 = reinterpret_cast<const float4*>(map + (img_base + tp.o10) * ld);
  const float4* b11 = reinterpret_cast<const float4*>(map + (img_base + tp.o11) * ld);
  for (int c4 = threadIdx.x; c4 < CG; c4 += GB_THREADS) {
    const float4 q0 = __ldg(b00 + c4), q1 = __ldg(b01 + c4), q2 = __ldg(b10 + c4), q3 = __ldg(b11 + c4);
    const float4 r = bilerp4(q0, q1, q2, q3, tp.w00, tp.w01, tp.w10, tp.w11);
    __stcs(dst + (size_t)cap * CG + c4, r);
  }
}

// backward: grad_map[b][cell][c] += grad_out[g][slot][c] for the copied cells, bilinear scatter for the centre slot.
// one warp per (g, slot); vector atomics (red.global.add.v4.f32).  Summation order across bags is not fixed (atomics),
// like the reference's index_put / grid_sample backward on CUDA.
__global__ void __launch_bounds__(256)
grid_bag_bwd_kernel(const float* __restrict__ grad_out, int H, int W, int C, int ld, const float* __restrict__ centers,
                    const int32_t* __restrict__ bag_img, const int32_t* __restrict__ cell_idx, long long S, int Kt,
                    float stride, float* __restrict__ grad_map) {
  const int CG = C >> 2;
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = warp_global; s < S; s += n_warps) {
    const int cell = cell_idx[s];
    if (cell == -1) continue;
    const int g = (int)(s / Kt);
    const size_t img_base = (size_t)bag_img[g] * H * W;
    const float4* go = reinterpret_cast<const float4*>(grad_out + (size_t)s * C);
    if (cell >= 0) {
      float* base = grad_map + (img_base + cell) * ld;
      for (int c4 = lane; c4 < CG; c4 += 32) atomicAdd(reinterpret_cast<float4*>(base) + c4, __ldcs(go + c4));
    } else {
      const Taps tp = make_taps(centers[2 * g], centers[2 * g + 1], stride, H, W);
      const int offs[4] = {tp.o00, tp.o01, tp.o10, tp.o11};
      const float ws[4] = {tp.w00, tp.w01, tp.w10, tp.w11};
      for (int c4 = lane; c4 < CG; c4 += 32) {
        const float4 gq = __ldcs(go + c4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (ws[q] != 0.f) {
            const float4 v = make_float4(gq.x * ws[q], gq.y * ws[q], gq.z * ws[q], gq.w * ws[q]);
            atomicAdd(reinterpret_cast<float4*>(grad_map + (img_base + offs[q]) * ld) + c4, v);
          }
        }
      }
    }
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" int ptb_cpr_grid_bag(const float* map, int B, int H, int W, int C, int ld, const float* centers,
                                const int32_t* bag_img, int G, float stride, float radius_px, int max_pos_num,
                                float* out_feats, float* out_pts, uint8_t* out_valid, int32_t* out_cell, int32_t* overflow,
                                void* stream) {
  PTB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && G >= 0 && max_pos_num > 0, "shape");
  PTB_REQUIRE(stride > 0.f && radius_px >= 0.f, "stride / radius");
  PTB_REQUIRE(max_pos_num + 1 <= GB_MAX_SLOTS, "max_pos_num too large for shared memory (4095)");
  PTB_REQUIRE(!out_feats || (map && C % 4 == 0 && ld % 4 == 0 && ld >= C), "C and ld must be multiples of 4, ld >= C");
  PTB_REQUIRE(((uintptr_t)map % 16 == 0) && ((uintptr_t)out_feats % 16 == 0), "map/out_feats must be 16-byte aligned");
  if (G == 0) return 0;
  PTB_REQUIRE(centers && bag_img, "NULL input");
  const int cap = max_pos_num + 1;         // the reference allocates max_pos_num + num_refine cell slots (cpr_head.py:326)
  grid_bag_kernel<<<G, GB_THREADS, (size_t)cap * sizeof(int32_t), (cudaStream_t)stream>>>(
      map, H, W, C, ld, centers, bag_img, stride, radius_px, cap, out_feats, out_pts, out_valid, out_cell, overflow);
  return check_launch("ptb_cpr_grid_bag");
}

extern "C" int ptb_cpr_grid_bag_bwd(const float* grad_out, int B, int H, int W, int C, int ld, const float* centers,
                                    const int32_t* bag_img, const int32_t* cell_idx, int G, int Kt, float stride,
                                    float* grad_map, void* stream) {
  PTB_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && G >= 0 && Kt > 0, "shape");
  PTB_REQUIRE(C % 4 == 0 && ld % 4 == 0 && ld >= C, "C and ld must be multiples of 4, ld >= C");
  PTB_REQUIRE(((uintptr_t)grad_out % 16 == 0) && ((uintptr_t)grad_map % 16 == 0), "16-byte alignment");
  if (G == 0) return 0;
  PTB_REQUIRE(grad_out && centers && bag_img && cell_idx && grad_map, "NULL input");
  const long long S = (long long)G * Kt;
  long long blocks = (S + 7) / 8;
  const long long max_blocks = (long long)sm_count() * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  grid_bag_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_out, H, W, C, ld, centers, bag_img, cell_idx, S, Kt,
                                                                         stride, grad_map);
  return check_launch("ptb_cpr_grid_bag_bwd");
}
