// Backward of the CPR training loss w.r.t. the logit map, gather-formulated and DETERMINISTIC (round 2; replaces round 1's chain
// mil_bwd -> gfocal_bwd -> bag_gather_bwd (fp32 vector atomics into a zeroed map) -> gfocal_bwd(neg)).
//
//   d loss / d lmap[cell][ch] =   sum over (bag g, sample k, tap t on this cell)  w_t * d loss / d bag_logit[g][k][ch]      (grid_sample backward)
//                               + [ch < N] neg-loss term of the cell itself                                                (cpr_head.py:1219-1228)
//   d/d cls logit = gp * pi * sg (1 - sg)   [+ centre sample: gt-loss term, cpr_head.py:1159-1184]      gp = dLoss/dprob[g][c] (gfocal')
//   d/d ins logit = gp * pi * (sg - p)                                                                 pi = e w / T, e = exp(ins - m)
//   (MILLoss.forward, multi_instance_learning_loss.py:153-203; m, 1/T, p per (bag, class) come from the forward kernel)
//
// One CTA owns a tile of 8 x 8 map cells and ALL channels of it.  It lists the bags of its image whose sample window reaches the
// tile (a handful), lets its warps walk those bags' samples 32 at a time (lane = sample: tap geometry, which taps land in the tile),
// and for every sample with a tap in the tile the 32 lanes turn to the classes: 2 coalesced loads of the sampled logits, the
// per-sample gradient, and one shared-memory atomic per (tap, channel) into a 64-BIT FIXED-POINT accumulator (2^-48 units).
// Integer addition is associative, so the result does not depend on the order in which warps arrive: bit-identical run to run,
// with no global atomics, no zero-initialised gradient map and no materialised (G,K,2N) gradient tensor (740 MB written and re-read
// in round 1).  Samples whose taps straddle a tile border are evaluated by each tile they touch (~1.25x).
#include "ptb_common.cuh"
#include <math_constants.h>

namespace ptb {

constexpr int LB_T = 8;                       // tile side (cells)
constexpr int LB_CELLS = LB_T * LB_T;
constexpr int LB_THREADS = 320;
constexpr int LB_WARPS = LB_THREADS / 32;
constexpr int LB_MAXCAND = 1024;              // GT indices examined per pass
constexpr int LB_NIT = 8;                     // class iterations per lane: up to 256 classes
constexpr float LB_FIX = 281474976710656.f;   // 2^48
constexpr float LB_UNFIX = 1.f / 281474976710656.f;

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gfocal_dp_f(float p, float q, float eps) {      // d/dp of -( (p-q)^2 (q log(p+eps) + (1-q) log(1-p+eps)) )
  const float d = p - q;
  const float L = q * __logf(p + eps) + (1.f - q) * __logf(1.f - p + eps);
  const float dL = __fdividef(q, p + eps) - __fdividef(1.f - q, 1.f - p + eps);
  return -(2.f * d * L + d * d * dL);
}

struct LossBwdArgs {
  const float* bl;          // [G][K][LD] sampled logits (cls at 0, ins at NP)
  const float* weight;      // [G][K]
  const float* mt;          // [G][N][2]  (max ins, 1/T or 0)
  const float* bag_prob;    // [G][N]
  const float* lw;          // [G] label weight (any sample weight > 0)
  const int32_t* labels;    // [G]
  const float* centers;     // [G][2]
  const int32_t* img_ptr;   // [B+1]
  const float* offsets;     // [K][2]
  const float* scale_mil;   // [1] or NULL (no MIL term)
  const float* scale_gt;    // [1] or NULL
  const float* wc;          // [G] validity of the centre sample (gt loss) or NULL
  const float* lmap;        // [B][H][W][LD] logit map (neg term) or NULL
  const uint8_t* neg_mask;  // [B][H][W][N]
  const float* scale_neg;   // [1]
  float* dlmap;             // [B][H][W][LD]
  int H, W, N, NP, LD, K;
  float stride, reach_px, eps;
};

__global__ void __launch_bounds__(LB_THREADS, 2)
cpr_loss_bwd_tile_kernel(const LossBwdArgs a) {
  extern __shared__ unsigned long long acc[];          // [LB_CELLS][LD]
  __shared__ int s_cand[LB_MAXCAND];
  __shared__ int s_ncand;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * LB_T, ty0 = blockIdx.y * LB_T;
  const int H = a.H, W = a.W, N = a.N, NP = a.NP, LD = a.LD, K = a.K;
  const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
  for (int i = tid; i < LB_CELLS * LD; i += LB_THREADS) acc[i] = 0ull;
  const float s_mil = a.scale_mil ? a.scale_mil[0] : 0.f;
  const float s_gt = (a.scale_gt && a.wc) ? a.scale_gt[0] : 0.f;
  const int g_lo = a.img_ptr[b], g_hi = a.img_ptr[b + 1];
  const int n_chunks = (K + 31) / 32;
  const int nit = (N + 31) / 32;

  for (int seg = g_lo; seg < g_hi; seg += LB_MAXCAND) {
    // ---- phase A: bags of this image whose window reaches the tile (order irrelevant: the accumulation is exact integer arithmetic)
    __syncthreads();
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    for (int g = seg + tid; g < min(seg + LB_MAXCAND, g_hi); g += LB_THREADS) {
      const float cx = a.centers[2 * g], cy = a.centers[2 * g + 1];
      const int x_lo = (int)floorf(sample_coord(__fadd_rn(-a.reach_px, cx), a.stride, (float)W, hw));
      const int x_hi = min((int)floorf(sample_coord(__fadd_rn(a.reach_px, cx), a.stride, (float)W, hw)) + 1, W - 1);
      const int y_lo = (int)floorf(sample_coord(__fadd_rn(-a.reach_px, cy), a.stride, (float)H, hh));
      const int y_hi = min((int)floorf(sample_coord(__fadd_rn(a.reach_px, cy), a.stride, (float)H, hh)) + 1, H - 1);
      if (x_hi >= tx0 && x_lo < tx0 + LB_T && y_hi >= ty0 && y_lo < ty0 + LB_T) s_cand[atomicAdd(&s_ncand, 1)] = g;
    }
    __syncthreads();
    const int n_items = s_ncand * n_chunks;
    // ---- phase B: (bag, 32-sample chunk) items, one warp each
    for (int item = warp; item < n_items; item += LB_WARPS) {
      const int g = s_cand[item / n_chunks];
      const int k0 = (item % n_chunks) * 32;
      const int k = k0 + lane;
      const bool act = k < K;
      // lane = sample: tap geometry and which taps land in this tile
      int cell[4] = {-1, -1, -1, -1};
      float tw[4] = {0.f, 0.f, 0.f, 0.f};
      float wk = 0.f;
      if (act) {
        const float px = __fadd_rn(a.offsets[2 * k], a.centers[2 * g]), py = __fadd_rn(a.offsets[2 * k + 1], a.centers[2 * g + 1]);
        const float ix = sample_coord(px, a.stride, (float)W, hw), iy = sample_coord(py, a.stride, (float)H, hh);
        const float x0f = floorf(ix), y0f = floorf(iy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
        const float ex = __fsub_rn(__fadd_rn(x0f, 1.f), ix), wx = __fsub_rn(ix, x0f);
        const float ey = __fsub_rn(__fadd_rn(y0f, 1.f), iy), wy = __fsub_rn(iy, y0f);
        tw[0] = __fmul_rn(ex, ey); tw[1] = __fmul_rn(wx, ey); tw[2] = __fmul_rn(ex, wy); tw[3] = __fmul_rn(wx, wy);
        const int lx0 = x0 - tx0, lx1 = x1 - tx0, ly0 = y0 - ty0, ly1 = y1 - ty0;
        const bool inx0 = (unsigned)lx0 < (unsigned)LB_T, inx1 = (unsigned)lx1 < (unsigned)LB_T;
        const bool iny0 = (unsigned)ly0 < (unsigned)LB_T, iny1 = (unsigned)ly1 < (unsigned)LB_T;
        if (inx0 && iny0) cell[0] = ly0 * LB_T + lx0;
        if (inx1 && iny0) cell[1] = ly0 * LB_T + lx1;
        if (inx0 && iny1) cell[2] = ly1 * LB_T + lx0;
        if (inx1 && iny1) cell[3] = ly1 * LB_T + lx1;
        wk = a.weight[(size_t)g * K + k];
      }
      const bool touches = (cell[0] >= 0) || (cell[1] >= 0) || (cell[2] >= 0) || (cell[3] >= 0);
      unsigned todo = __ballot_sync(0xffffffffu, act && touches);
      if (todo == 0u) continue;
      // per-bag class coefficients of this lane (classes lane, lane + 32, ...)
      const int lab = a.labels[g];
      const float lwg = a.lw[g];
      const float sgt = s_gt != 0.f ? s_gt * a.wc[g] : 0.f;
      float cm[LB_NIT], cinvT[LB_NIT], cp[LB_NIT], cgp[LB_NIT];
#pragma unroll
      for (int i = 0; i < LB_NIT; ++i) {
        const int c = lane + 32 * i;
        cm[i] = 0.f; cinvT[i] = 0.f; cp[i] = 0.f; cgp[i] = 0.f;
        if (i < nit && c < N) {
          const float2 mt = *reinterpret_cast<const float2*>(a.mt + ((size_t)g * N + c) * 2);
          cm[i] = mt.x; cinvT[i] = mt.y;
          cp[i] = a.bag_prob[(size_t)g * N + c];
          cgp[i] = s_mil * lwg * gfocal_dp_f(cp[i], c == lab ? 1.f : 0.f, a.eps);
        }
      }
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int ks = k0 + src;
        int sc[4];
        float sw[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { sc[t] = __shfl_sync(0xffffffffu, cell[t], src); sw[t] = __shfl_sync(0xffffffffu, tw[t], src); }
        const float swk = __shfl_sync(0xffffffffu, wk, src);
        const float* row = a.bl + ((size_t)g * K + ks) * LD;
        const bool centre = (ks == K - 1) && sgt != 0.f;
#pragma unroll
        for (int i = 0; i < LB_NIT; ++i) {
          const int c = lane + 32 * i;
          if (i < nit && c < N) {
            const float xc = __ldg(row + c), xi = __ldg(row + NP + c);
            const float sg = fast_sigmoid(xc);
            const float pi = __expf(xi - cm[i]) * swk * cinvT[i];
            const float gpi = cgp[i] * pi;
            float dc = gpi * sg * (1.f - sg);
            const float di = gpi * (sg - cp[i]);
            if (centre) dc += sgt * gfocal_dp_f(sg, c == lab ? 1.f : 0.f, a.eps) * sg * (1.f - sg);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (sc[t] >= 0) {
                unsigned long long* p = acc + (size_t)sc[t] * LD;
                atomicAdd(p + c, (unsigned long long)__float2ll_rn(sw[t] * dc * LB_FIX));
                atomicAdd(p + NP + c, (unsigned long long)__float2ll_rn(sw[t] * di * LB_FIX));
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- phase D: fixed point -> fp32, + neg-loss term of the cell itself, write the tile (every channel of every in-map cell)
  const float s_neg = (a.lmap && a.scale_neg) ? a.scale_neg[0] : 0.f;
  for (int i = tid; i < LB_CELLS * LD; i += LB_THREADS) {
    const int cl = i / LD, ch = i - cl * LD;
    const int y = ty0 + cl / LB_T, x = tx0 + cl % LB_T;
    if (y >= H || x >= W) continue;
    const size_t cellg = ((size_t)b * H + y) * W + x;
    float v = __ll2float_rn((long long)acc[i]) * LB_UNFIX;
    if (s_neg != 0.f && ch < N && a.neg_mask[cellg * N + ch]) {
      const float sg = fast_sigmoid(a.lmap[cellg * LD + ch]);
      v += s_neg * gfocal_dp_f(sg, 0.f, a.eps) * sg * (1.f - sg);
    }
    a.dlmap[cellg * LD + ch] = v;
  }
}

}  // namespace ptb

using namespace ptb;

extern "C" int ptb_cpr_loss_bwd_map(const float* bag_logits, const float* weight, const float* mil_mt, const float* bag_prob,
                                    const float* label_weight, const int32_t* labels, const float* centers, const int32_t* img_ptr,
                                    const float* offsets, int B, int H, int W, int G, int K, int num_classes, int ins_off, int ld,
                                    float stride, float reach_px, float eps, const float* scale_mil, const float* scale_gt,
                                    const float* valid_center, const float* logit_map, const uint8_t* neg_mask, const float* scale_neg,
                                    float* grad_map, void* stream) {
  PTB_REQUIRE(B > 0 && H > 0 && W > 0 && G >= 0 && K > 0 && num_classes > 0 && num_classes <= 32 * LB_NIT, "shape (num_classes <= 256)");
  PTB_REQUIRE(ld >= ins_off + num_classes && ins_off >= num_classes && stride > 0.f && reach_px >= 0.f, "ld / ins_off / stride");
  PTB_REQUIRE(img_ptr && grad_map && (G == 0 || (bag_logits && weight && mil_mt && bag_prob && label_weight && labels && centers && offsets)),
              "NULL input");
  PTB_REQUIRE(!logit_map || (neg_mask && scale_neg), "the neg term needs logit_map, neg_mask and scale_neg");
  LossBwdArgs a;
  a.bl = bag_logits; a.weight = weight; a.mt = mil_mt; a.bag_prob = bag_prob; a.lw = label_weight; a.labels = labels;
  a.centers = centers; a.img_ptr = img_ptr; a.offsets = offsets; a.scale_mil = scale_mil; a.scale_gt = scale_gt; a.wc = valid_center;
  a.lmap = logit_map; a.neg_mask = neg_mask; a.scale_neg = scale_neg; a.dlmap = grad_map;
  a.H = H; a.W = W; a.N = num_classes; a.NP = ins_off; a.LD = ld; a.K = K; a.stride = stride; a.reach_px = reach_px; a.eps = eps;
  const size_t smem = (size_t)LB_CELLS * ld * sizeof(unsigned long long);
  PTB_REQUIRE(smem <= 200 * 1024, "ld too large for the tile accumulator");
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute(cpr_loss_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail("%s", "ptb_cpr_loss_bwd_map: shared memory opt-in failed");
  dim3 grid((W + LB_T - 1) / LB_T, (H + LB_T - 1) / LB_T, B);
  cpr_loss_bwd_tile_kernel<<<grid, LB_THREADS, smem, (cudaStream_t)stream>>>(a);
  return check_launch("ptb_cpr_loss_bwd_map");
}
