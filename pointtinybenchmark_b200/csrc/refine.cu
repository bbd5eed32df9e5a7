// Point refinement (PointRefiner.refine_single, cpr_head.py:780-850) — stage form and fused form.
//
// One CTA (4 warps) per GT.  A warp owns one bag sample at a time:
//   * lanes span the classes (128-bit loads of channels-last logits / probabilities) -> argmax for the classify
//     filter (cpr_head.py:745-756) by shuffle reduction, prob of the GT's label by broadcast;
//   * lanes span the same-(image,label) GT group for the nearest filter (cpr_head.py:711-743), distances in
//     torch.cdist's fp32 matmul formulation, first-index argmin by shuffle reduction;
//   * thresholds (cpr_head.py:823) and inside-image (773-778) are scalar;
// then the CTA reduces the surviving samples: weights, weighted mean, score, not_refine fallback (829-838).
// The fused form never materialises the (G,K,classes) probability tensor: it bilinearly samples the class-logit map
// (linearity: Linear(bilinear(feat)) == bilinear(Linear(feat)), border padding keeps the 4 weights summing to 1).
#include "ptb_common.cuh"
#include <math_constants.h>

namespace ptb {

constexpr int RF_THREADS = 128;          // stage kernel
constexpr int RF_WARPS = RF_THREADS / 32;
constexpr int RF_MAXWARPS = 32;

struct ArgMin {
  float d;
  int i;
};
__device__ __forceinline__ ArgMin warp_argmin_first(ArgMin a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, a.d, o);
    const int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
    if (od < a.d || (od == a.d && oi < a.i)) { a.d = od; a.i = oi; }
  }
  return a;
}
// argmax with lowest index on ties (torch.max(dim) on CPU keeps the first maximum)
__device__ __forceinline__ ArgMin warp_argmax_first(ArgMin a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, a.d, o);
    const int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
    if (od > a.d || (od == a.d && oi < a.i)) { a.d = od; a.i = oi; }
  }
  return a;
}

__device__ __forceinline__ float block_sum(float v, float* red /*[RF_MAXWARPS]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  const int nw = (int)(blockDim.x >> 5);
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = red[0];
  const int nw = (int)(blockDim.x >> 5);
  for (int w = 1; w < nw; ++w) s = fmaxf(s, red[w]);
  return s;
}

// nearest filter for one sample, evaluated by a full warp.  Candidates = (member j, refine r) centres in group order.
//   cand(j, r, &cx, &cy) supplies centre coordinates.  Returns true iff argmin == own.
template <class CandFn>
__device__ __forceinline__ bool nearest_is_own(float px, float py, int t, int R, int own, bool use_mm, CandFn cand, int lane) {
  const float pn = sq_norm2(px, py);
  ArgMin best;
  best.d = CUDART_INF_F;
  best.i = 0x7fffffff;
  const int total = t * R;
  for (int q = lane; q < total; q += 32) {
    float cx, cy;
    cand(q / R, q % R, cx, cy);
    const float d = use_mm ? cdist_mm(px, py, pn, cx, cy, sq_norm2(cx, cy)) : cdist_direct(px, py, cx, cy);
    if (d < best.d) { best.d = d; best.i = q; }   // ascending q per lane: keeps the first minimum
  }
  best = warp_argmin_first(best);
  return best.i == own;
}

// tail shared by both forms: pm/x/y in shared memory
__device__ __forceinline__ void refine_tail(const float* pm, const float* sx, const float* sy, int Kt, float gt_x, float gt_y,
                                            const uint8_t* not_refine_in, int g, const ptb_refine_cfg& cfg, float* red,
                                            float* out_pts, float* out_score, uint8_t* out_not_refine, uint8_t* out_chosen) {
  const int tid = threadIdx.x;
  float s = 0.f, c = 0.f, mx = 0.f;
  for (int k = tid; k < Kt; k += blockDim.x) {
    const float v = pm[k];
    s += v;
    c += (v > 0.f) ? 1.f : 0.f;
    mx = fmaxf(mx, v);
  }
  const float sum = block_sum(s, red);
  const float cnt = block_sum(c, red);
  const float denom = __fadd_rn(sum, 1e-8f);        // cpr_head.py:832
  float ax = 0.f, ay = 0.f;
  for (int k = tid; k < Kt; k += blockDim.x) {
    const float w = __fdiv_rn(pm[k], denom);
    ax += __fmul_rn(sx[k], w);
    ay += __fmul_rn(sy[k], w);
    if (out_chosen) out_chosen[(size_t)g * Kt + k] = w > 0.f;     // cpr_head.py:849
  }
  const float rx = block_sum(ax, red);
  const float ry = block_sum(ay, red);
  float score = __fdiv_rn(sum, __fadd_rn(cnt, 1e-8f));   // cpr_head.py:835
  bool nr = score < cfg.refine_th;                        // cpr_head.py:836
  if (not_refine_in) nr = nr || (not_refine_in[g] != 0);
  if (cfg.flags & 4) {                                    // return_score_type == 'max' (cpr_head.py:840-842)
    const float m = block_max(mx, red);
    score = (m == 0.f) ? __fmul_rn(cfg.refine_th, 0.5f) : m;
  }
  if (tid == 0) {
    out_pts[2 * g] = nr ? gt_x : rx;
    out_pts[2 * g + 1] = nr ? gt_y : ry;
    out_score[g] = score;
    out_not_refine[g] = nr;
  }
}

// ------------------------------------------------------------------------------------------------
// stage form: probabilities given
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RF_THREADS)
refine_stage_kernel(const float* __restrict__ prob, const float* __restrict__ pts, const uint8_t* __restrict__ valid,
                    int Kt, int K, int ncls, const int32_t* __restrict__ labels, const int32_t* __restrict__ bag_img,
                    const int32_t* __restrict__ img_hw, const int32_t* __restrict__ grp_of,
                    const int32_t* __restrict__ grp_ptr, const int32_t* __restrict__ grp_idx,
                    const uint8_t* __restrict__ not_refine_in, ptb_refine_cfg cfg, float* __restrict__ out_pts,
                    float* __restrict__ out_score, uint8_t* __restrict__ out_not_refine, uint8_t* __restrict__ out_chosen,
                    uint8_t* __restrict__ out_merge_valid) {
  extern __shared__ float sm[];
  float* pm = sm;            // [Kt]
  float* sx = sm + Kt;       // [Kt]
  float* sy = sm + 2 * Kt;   // [Kt]
  __shared__ float red[RF_MAXWARPS];
  const int g = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = labels[g], b = bag_img[g];
  const int R = Kt / K;
  const float ih = (float)img_hw[2 * b], iw = (float)img_hw[2 * b + 1];
  const int gi = grp_of[g];
  const int m0 = grp_ptr[gi], t = grp_ptr[gi + 1] - m0;
  int pos = 0;
  for (int j = 0; j < t; ++j) if (grp_idx[m0 + j] == g) pos = j;
  const bool use_mm = ((long long)t * R * K > 25) || (t * R > 25);
  const float* gprob = prob + (size_t)g * Kt * ncls;
  const float pg = gprob[(size_t)(K - 1) * ncls + l];           // centre of refine 0 (cpr_head.py:807,821)
  const float pg_a = __fmul_rn(pg, cfg.gt_alpha);

  for (int s = warp; s < Kt; s += RF_WARPS) {
    const float* pr = gprob + (size_t)s * ncls;
    const float px = pts[((size_t)g * Kt + s) * 3], py = pts[((size_t)g * Kt + s) * 3 + 1];
    bool m = valid[(size_t)g * Kt + s] != 0;
    if (cfg.flags & 2) {
      ArgMin a;
      a.d = -CUDART_INF_F;
      a.i = 0x7fffffff;
      for (int c = lane; c < ncls; c += 32) {
        const float v = pr[c];
        if (v > a.d) { a.d = v; a.i = c; }
      }
      a = warp_argmax_first(a);
      m = m && (a.i == l);
    }
    if ((cfg.flags & 1) && t > 1) {
      auto cand = [&](int j, int r, float& cx, float& cy) {
        const size_t o = ((size_t)grp_idx[m0 + j] * Kt + (size_t)r * K + (K - 1)) * 3;
        cx = pts[o]; cy = pts[o + 1];
      };
      m = m && nearest_is_own(px, py, t, R, pos * R + s / K, use_mm, cand, lane);
    }
    const float p = pr[l];
    m = m && (p > cfg.merge_th) && (p > pg_a);
    m = m && (px < iw) && (px >= 0.f) && (py < ih) && (py >= 0.f);
    if (lane == 0) {
      pm[s] = m ? p : 0.f;      // bag_cls_prob * merge_valid.float()   (p * 1.0 or p * 0.0)
      sx[s] = px; sy[s] = py;
      if (out_merge_valid) out_merge_valid[(size_t)g * Kt + s] = m;
    }
  }
  __syncthreads();
  const float gx = pts[((size_t)g * Kt + (K - 1)) * 3], gy = pts[((size_t)g * Kt + (K - 1)) * 3 + 1];
  refine_tail(pm, sx, sy, Kt, gx, gy, not_refine_in, g, cfg, red, out_pts, out_score, out_not_refine, out_chosen);
}

// ------------------------------------------------------------------------------------------------
// fused form: logits sampled from the map on the fly (num_refine == 1).
// One CTA per GT; a warp takes 32 samples at a time and alternates between two lane mappings:
//   "owner" phases (lane = sample): coordinates, validity, nearest filter, sigmoid + thresholds   -> done once per sample;
//   "class" phase (8 lanes per sample, 4 samples per round, 8 rounds): the 8 lanes read 128 contiguous bytes of each of the
//     4 bilinear taps per step (channels-last logit map: one L1 line per tap instead of one line per lane), every lane keeps
//     the arg-max / runner-up of ITS classes on the LOGITS, a 3-step xor-shuffle merges them and the owner lane picks them up.
// ncu of the previous mappings: one thread per sample was L1-bound (32 lines per load instruction), 8 lanes per sample for
// everything was issue-bound (the per-sample scalar work replicated 8x: 770 SASS instructions per 4 samples).
// The map is L1/L2 resident: the 289 samples of a bag share an 18x18-cell window.
// ------------------------------------------------------------------------------------------------
constexpr int RF_SUB = 8;                 // lanes per sample in the class phase
constexpr int RF_SPW = 32 / RF_SUB;       // samples per round

__global__ void __launch_bounds__(1024)
refine_fused_kernel(const float* __restrict__ lmap, int H, int W, int ncls, int ld, const float* __restrict__ centers,
                    const int32_t* __restrict__ labels, const int32_t* __restrict__ bag_img,
                    const float* __restrict__ offsets, int K, float stride, const int32_t* __restrict__ pad_hw,
                    const int32_t* __restrict__ img_hw, const int32_t* __restrict__ grp_of,
                    const int32_t* __restrict__ grp_ptr, const int32_t* __restrict__ grp_idx,
                    const uint8_t* __restrict__ not_refine_in, ptb_refine_cfg cfg, float* __restrict__ out_pts,
                    float* __restrict__ out_score, uint8_t* __restrict__ out_not_refine, uint8_t* __restrict__ out_chosen) {
  extern __shared__ float sm[];
  float* pm = sm;
  float* sx = sm + K;
  float* sy = sm + 2 * K;
  float* pl = sm + 3 * K;     // prob of the GT label per sample (thresholds need the centre's first)
  uint8_t* mk = reinterpret_cast<uint8_t*>(sm + 4 * K);   // partial mask per sample
  __shared__ float red[RF_MAXWARPS];
  const int g = blockIdx.x;
  const int l = labels[g], b = bag_img[g];
  const float ih = (float)img_hw[2 * b], iw = (float)img_hw[2 * b + 1];
  const float ph = (float)pad_hw[2 * b], pw = (float)pad_hw[2 * b + 1];
  const int gi = grp_of[g];
  const int m0 = grp_ptr[gi], t = grp_ptr[gi + 1] - m0;
  const bool use_mm = ((long long)t * K > 25) || (t > 25);
  const float cxg = centers[2 * g], cyg = centers[2 * g + 1];
  const float ox_last = offsets[2 * (K - 1)], oy_last = offsets[2 * (K - 1) + 1];
  const float* img_map = lmap + (size_t)b * H * W * ld;
  const int cg4 = (ncls + 3) >> 2;
  const int lane = threadIdx.x & 31, sub = lane & (RF_SUB - 1), slot = lane / RF_SUB;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int l4 = l >> 2, lq = l & 3;                  // float4 / component that holds the label's logit
  const int l_sub = l4 & (RF_SUB - 1);                // ... and the sub-lane that reads it
  const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;

  for (int s0 = warp * 32; s0 < K; s0 += nwarps * 32) {              // warp-uniform trip count
    // ---- owner phase 1: lane = sample
    const bool act = s0 + lane < K;
    const int s = act ? s0 + lane : K - 1;                            // idle lanes shadow the centre sample
    const float px = __fadd_rn(offsets[2 * s], cxg), py = __fadd_rn(offsets[2 * s + 1], cyg);
    const float ix = sample_coord(px, stride, (float)W, hw), iy = sample_coord(py, stride, (float)H, hh);
    // ---- class phase
    float my_best = 0.f, my_runner = 0.f, my_llab = 0.f;
    int my_besti = 0;
#pragma unroll 1
    for (int r = 0; r < 32 / RF_SPW; ++r) {
      const int src = RF_SPW * r + slot;
      const Taps tp = make_taps_at(__shfl_sync(0xffffffffu, ix, src), __shfl_sync(0xffffffffu, iy, src), H, W);
      const float4* b00 = reinterpret_cast<const float4*>(img_map + (size_t)tp.o00 * ld);
      const float4* b01 = reinterpret_cast<const float4*>(img_map + (size_t)tp.o01 * ld);
      const float4* b10 = reinterpret_cast<const float4*>(img_map + (size_t)tp.o10 * ld);
      const float4* b11 = reinterpret_cast<const float4*>(img_map + (size_t)tp.o11 * ld);
      float best = -CUDART_INF_F, runner = -CUDART_INF_F, llab = 0.f;
      int besti = 0x7fffffff;
      for (int c4 = sub; c4 < cg4; c4 += RF_SUB) {
        const float4 q0 = __ldg(b00 + c4);
        const float4 q1 = __ldg(b01 + c4);
        const float4 q2 = __ldg(b10 + c4);
        const float4 q3 = __ldg(b11 + c4);
        const float4 lg4 = bilerp4(q0, q1, q2, q3, tp.w00, tp.w01, tp.w10, tp.w11);
        float lg[4] = {lg4.x, lg4.y, lg4.z, lg4.w};
        if (c4 == l4) llab = lq == 0 ? lg[0] : lq == 1 ? lg[1] : lq == 2 ? lg[2] : lg[3];
        if (4 * c4 + 3 >= ncls) {                                       // row padding beyond num_classes never competes
#pragma unroll
          for (int q = 1; q < 4; ++q)
            if (4 * c4 + q >= ncls) lg[q] = -CUDART_INF_F;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                   // ascending class inside a lane: first maximum wins
          const float v = lg[q];
          runner = fmaxf(runner, fminf(best, v));
          besti = v > best ? 4 * c4 + q : besti;
          best = fmaxf(best, v);
        }
      }
      // merge the 8 lanes of the sample: highest logit, ties -> lowest class; runner = best logit of all OTHER classes
#pragma unroll
      for (int d = 1; d < RF_SUB; d <<= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, d);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, d);
        const float orun = __shfl_xor_sync(0xffffffffu, runner, d);
        const bool take = ob > best || (ob == best && oi < besti);
        runner = fmaxf(fmaxf(runner, orun), take ? best : ob);
        best = take ? ob : best;
        besti = take ? oi : besti;
      }
      // hand the result to the owner lane (lanes 4r .. 4r+3 own the samples of this round)
      const int from = (lane & (RF_SPW - 1)) * RF_SUB;
      const float tb = __shfl_sync(0xffffffffu, best, from);
      const int ti = __shfl_sync(0xffffffffu, besti, from);
      const float tr = __shfl_sync(0xffffffffu, runner, from);
      const float tl = __shfl_sync(0xffffffffu, llab, from + l_sub);
      if ((lane / RF_SPW) == r) { my_best = tb; my_besti = ti; my_runner = tr; my_llab = tl; }
    }
    // ---- owner phase 2: lane = sample.
    // The reference takes the FIRST maximum of the PROBABILITIES (cpr_head.py:745-756): a class c < besti whose logit is a
    // hair below the best can round to the same fp32 sigmoid (always when both saturate to 1.0).  t0 bounds that region
    // from below with 8x slack (ulp(p) / (p(1-p)) in logit units); only if another class reaches it are sigmoids compared.
    const float pmax = sigmoidf_acc(my_best);
    int besti = my_besti;
    if (cfg.flags & 2) {
      float t0;
      if (my_best <= 0.f) t0 = my_best - 2e-6f;
      else { const float q1m = __fsub_rn(1.f, pmax); t0 = q1m > 0.f ? my_best - __fdiv_rn(1e-6f, q1m) : 16.f; }
      if (my_runner >= t0) {                                           // rare (divergent, serial over the classes)
        const Taps tp = make_taps_at(ix, iy, H, W);
        const float* r00 = img_map + (size_t)tp.o00 * ld;
        const float* r01 = img_map + (size_t)tp.o01 * ld;
        const float* r10 = img_map + (size_t)tp.o10 * ld;
        const float* r11 = img_map + (size_t)tp.o11 * ld;
        for (int c = 0; c < besti; ++c) {
          const float v = __fmaf_rn(__ldg(r11 + c), tp.w11, __fmaf_rn(__ldg(r10 + c), tp.w10,
                          __fmaf_rn(__ldg(r01 + c), tp.w01, __fmul_rn(__ldg(r00 + c), tp.w00))));
          if (v >= t0 && sigmoidf_acc(v) >= pmax) { besti = c; break; }
        }
      }
    }
    const float p_label = (l == besti) ? pmax : sigmoidf_acc(my_llab);
    bool m = (0.f <= px) && (px < pw) && (0.f <= py) && (py < ph);   // bag_valid (cpr_head.py:179)
    if (cfg.flags & 2) m = m && (besti == l);
    if ((cfg.flags & 1) && t > 1) {
      // nearest filter (cpr_head.py:711-743): candidates = same-(image,label) GT centres in ascending GT order
      const float pn = sq_norm2(px, py);
      float bd = CUDART_INF_F;
      int bj = -1;
      for (int j = 0; j < t; ++j) {                                    // t is CTA-uniform
        const int gj = grp_idx[m0 + j];
        const float cx = __fadd_rn(ox_last, centers[2 * gj]), cy = __fadd_rn(oy_last, centers[2 * gj + 1]);
        const float d = use_mm ? cdist_mm(px, py, pn, cx, cy, sq_norm2(cx, cy)) : cdist_direct(px, py, cx, cy);
        if (d < bd) { bd = d; bj = gj; }
      }
      m = m && (bj == g);
    }
    m = m && (px < iw) && (px >= 0.f) && (py < ih) && (py >= 0.f);
    if (act) { pl[s] = p_label; mk[s] = m; sx[s] = px; sy[s] = py; }
  }
  __syncthreads();
  const float pg_a = __fmul_rn(pl[K - 1], cfg.gt_alpha);
  for (int s = threadIdx.x; s < K; s += blockDim.x) {
    const float p = pl[s];
    const bool m = mk[s] && (p > cfg.merge_th) && (p > pg_a);
    pm[s] = m ? p : 0.f;
  }
  __syncthreads();
  refine_tail(pm, sx, sy, K, sx[K - 1], sy[K - 1], not_refine_in, g, cfg, red, out_pts, out_score, out_not_refine, out_chosen);
}

// ------------------------------------------------------------------------------------------------
// same-(image,label) GT groups as CSR — the device-side group_by_label (cpr_head.py:64-70 does labels.cpu()).
// One CTA per image: a stable counting sort by label (members of a group stay in ascending GT order); groups are numbered
// image-major, label-minor, so CTA b first counts the distinct labels of images 0..b-1 (a few thousand labels: cheaper than
// a second launch or a grid-wide scan).  Deterministic: no cross-CTA communication at all.
// ------------------------------------------------------------------------------------------------
constexpr int LG_THREADS = 512;
constexpr int LG_MAX_N = 8192;       // GTs per image held in shared memory
constexpr int LG_MAX_C = 1024;

__global__ void __launch_bounds__(LG_THREADS)
label_groups_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ img_ptr, int B, int G, int C,
                    int32_t* __restrict__ grp_of, int32_t* __restrict__ grp_ptr, int32_t* __restrict__ grp_idx) {
  __shared__ int s_lab[LG_MAX_N];
  __shared__ int s_cnt[LG_MAX_C], s_start[LG_MAX_C], s_rank[LG_MAX_C];
  __shared__ int s_groups, s_base;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  if (tid == 0) s_base = 0;
  // distinct labels of the images before this one
  for (int pb = 0; pb < b; ++pb) {
    const int q0 = img_ptr[pb], qn = img_ptr[pb + 1] - q0;
    for (int c = tid; c < C; c += LG_THREADS) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < qn; i += LG_THREADS) s_cnt[labels[q0 + i]] = 1;       // benign same-value race
    __syncthreads();
    int mine = 0;
    for (int c = tid; c < C; c += LG_THREADS) mine += s_cnt[c];
    mine = warp_sum_int(mine);
    if ((tid & 31) == 0 && mine) atomicAdd(&s_base, mine);                      // integer sum: order-independent
    __syncthreads();
  }
  const int group_base = s_base;
  const int g0 = img_ptr[b], n = img_ptr[b + 1] - g0;
  for (int c = tid; c < C; c += LG_THREADS) s_cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += LG_THREADS) {
    const int l = labels[g0 + i];
    s_lab[i] = l;
    atomicAdd(&s_cnt[l], 1);
  }
  __syncthreads();
  if (tid == 0) {           // C <= 1024: serial exclusive scans are negligible
    int run = 0, ng = 0;
    for (int c = 0; c < C; ++c) {
      s_start[c] = run;
      s_rank[c] = ng;
      if (s_cnt[c] > 0) {
        grp_ptr[group_base + ng] = g0 + run;
        ++ng;
      }
      run += s_cnt[c];
    }
    s_groups = ng;
  }
  __syncthreads();
  for (int i = tid; i < n; i += LG_THREADS) {
    const int l = s_lab[i];
    int before = 0;
    for (int j = 0; j < i; ++j) before += (s_lab[j] == l);      // stable rank inside the label
    grp_idx[g0 + s_start[l] + before] = g0 + i;
    grp_of[g0 + i] = group_base + s_rank[l];
  }
  if (b == B - 1)
    for (int k = group_base + s_groups + tid; k <= G; k += LG_THREADS) grp_ptr[k] = G;   // closing entry (+ unused tail)
}

}  // namespace ptb

using namespace ptb;

extern "C" int ptb_cpr_refine(const float* bag_prob, const float* bag_pts, const uint8_t* bag_valid, int G, int Kt, int K,
                              int num_classes, const int32_t* labels, const int32_t* bag_img, const int32_t* img_hw,
                              const int32_t* grp_of, const int32_t* grp_ptr, const int32_t* grp_idx,
                              const uint8_t* not_refine_in, ptb_refine_cfg cfg, float* out_pts, float* out_score,
                              uint8_t* out_not_refine, uint8_t* out_chosen, uint8_t* out_merge_valid, void* stream) {
  PTB_REQUIRE(G >= 0 && Kt > 0 && K > 0 && Kt % K == 0 && num_classes > 0, "shape");
  if (G == 0) return 0;
  PTB_REQUIRE(bag_prob && bag_pts && bag_valid && labels && bag_img && img_hw && grp_of && grp_ptr && grp_idx, "NULL input");
  PTB_REQUIRE(out_pts && out_score && out_not_refine, "NULL output");
  const size_t smem = (size_t)3 * Kt * sizeof(float);
  PTB_REQUIRE(smem <= 48 * 1024, "bag too large for shared memory");
  refine_stage_kernel<<<G, RF_THREADS, smem, (cudaStream_t)stream>>>(bag_prob, bag_pts, bag_valid, Kt, K, num_classes, labels,
                                                                   bag_img, img_hw, grp_of, grp_ptr, grp_idx, not_refine_in,
                                                                   cfg, out_pts, out_score, out_not_refine, out_chosen,
                                                                   out_merge_valid);
  return check_launch("ptb_cpr_refine");
}

extern "C" int ptb_cpr_refine_fused(const float* logit_map, int B, int H, int W, int num_classes, int ld, const float* centers,
                                    const int32_t* labels, const int32_t* bag_img, int G, const float* offsets, int K,
                                    float stride, const int32_t* pad_hw, const int32_t* img_hw, const int32_t* grp_of,
                                    const int32_t* grp_ptr, const int32_t* grp_idx, const uint8_t* not_refine_in,
                                    ptb_refine_cfg cfg, float* out_pts, float* out_score, uint8_t* out_not_refine,
                                    uint8_t* out_chosen, void* stream) {
  PTB_REQUIRE(B > 0 && H > 0 && W > 0 && G >= 0 && K > 0 && num_classes > 0, "shape");
  PTB_REQUIRE(ld % 4 == 0 && ld >= ((num_classes + 3) / 4) * 4, "ld must be a multiple of 4 covering num_classes");
  PTB_REQUIRE((uintptr_t)logit_map % 16 == 0, "logit_map must be 16-byte aligned");
  if (G == 0) return 0;
  PTB_REQUIRE(logit_map && centers && labels && bag_img && offsets && pad_hw && img_hw && grp_of && grp_ptr && grp_idx,
              "NULL input");
  PTB_REQUIRE(out_pts && out_score && out_not_refine, "NULL output");
  const size_t smem = (size_t)4 * K * sizeof(float) + (size_t)((K + 15) / 16) * 16;
  PTB_REQUIRE(smem <= 48 * 1024, "bag too large for shared memory");
  int warps = (K + 31) / 32;              // a warp takes 32 samples per pass
  if (warps < 2) warps = 2;
  if (warps > 16) warps = (warps + 1) / 2 > 16 ? 16 : (warps + 1) / 2;
  const int threads = warps * 32;
  refine_fused_kernel<<<G, threads, smem, (cudaStream_t)stream>>>(logit_map, H, W, num_classes, ld, centers, labels, bag_img,
                                                                   offsets, K, stride, pad_hw, img_hw, grp_of, grp_ptr,
                                                                   grp_idx, not_refine_in, cfg, out_pts, out_score,
                                                                   out_not_refine, out_chosen);
  return check_launch("ptb_cpr_refine_fused");
}

extern "C" int ptb_label_groups(const int32_t* labels, const int32_t* img_ptr, int B, int G, int num_classes, int max_per_image,
                                int32_t* grp_of, int32_t* grp_ptr, int32_t* grp_idx, void* stream) {
  PTB_REQUIRE(B > 0 && G >= 0 && num_classes > 0, "shape");
  PTB_REQUIRE(num_classes <= LG_MAX_C, "num_classes > 1024 not supported");
  PTB_REQUIRE(max_per_image <= LG_MAX_N, "more than 8192 GT points per image not supported");
  PTB_REQUIRE(img_ptr && grp_ptr && (G == 0 || (labels && grp_of && grp_idx)), "NULL input");
  label_groups_kernel<<<B, LG_THREADS, 0, (cudaStream_t)stream>>>(labels, img_ptr, B, G, num_classes, grp_of, grp_ptr, grp_idx);
  return check_launch("ptb_label_groups");
}
