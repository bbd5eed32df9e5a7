// P2P head point path: decode + per-image top-k (p2p_head.py:125-170, 362-376), Hungarian cost matrix
// (match_cost.py:94-99, 197-214), PointAssigner (point_assigner.py:23-133) and the elementwise losses
// (focal_loss.py:11-56, smooth_l1_loss.py:25-31).  All HBM-bound scan / select work: coalesced channels-last
// reads, warp-shuffle reductions, radix select + bitonic sort in shared memory (no library sort).
#include "ptb_common.cuh"
#include "topk_select.cuh"
#include <math_constants.h>

namespace ptb {

// ------------------------------------------------------------------------------------------------
// deterministic sum helper: fixed grid, per-block partial, last block adds them in index order
// ------------------------------------------------------------------------------------------------
constexpr int SUM_BLOCKS = SCRATCH_BLOCKS;      // partials + counter: per-stream scratch block (ptb_common.cuh), not file-scope globals

__device__ __forceinline__ void block_partial_finish(float acc, SumScratch& sc, float* out) {
  __shared__ float red[32];
  __shared__ bool last;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    sc.partials[blockIdx.x] = t;
    __threadfence();
    last = (atomicAdd(&sc.done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.f;
    for (int b = 0; b < (int)gridDim.x; ++b) t += reinterpret_cast<volatile float*>(sc.partials)[b];
    out[0] += t;
    sc.done = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// decode + top-k
// ------------------------------------------------------------------------------------------------
// key[b][q] = max_c sigmoid(cls[b][cell][a*C + c]),  q = cell*k + a.   One warp per proposal.
__global__ void __launch_bounds__(256)
p2p_score_kernel(const float* __restrict__ cls_map, long long BQ, int k, int C, float* __restrict__ key) {
  const long long wq = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wq >= BQ) return;
  const float* row = cls_map + wq * C;     // [B][H][W][k*C] flattened: (b,cell,a) -> contiguous C floats
  float mx = -CUDART_INF_F;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, sigmoidf_acc(row[c]));
  mx = warp_max(mx);
  if (lane == 0) key[wq] = mx;
  (void)k;
}

// (radix select + bitonic sort: topk_select.cuh, shared with the RPN proposal path of rpn.cu)

// gather: decode the selected proposals
__global__ void __launch_bounds__(256)
p2p_gather_kernel(const float* __restrict__ cls_map, const float* __restrict__ reg_map, int H, int W, int C, int k,
                  const float* __restrict__ point_anchor, float stride, float gamma, const int32_t* __restrict__ img_hw,
                  const float* __restrict__ scale_xy, int P, int identity, const int32_t* __restrict__ idx,
                  int32_t* __restrict__ out_idx, float* __restrict__ out_pts, float* __restrict__ out_scores, long long BP) {
  const long long wr = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wr >= BP) return;
  const int b = (int)(wr / P);
  const int r = (int)(wr - (long long)b * P);
  const int q = identity ? r : idx[wr];
  const int Q = H * W * k;
  const int cell = q / k, a = q - cell * k;
  const int i = cell / W, j = cell - i * W;
  if (lane == 0) {
    if (identity) out_idx[wr] = q;
    // p2p_head.py:155-165 with PointGenerator.grid_points (no half-stride): anchor = (j*s, i*s) + point_anchor*s
    const float ax = __fadd_rn(__fmul_rn((float)j, stride), __fmul_rn(point_anchor[2 * a], stride));
    const float ay = __fadd_rn(__fmul_rn((float)i, stride), __fmul_rn(point_anchor[2 * a + 1], stride));
    const float* rg = reg_map + ((size_t)b * H * W + cell) * (2 * k) + 2 * a;
    float x = __fadd_rn(ax, __fmul_rn(__fmul_rn(rg[0], gamma), stride));
    float y = __fadd_rn(ay, __fmul_rn(__fmul_rn(rg[1], gamma), stride));
    x = fminf(fmaxf(x, 0.f), (float)img_hw[2 * b + 1]);      // p2p_head.py:374-375
    y = fminf(fmaxf(y, 0.f), (float)img_hw[2 * b]);
    if (scale_xy) { x = __fdiv_rn(x, scale_xy[2 * b]); y = __fdiv_rn(y, scale_xy[2 * b + 1]); }
    out_pts[wr * 2] = x; out_pts[wr * 2 + 1] = y;
  }
  const float* row = cls_map + ((size_t)b * Q + q) * C;
  float* orow = out_scores + wr * C;
  for (int c = lane; c < C; c += 32) orow[c] = sigmoidf_acc(row[c]);
}

// ------------------------------------------------------------------------------------------------
// cost matrix
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cost_matrix_kernel(const float* __restrict__ cls, const float* __restrict__ pts, int ldp, const int32_t* __restrict__ row_idx,
                   long long n_rows, int C, const float* __restrict__ gts, const int32_t* __restrict__ gt_labels, int n_gt,
                   float w_cls, float alpha, float gamma, float eps, float w_dis, float fx, float fy, float* __restrict__ cost) {
  const long long total = n_rows * n_gt;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const long long r = e / n_gt;
    const int g = (int)(e - r * n_gt);
    const long long q = row_idx ? row_idx[r] : r;
    const float p = sigmoidf_acc(cls[q * C + gt_labels[g]]);
    const float pg = (gamma == 2.f) ? __fmul_rn(p, p) : powf(p, gamma);
    const float omp = __fsub_rn(1.f, p);
    const float og = (gamma == 2.f) ? __fmul_rn(omp, omp) : powf(omp, gamma);
    // match_cost.py:95-99
    const float neg = __fmul_rn(__fmul_rn(-logf(__fadd_rn(omp, eps)), __fsub_rn(1.f, alpha)), pg);
    const float pos = __fmul_rn(__fmul_rn(-logf(__fadd_rn(p, eps)), alpha), og);
    const float cc = __fmul_rn(__fsub_rn(pos, neg), w_cls);
    const float dx = fabsf(__fsub_rn(__fdiv_rn(pts[q * ldp], fx), __fdiv_rn(gts[2 * g], fx)));
    const float dy = fabsf(__fsub_rn(__fdiv_rn(pts[q * ldp + 1], fy), __fdiv_rn(gts[2 * g + 1], fy)));
    const float dc = __fmul_rn(__fadd_rn(dx, dy), w_dis);    // cdist p=1, match_cost.py:213-214
    cost[e] = __fadd_rn(cc, dc);
  }
}

// ------------------------------------------------------------------------------------------------
// PointAssigner
// ------------------------------------------------------------------------------------------------
struct PaScratch {
  int lmin, lmax;
};

__global__ void pa_init_kernel(const float* __restrict__ points, int N, unsigned long long* __restrict__ best, PaScratch* sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && blockIdx.x == 0) { /* lmin/lmax set by host memset pattern below */ }
  if (i < N) {
    best[i] = 0xFFFFFFFFFFFFFFFFull;
    const int lv = (int)log2f(points[3 * i + 2]);
    atomicMin(&sc->lmin, lv);
    atomicMax(&sc->lmax, lv);
  }
}
__global__ void pa_reset_kernel(PaScratch* sc) { sc->lmin = 0x7fffffff; sc->lmax = -0x7fffffff; }

// one CTA per GT: pos_num nearest points of its level claim the GT through a packed 64-bit atomicMin
__global__ void __launch_bounds__(256)
pa_assign_kernel(const float* __restrict__ points, int N, const float* __restrict__ gts, float scale, int pos_num,
                 const PaScratch* __restrict__ sc, unsigned long long* __restrict__ best) {
  __shared__ unsigned long long red[8];
  __shared__ unsigned long long chosen_prev;
  const int j = blockIdx.x;
  const float x1 = gts[4 * j], y1 = gts[4 * j + 1], x2 = gts[4 * j + 2], y2 = gts[4 * j + 3];
  const float cx = __fdiv_rn(__fadd_rn(x1, x2), 2.f), cy = __fdiv_rn(__fadd_rn(y1, y2), 2.f);
  const float w = fmaxf(__fsub_rn(x2, x1), 1e-6f), h = fmaxf(__fsub_rn(y2, y1), 1e-6f);
  int lvl = (int)(__fdiv_rn(__fadd_rn(log2f(__fdiv_rn(w, scale)), log2f(__fdiv_rn(h, scale))), 2.f));
  lvl = min(max(lvl, sc->lmin), sc->lmax);
  unsigned long long prev = 0ull;   // selections are strictly increasing in (dist, idx): pick the next one above `prev`
  for (int round = 0; round < pos_num; ++round) {
    unsigned long long mine = 0xFFFFFFFFFFFFFFFFull;
    for (int i = threadIdx.x; i < N; i += 256) {
      if ((int)log2f(points[3 * i + 2]) != lvl) continue;
      const float dx = __fdiv_rn(__fsub_rn(points[3 * i], cx), w), dy = __fdiv_rn(__fsub_rn(points[3 * i + 1], cy), h);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)i;
      if ((round == 0 || key > prev) && key < mine) mine = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
      if (other < mine) mine = other;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long m = red[0];
      for (int wv = 1; wv < 8; ++wv) if (red[wv] < m) m = red[wv];
      chosen_prev = m;
      if (m != 0xFFFFFFFFFFFFFFFFull) {
        const unsigned int pi = (unsigned int)(m & 0xFFFFFFFFull);
        const unsigned long long claim = (m & 0xFFFFFFFF00000000ull) | (unsigned int)j;   // (dist, gt index)
        atomicMin(&best[pi], claim);
      }
    }
    __syncthreads();
    prev = chosen_prev;
    if (prev == 0xFFFFFFFFFFFFFFFFull) break;
  }
}

__global__ void pa_finish_kernel(const unsigned long long* __restrict__ best, int N, int64_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = best[i] == 0xFFFFFFFFFFFFFFFFull ? 0 : (int64_t)(best[i] & 0xFFFFFFFFull) + 1;
}

// ------------------------------------------------------------------------------------------------
// elementwise losses with sums
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
focal_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels, const float* __restrict__ weight, long long M,
             int C, float gamma, float alpha, float* loss_sum, const float* __restrict__ scale, float* __restrict__ grad,
             SumScratch* __restrict__ scr) {
  const long long total = M * C;
  const float sc = (grad && scale) ? scale[0] : 1.f;
  float acc = 0.f;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const long long m = e / C;
    const int c = (int)(e - m * C);
    const float w = weight ? weight[m] : 1.f;
    const float t = (labels[m] == c) ? 1.f : 0.f;
    const float v = x[e];
    const float p = sigmoidf_acc(v);
    const float pt = (1.f - p) * t + p * (1.f - t);
    const float a = alpha * t + (1.f - alpha) * (1.f - t);
    const float ptg = (gamma == 2.f) ? pt * pt : powf(pt, gamma);
    const float bce = fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
    acc += bce * (a * ptg) * w;
    if (grad) {
      const float dpt = (t > 0.5f ? -1.f : 1.f) * p * (1.f - p);
      const float ptg1 = (gamma == 2.f) ? 2.f * pt : gamma * powf(pt, gamma - 1.f);
      grad[e] = sc * w * a * (ptg1 * dpt * bce + ptg * (p - t));
    }
  }
  if (loss_sum) block_partial_finish(acc, *scr, loss_sum);
}

__global__ void __launch_bounds__(256)
smooth_l1_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ weight, long long n,
                 float inv_norm, float beta, float* loss_sum, const float* __restrict__ scale, float* __restrict__ grad,
                 SumScratch* __restrict__ scr) {
  const float sc = (grad && scale) ? scale[0] : 1.f;
  float acc = 0.f;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long long)gridDim.x * 256) {
    const float w = weight ? weight[e] : 1.f;
    const float diff = (pred[e] - target[e]) * inv_norm;
    const float d = fabsf(diff);
    acc += (d < beta ? 0.5f * d * d / beta : d - 0.5f * beta) * w;
    if (grad) grad[e] = sc * w * inv_norm * (d < beta ? diff / beta : (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)));
  }
  if (loss_sum) block_partial_finish(acc, *scr, loss_sum);
}

}  // namespace ptb

using namespace ptb;

extern "C" uint64_t ptb_p2p_decode_topk_workspace(int B, int H, int W, int k) {
  return (uint64_t)B * H * W * k * sizeof(float) + (uint64_t)B * TOPK_MAX * sizeof(int32_t);
}

extern "C" int ptb_p2p_decode_topk(const float* cls_map, const float* reg_map, int B, int H, int W, int num_classes, int k,
                                   const float* point_anchor, float stride, float pts_gamma, const int32_t* img_hw,
                                   const float* scale_xy, int nms_pre, int32_t* out_topk_idx, float* out_pts, float* out_scores,
                                   void* workspace, uint64_t workspace_bytes, void* stream) {
  PTB_REQUIRE(B > 0 && H > 0 && W > 0 && num_classes > 0 && k > 0, "shape");
  PTB_REQUIRE(cls_map && reg_map && point_anchor && img_hw && out_topk_idx && out_pts && out_scores, "NULL input");
  const int Q = H * W * k;
  const bool identity = !(nms_pre > 0 && nms_pre < Q);
  const int P = identity ? Q : nms_pre;
  PTB_REQUIRE(identity || nms_pre <= TOPK_MAX, "nms_pre > 4096 not supported");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (!identity) {
    PTB_REQUIRE(workspace && workspace_bytes >= ptb_p2p_decode_topk_workspace(B, H, W, k), "workspace too small");
    float* key = reinterpret_cast<float*>(workspace);
    const long long BQ = (long long)B * Q;
    p2p_score_kernel<<<(unsigned)((BQ * 32 + 255) / 256), 256, 0, st>>>(cls_map, BQ, k, num_classes, key);
    if ((rc = check_launch("ptb_p2p_decode_topk/score"))) return rc;
    p2p_select_kernel<<<B, SEL_THREADS, 0, st>>>(key, Q, P, out_topk_idx, P);
    if ((rc = check_launch("ptb_p2p_decode_topk/select"))) return rc;
  }
  const long long BP = (long long)B * P;
  p2p_gather_kernel<<<(unsigned)((BP * 32 + 255) / 256), 256, 0, st>>>(cls_map, reg_map, H, W, num_classes, k, point_anchor,
                                                                     stride, pts_gamma, img_hw, scale_xy, P, identity ? 1 : 0,
                                                                     out_topk_idx, out_topk_idx, out_pts, out_scores, BP);
  return check_launch("ptb_p2p_decode_topk/gather");
}

extern "C" int ptb_p2p_cost_matrix(const float* cls_logits, const float* pts, int ldp, const int32_t* row_idx, int n_rows,
                                   int num_classes, const float* gts, const int32_t* gt_labels, int n_gt, float w_cls,
                                   float alpha, float gamma, float eps, float w_dis, float fx, float fy, float* cost,
                                   void* stream) {
  PTB_REQUIRE(n_rows >= 0 && n_gt >= 0 && num_classes > 0 && ldp >= 2, "shape");
  if (n_rows == 0 || n_gt == 0) return 0;
  PTB_REQUIRE(cls_logits && pts && gts && gt_labels && cost, "NULL input");
  long long blocks = ((long long)n_rows * n_gt + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cost_matrix_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(cls_logits, pts, ldp, row_idx, n_rows, num_classes, gts,
                                                                       gt_labels, n_gt, w_cls, alpha, gamma, eps, w_dis, fx, fy,
                                                                       cost);
  return check_launch("ptb_p2p_cost_matrix");
}

extern "C" uint64_t ptb_point_assigner_workspace(int N, int n) {
  (void)n;
  return (uint64_t)N * sizeof(unsigned long long) + 64;
}

extern "C" int ptb_point_assigner(const float* points, int N, const float* gt_bboxes, int n, float scale, int pos_num,
                                  int64_t* out_gt_inds, void* workspace, uint64_t workspace_bytes, void* stream) {
  PTB_REQUIRE(N >= 0 && n >= 0 && pos_num > 0 && scale > 0.f, "shape");
  if (N == 0) return 0;
  PTB_REQUIRE(points && out_gt_inds, "NULL input");
  PTB_REQUIRE(workspace && workspace_bytes >= ptb_point_assigner_workspace(N, n), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  PaScratch* sc = reinterpret_cast<PaScratch*>(workspace);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + 64);
  int rc;
  pa_reset_kernel<<<1, 1, 0, st>>>(sc);
  if ((rc = check_launch("ptb_point_assigner/reset"))) return rc;
  pa_init_kernel<<<(N + 255) / 256, 256, 0, st>>>(points, N, best, sc);
  if ((rc = check_launch("ptb_point_assigner/init"))) return rc;
  if (n > 0) {
    PTB_REQUIRE(gt_bboxes, "NULL gt_bboxes");
    pa_assign_kernel<<<n, 256, 0, st>>>(points, N, gt_bboxes, scale, pos_num, sc, best);
    if ((rc = check_launch("ptb_point_assigner/assign"))) return rc;
  }
  pa_finish_kernel<<<(N + 255) / 256, 256, 0, st>>>(best, N, out_gt_inds);
  return check_launch("ptb_point_assigner/finish");
}

extern "C" int ptb_sigmoid_focal_fwd_bwd(const float* logits, const int64_t* labels, const float* weight, int64_t M,
                                         int num_classes, float gamma, float alpha, float* loss_sum, const float* scale,
                                         float* grad, void* stream) {
  PTB_REQUIRE(M >= 0 && num_classes > 0, "shape");
  if (M == 0) return 0;
  PTB_REQUIRE(logits && labels && (loss_sum || grad), "NULL input");
  StreamScratch* scr = stream_scratch(stream);
  if (!scr) return 1;
  focal_kernel<<<SUM_BLOCKS, 256, 0, (cudaStream_t)stream>>>(logits, labels, weight, M, num_classes, gamma, alpha, loss_sum, scale,
                                                           grad, &scr->focal);
  return check_launch("ptb_sigmoid_focal_fwd_bwd");
}

extern "C" int ptb_smooth_l1_fwd_bwd(const float* pred, const float* target, const float* weight, int64_t M, float inv_norm,
                                     float beta, float* loss_sum, const float* scale, float* grad, void* stream) {
  PTB_REQUIRE(M >= 0 && beta > 0.f, "shape");
  if (M == 0) return 0;
  PTB_REQUIRE(pred && target && (loss_sum || grad), "NULL input");
  StreamScratch* scr = stream_scratch(stream);
  if (!scr) return 1;
  smooth_l1_kernel<<<SUM_BLOCKS, 256, 0, (cudaStream_t)stream>>>(pred, target, weight, M * 2, inv_norm, beta, loss_sum, scale, grad,
                                                               &scr->sl1);
  return check_launch("ptb_smooth_l1_fwd_bwd");
}
