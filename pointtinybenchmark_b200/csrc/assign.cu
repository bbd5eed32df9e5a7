// MaxIoUAssigner for dense anchors (reference mmdet/core/bbox/assigners/max_iou_assigner.py:60-212 over
// bbox_overlaps, mmdet/core/bbox/iou_calculators/iou2d_calculator.py:211-256).  The reference materialises the (k, n) IoU matrix
// (k GTs x n anchors: 81 840 x k fp32 at BASELINE.json configs[3]) and walks it several times from Python; here the matrix is
// never written:
//   K1  miou_scan_kernel   thread per anchor: ignore test (IoF against the ignore boxes), running max / first arg-max over the
//                          GTs, and per GT the (max IoU, lowest anchor index) as one packed 64-bit key: warp shuffle max ->
//                          shared-memory atomicMax -> one global atomicMax per (CTA, GT).  Integer atomics only: deterministic.
//   K2  miou_assign_kernel thread per anchor: negative / positive rule, then the low-quality matching loop over the GTs in the
//                          reference's order (later GTs overwrite earlier ones), recomputing the anchor's IoU with each GT
//                          bit-identically; labels.
// IoU arithmetic = the reference's torch ops in the same order (no FMA): area = (x2-x1)*(y2-y1); wh = clamp(min(rb)-max(lt), 0);
// overlap = w*h; union = max((area_g + area_b) - overlap, 1e-6); iou = overlap / union.  Integer outputs are bit-exact targets.
#include "ptb_common.cuh"
#include <math_constants.h>

namespace ptb {

constexpr int MI_THREADS = 256;
constexpr int MI_TILE = 256;       // GTs per shared-memory tile

struct MiouCfg {
  float pos_iou_thr, neg_lo, neg_hi, min_pos_iou, ignore_iof_thr;
  int gt_max_assign_all, match_low_quality, ignore_mode;   // ignore_mode: 0 off, 1 iof w.r.t. the anchor, 2 w.r.t. the ignore box
};

__device__ __forceinline__ float box_area(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

__device__ __forceinline__ float box_overlap(const float4 a, const float4 b) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  return __fmul_rn(w, h);
}
// bbox_overlaps(mode='iou') element (gt g, anchor b)
__device__ __forceinline__ float box_iou(const float4 g, float area_g, const float4 b, float area_b) {
  const float ov = box_overlap(g, b);
  const float uni = fmaxf(__fsub_rn(__fadd_rn(area_g, area_b), ov), 1e-6f);
  return __fdiv_rn(ov, uni);
}
// monotone map float -> uint32 (handles the -1 of ignored anchors)
__device__ __forceinline__ unsigned int f2ord(float v) {
  const unsigned int b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

__global__ void __launch_bounds__(MI_THREADS)
miou_scan_kernel(const float4* __restrict__ bboxes, int N, const float4* __restrict__ gts, int n, const float4* __restrict__ ign, int m,
                 MiouCfg cfg, float* __restrict__ max_ov, int32_t* __restrict__ argmax, uint8_t* __restrict__ ignored,
                 unsigned long long* __restrict__ gt_key /*[n], zeroed*/) {
  __shared__ float4 s_gt[MI_TILE];
  __shared__ unsigned long long s_key[MI_TILE];
  const int b = blockIdx.x * MI_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool have = b < N;
  const float4 bb = have ? bboxes[b] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float area_b = box_area(bb);
  bool ig = false;
  if (have && cfg.ignore_mode) {           // ignore_overlaps.max over the ignore boxes > thr  (max_iou_assigner.py:107-117)
    float mx = -CUDART_INF_F;
    for (int j = 0; j < m; ++j) {
      const float4 q = ign[j];
      const float area = cfg.ignore_mode == 1 ? area_b : box_area(q);     // iof: union = area of bboxes1
      mx = fmaxf(mx, __fdiv_rn(box_overlap(bb, q), fmaxf(area, 1e-6f)));
    }
    ig = mx > cfg.ignore_iof_thr;
  }
  float best = -CUDART_INF_F;
  int besti = 0;
  for (int t0 = 0; t0 < n; t0 += MI_TILE) {
    const int tn = min(MI_TILE, n - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < tn; j += MI_THREADS) { s_gt[j] = gts[t0 + j]; s_key[j] = 0ull; }
    __syncthreads();
    for (int j = 0; j < tn; ++j) {
      const float4 g = s_gt[j];
      const float v = ig ? -1.f : box_iou(g, box_area(g), bb, area_b);
      if (have && v > best) { best = v; besti = t0 + j; }                 // first maximum (torch.max(dim=0))
      unsigned long long key = have ? (((unsigned long long)f2ord(v) << 32) | (unsigned int)(0xFFFFFFFFu - (unsigned int)b)) : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
      }
      if (lane == 0 && key) atomicMax(&s_key[j], key);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < tn; j += MI_THREADS)
      if (s_key[j]) atomicMax(&gt_key[t0 + j], s_key[j]);
  }
  if (have) {
    max_ov[b] = best;
    argmax[b] = besti;
    ignored[b] = ig;
  }
}

__global__ void __launch_bounds__(MI_THREADS)
miou_assign_kernel(const float4* __restrict__ bboxes, int N, const float4* __restrict__ gts, int n, const int32_t* __restrict__ gt_labels,
                   MiouCfg cfg, const float* __restrict__ max_ov, const int32_t* __restrict__ argmax,
                   const uint8_t* __restrict__ ignored, const unsigned long long* __restrict__ gt_key,
                   long long* __restrict__ out_gt_inds, long long* __restrict__ out_labels) {
  __shared__ float4 s_gt[MI_TILE];
  __shared__ float s_gmax[MI_TILE];
  __shared__ int s_garg[MI_TILE];
  const int b = blockIdx.x * MI_THREADS + threadIdx.x;
  const bool have = b < N;
  const float4 bb = have ? bboxes[b] : make_float4(0.f, 0.f, 0.f, 0.f);
  const float area_b = box_area(bb);
  const bool ig = have ? ignored[b] != 0 : false;
  long long a = -1;                                                        // 1. assign -1 by default
  if (have) {
    const float mo = max_ov[b];
    if (mo >= cfg.neg_lo && mo < cfg.neg_hi) a = 0;                        // 2. negatives
    if (mo >= cfg.pos_iou_thr) a = (long long)argmax[b] + 1;              // 3. positives
  }
  if (cfg.match_low_quality) {                                             // 4. every GT claims its best anchor(s), in GT order
    for (int t0 = 0; t0 < n; t0 += MI_TILE) {
      const int tn = min(MI_TILE, n - t0);
      __syncthreads();
      for (int j = threadIdx.x; j < tn; j += MI_THREADS) {
        s_gt[j] = gts[t0 + j];
        const unsigned long long k = gt_key[t0 + j];
        s_gmax[j] = ord2f((unsigned int)(k >> 32));
        s_garg[j] = (int)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull));
      }
      __syncthreads();
      if (have) {
        for (int j = 0; j < tn; ++j) {
          const float gm = s_gmax[j];
          if (gm >= cfg.min_pos_iou) {
            if (cfg.gt_max_assign_all) {
              const float4 g = s_gt[j];
              const float v = ig ? -1.f : box_iou(g, box_area(g), bb, area_b);
              if (v == gm) a = t0 + j + 1;
            } else if (s_garg[j] == b) {
              a = t0 + j + 1;
            }
          }
        }
      }
    }
  }
  if (have) {
    out_gt_inds[b] = a;
    if (out_labels) out_labels[b] = a > 0 ? (long long)gt_labels[a - 1] : -1;
  }
}

__global__ void miou_fill_kernel(long long* __restrict__ gt_inds, float* __restrict__ max_ov, long long* __restrict__ labels, int N,
                                 long long v) {
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b < N) {
    gt_inds[b] = v;
    max_ov[b] = 0.f;
    if (labels) labels[b] = -1;
  }
}

// the IoU / IoF matrix itself (BboxOverlaps2D, is_aligned=False): out[i][j] for boxes1[i], boxes2[j]
__global__ void __launch_bounds__(256)
bbox_overlaps_kernel(const float4* __restrict__ b1, int m, const float4* __restrict__ b2, int n, int iof, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)m * n) return;
  const int i = (int)(idx / n), j = (int)(idx - (long long)i * n);
  const float4 p = b1[i], q = b2[j];
  const float ov = box_overlap(p, q);
  const float uni = iof ? box_area(p) : __fsub_rn(__fadd_rn(box_area(p), box_area(q)), ov);
  out[idx] = __fdiv_rn(ov, fmaxf(uni, 1e-6f));
}

}  // namespace ptb

using namespace ptb;

extern "C" uint64_t ptb_max_iou_assign_workspace(int N, int n_gt) {
  if (N < 0 || n_gt < 0) return 0;
  return (uint64_t)n_gt * 8 + (uint64_t)N * 4 + (uint64_t)N + 64;
}

extern "C" int ptb_max_iou_assign(const float* bboxes, int N, const float* gt_bboxes, int n_gt, const int32_t* gt_labels,
                                  const float* gt_bboxes_ignore, int n_ignore, float pos_iou_thr, float neg_iou_lo, float neg_iou_hi,
                                  float min_pos_iou, int gt_max_assign_all, int match_low_quality, float ignore_iof_thr,
                                  int ignore_wrt_candidates, int64_t* out_gt_inds, float* out_max_overlaps, int64_t* out_labels,
                                  void* workspace, uint64_t workspace_bytes, void* stream) {
  PTB_REQUIRE(N >= 0 && n_gt >= 0 && n_ignore >= 0, "shape");
  if (N == 0) return 0;
  PTB_REQUIRE(bboxes && out_gt_inds && out_max_overlaps, "NULL input");
  PTB_REQUIRE(((uintptr_t)bboxes % 16 == 0) && ((uintptr_t)gt_bboxes % 16 == 0) && ((uintptr_t)gt_bboxes_ignore % 16 == 0),
              "boxes must be 16-byte aligned [.][4] fp32 arrays");
  PTB_REQUIRE(!out_labels || gt_labels || n_gt == 0, "gt_labels required for out_labels");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (N + MI_THREADS - 1) / MI_THREADS;
  if (n_gt == 0) {                     // no ground truth: everything is background (max_iou_assigner.py:145-161)
    miou_fill_kernel<<<(N + 255) / 256, 256, 0, st>>>(reinterpret_cast<long long*>(out_gt_inds), out_max_overlaps,
                                                      reinterpret_cast<long long*>(out_labels), N, 0);
    return check_launch("ptb_max_iou_assign/fill");
  }
  PTB_REQUIRE(gt_bboxes && workspace && workspace_bytes >= ptb_max_iou_assign_workspace(N, n_gt), "workspace too small / NULL gt");
  unsigned long long* gt_key = reinterpret_cast<unsigned long long*>(workspace);
  int32_t* argmax = reinterpret_cast<int32_t*>(gt_key + n_gt);
  uint8_t* ignored = reinterpret_cast<uint8_t*>(argmax + N);
  MiouCfg cfg;
  cfg.pos_iou_thr = pos_iou_thr; cfg.neg_lo = neg_iou_lo; cfg.neg_hi = neg_iou_hi; cfg.min_pos_iou = min_pos_iou;
  cfg.ignore_iof_thr = ignore_iof_thr; cfg.gt_max_assign_all = gt_max_assign_all; cfg.match_low_quality = match_low_quality;
  cfg.ignore_mode = (ignore_iof_thr > 0.f && gt_bboxes_ignore && n_ignore > 0) ? (ignore_wrt_candidates ? 1 : 2) : 0;
  if (cudaMemsetAsync(gt_key, 0, (size_t)n_gt * 8, st) != cudaSuccess) return fail("%s", "ptb_max_iou_assign: cudaMemsetAsync failed");
  miou_scan_kernel<<<blocks, MI_THREADS, 0, st>>>(reinterpret_cast<const float4*>(bboxes), N, reinterpret_cast<const float4*>(gt_bboxes),
                                                  n_gt, reinterpret_cast<const float4*>(gt_bboxes_ignore), n_ignore, cfg,
                                                  out_max_overlaps, argmax, ignored, gt_key);
  int rc = check_launch("ptb_max_iou_assign/scan");
  if (rc) return rc;
  miou_assign_kernel<<<blocks, MI_THREADS, 0, st>>>(reinterpret_cast<const float4*>(bboxes), N, reinterpret_cast<const float4*>(gt_bboxes),
                                                    n_gt, gt_labels, cfg, out_max_overlaps, argmax, ignored, gt_key,
                                                    reinterpret_cast<long long*>(out_gt_inds), reinterpret_cast<long long*>(out_labels));
  return check_launch("ptb_max_iou_assign/assign");
}

extern "C" int ptb_bbox_overlaps(const float* boxes1, int m, const float* boxes2, int n, int mode_iof, float* out, void* stream) {
  PTB_REQUIRE(m >= 0 && n >= 0, "shape");
  if ((long long)m * n == 0) return 0;
  PTB_REQUIRE(boxes1 && boxes2 && out, "NULL input");
  PTB_REQUIRE(((uintptr_t)boxes1 % 16 == 0) && ((uintptr_t)boxes2 % 16 == 0), "boxes must be 16-byte aligned [.][4] fp32 arrays");
  const long long total = (long long)m * n;
  bbox_overlaps_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes1), m, reinterpret_cast<const float4*>(boxes2), n, mode_iof, out);
  return check_launch("ptb_bbox_overlaps");
}
