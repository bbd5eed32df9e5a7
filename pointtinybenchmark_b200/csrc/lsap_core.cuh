// HungarianAssignerV2's matching on the GPU: scipy.optimize.linear_sum_assignment restated as a
// one-CTA-per-image kernel body.  Reference call sites: mmdet/core/bbox/assigners/hungarian_assigner.py:229-270 (cost.cpu() +
// <= topk_k scipy solves per image).  Algorithm = scipy/optimize/rectangular_lsap (Crouse 2016 shortest augmenting paths, dual
// variables u, v in fp64), including scipy's transpose rule, its reverse-filled swap-with-last `remaining` list and its tie rule —
// restated and pinned against scipy in oracle/lsap.c, whose `keyed` variant is the formulation used here:
//
//   every Dijkstra step scans the remaining columns IN PARALLEL (thread t owns columns t, t+T, ...):
//       r = ((minVal + cost[i][j]) - u[i]) - v[j];  if (r < spc[j]) { path[j] = i; spc[j] = r; }
//   and the next column is the arg-max of a TOTAL ORDER (spc asc, then st desc) with st = +(it+1) for an unassigned column at
//   position `it` of scipy's `remaining` list and -(it+1) for an assigned one — exactly what scipy's sequential scan selects
//   (the last unassigned column among the minima, else the first).  `colstate[j]` holds st (0 = column already scanned), so one
//   int32 per column carries scipy's SC flag, its position in `remaining` and the assigned bit.
//   fp64 adds/subtracts only, in scipy's order: the duals and therefore every tie are bit-identical.
//
// This header is compiled twice: by nvcc into lsap.cu (T = blockDim.x threads, shuffles + shared memory for the reduction), and
// by g++ with -DPTB_LSAP_HOST_EMU into the host-logic test (one emulated thread: checks the bookkeeping — column states, swap
// removal, dual updates, augmentation, the <= topk_k rounds and the free-list compaction — against scipy on a box without a GPU).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#ifdef PTB_LSAP_HOST_EMU
#define LSAP_FN static inline
#define LSAP_INF ((double)INFINITY)
#define LSAP_POPC(x) __builtin_popcount(x)
#else
#define LSAP_FN __device__ __forceinline__
#define LSAP_INF (__longlong_as_double(0x7ff0000000000000LL))
#define LSAP_POPC(x) __popc(x)
#endif

namespace ptb_lsap {

constexpr int LSAP_U = 4;     // columns in flight per thread

struct Cand {      // candidate column of one Dijkstra step; st == 0: none
  double val;
  int st;
  int j;
};

LSAP_FN Cand better(const Cand a, const Cand b) {
  if (b.st == 0) return a;
  if (a.st == 0) return b;
  if (a.val < b.val) return a;
  if (b.val < a.val) return b;
  return a.st > b.st ? a : b;
}

struct Bcast {     // written by thread 0 after the reduction, read by every thread after the barrier
  double minVal;
  int j;           // chosen column
  int next_i;      // row to continue from, or -1 when j is a sink
  int err;
};

// workspace of one image; M = max(N, n), m = min(N, n)
struct Ws {
  double *u, *v, *spc;
  float* T;        // [n][N] transposed cost (present when n < N)
  float* T2;       // [n][N] the same matrix gathered through the free list for rounds > 0 (cluster kernel, lsap_cluster.cuh)
  int32_t *path, *row4col, *colstate, *remaining, *sc_list, *freelist, *col4row;
  int32_t *pathstamp, *remstamp;   // path[j] / remaining[it] hold a value of the CURRENT augmentation iff the stamp equals its row
  uint8_t* flags;                  // per column: bit 0 = assigned (row4col != -1), bit 1 = v[j] may be non-zero (column was scanned once)
};

#ifdef PTB_LSAP_HOST_EMU
#define LSAP_HD static inline
#else
#define LSAP_HD __host__ __device__ inline
#endif
LSAP_HD size_t ws_bytes(int64_t N, int64_t n) {
  const size_t M = (size_t)(N > n ? N : n), m = (size_t)(N > n ? n : N);
  size_t b = 8 * (m + 2 * M);
  b += 2 * ((((n < N) ? (size_t)n * (size_t)N * 4 : 0) + 7) & ~(size_t)7);      // T and T2
  b += 7 * ((M * 4 + 7) & ~(size_t)7) + (((size_t)N * 4 + 7) & ~(size_t)7) + ((m * 4 + 7) & ~(size_t)7) + ((M + 7) & ~(size_t)7);
  return b + 64;
}
LSAP_HD Ws ws_carve(void* base, int64_t N, int64_t n) {
  const size_t M = (size_t)(N > n ? N : n), m = (size_t)(N > n ? n : N);
  char* p = reinterpret_cast<char*>(base);
  Ws w;
  w.u = reinterpret_cast<double*>(p); p += 8 * m;
  w.v = reinterpret_cast<double*>(p); p += 8 * M;
  w.spc = reinterpret_cast<double*>(p); p += 8 * M;
  w.T = reinterpret_cast<float*>(p); p += ((((n < N) ? (size_t)n * (size_t)N * 4 : 0) + 7) & ~(size_t)7);
  w.T2 = reinterpret_cast<float*>(p); p += ((((n < N) ? (size_t)n * (size_t)N * 4 : 0) + 7) & ~(size_t)7);
  const size_t mi = (M * 4 + 7) & ~(size_t)7;
  w.path = reinterpret_cast<int32_t*>(p); p += mi;
  w.row4col = reinterpret_cast<int32_t*>(p); p += mi;
  w.colstate = reinterpret_cast<int32_t*>(p); p += mi;
  w.remaining = reinterpret_cast<int32_t*>(p); p += mi;
  w.sc_list = reinterpret_cast<int32_t*>(p); p += mi;
  w.freelist = reinterpret_cast<int32_t*>(p); p += (((size_t)N * 4 + 7) & ~(size_t)7);
  w.col4row = reinterpret_cast<int32_t*>(p); p += ((m * 4 + 7) & ~(size_t)7);
  w.pathstamp = reinterpret_cast<int32_t*>(p); p += mi;
  w.remstamp = reinterpret_cast<int32_t*>(p); p += mi;
  w.flags = reinterpret_cast<uint8_t*>(p);
  return w;
}

// ---- execution context: device = the CTA, host emulation = one thread ------------------------------------------------
#ifdef PTB_LSAP_HOST_EMU
struct Ctx {
  Bcast bc;
  int tid() const { return 0; }
  int nthreads() const { return 1; }
  int lane() const { return 0; }
  int warp_width() const { return 1; }
  void sync() {}
  unsigned ballot(bool f) { return f ? 1u : 0u; }
  Cand reduce(Cand c) { return c; }            // result valid on thread 0
  int exscan(int v, int* total) { *total = v; return 0; }
};
#else
struct Ctx {
  Bcast& bc;
  Cand* s_part;                                // [32]
  int* s_scan;                                 // [33]
  __device__ Ctx(Bcast& b, Cand* p, int* sc) : bc(b), s_part(p), s_scan(sc) {}
  __device__ int tid() const { return threadIdx.x; }
  __device__ int nthreads() const { return blockDim.x; }
  __device__ int lane() const { return threadIdx.x & 31; }
  __device__ int warp_width() const { return 32; }
  __device__ void sync() { __syncthreads(); }
  __device__ unsigned ballot(bool f) { return __ballot_sync(0xffffffffu, f); }
  __device__ static Cand warp_reduce(Cand c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Cand d;
      d.val = __shfl_xor_sync(0xffffffffu, c.val, o);
      d.st = __shfl_xor_sync(0xffffffffu, c.st, o);
      d.j = __shfl_xor_sync(0xffffffffu, c.j, o);
      c = better(c, d);
    }
    return c;
  }
  // block arg-max of the total order; contains one barrier; the result is valid in warp 0 (every lane)
  __device__ Cand reduce(Cand c) {
    c = warp_reduce(c);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) s_part[w] = c;
    __syncthreads();
    Cand r;
    r.val = 0.0; r.st = 0; r.j = -1;
    if (w == 0) {
      if ((int)(threadIdx.x & 31) < nw) r = s_part[threadIdx.x & 31];
      r = warp_reduce(r);
    }
    return r;
  }
  // block-wide exclusive prefix sum of one int per thread (two barriers); *total = the sum, valid in every thread
  __device__ int exscan(int v, int* total) {
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_scan[wi] = inc;
    __syncthreads();
    if (wi == 0) {
      int t = lane < nw ? s_scan[lane] : 0;
      int ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += x;
      }
      s_scan[lane] = ti - t;                     // exclusive warp offsets
      if (lane == 31) s_scan[32] = ti;
    }
    __syncthreads();
    *total = s_scan[32];
    return s_scan[wi] + inc - v;
  }
};
#endif

// cost element of (row i, column j) of the CURRENT round's matrix:
//   transposed (n < nfree): rows = GTs, columns = free proposals:  T[i*N + freelist[j]]
//   else                  : rows = free proposals, columns = GTs:  cost[freelist[i]*n + j]

// One linear_sum_assignment of R rows x C columns (R <= C).  Returns 0 / 1 (infeasible) / 3 (internal: broken path).  col4row[R] out.
template <class CTX>
LSAP_FN int solve(CTX& cx, const float* cost, const Ws& w, int N, int n, int R, int C, bool transposed) {
  const bool ident = !transposed || C == N;     // columns are GTs, or every proposal is still free: no free-list indirection
  const int tid = cx.tid(), T = cx.nthreads();
  for (int i = tid; i < R; i += T) { w.u[i] = 0.0; w.col4row[i] = -1; }
  for (int j = tid; j < C; j += T) { w.v[j] = 0.0; w.row4col[j] = -1; w.pathstamp[j] = -1; w.remstamp[j] = -1; w.flags[j] = 0; }
  cx.sync();
  for (int cur = 0; cur < R; ++cur) {
    int i = cur, nrem = C, nsc = 0, sink = -1;
    double minVal = 0.0;
    bool first = true;
    while (sink < 0) {
      const double ui = w.u[i];
      const float* crow = transposed ? w.T + (size_t)i * (size_t)N : cost + (size_t)w.freelist[i] * (size_t)n;
      Cand best;
      best.val = 0.0; best.st = 0; best.j = -1;
      // columns are visited LSAP_U at a time per thread with every load issued before the first use: the loop is bound by
      // L2 round trips (cost row, v, and the free-list indirection in rounds > 1), so memory-level parallelism is what counts
      if (first) {
        // nothing but the two shared-memory arrays is reset per augmentation: `path` and `remaining` carry row stamps instead
        // (default path[j] = cur, default remaining[it] = C-1-it), the assigned bit and "v[j] != 0 possible" come from flags[]
        for (int j0 = tid; j0 < C; j0 += LSAP_U * T) {
          float cf[LSAP_U];
          double vj[LSAP_U];
          int fl[LSAP_U];
#pragma unroll
          for (int q = 0; q < LSAP_U; ++q) {
            const int j = j0 + q * T;
            const int jj = j < C ? j : j0;
            fl[q] = w.flags[jj];
            cf[q] = crow[ident ? jj : w.freelist[jj]];
          }
#pragma unroll
          for (int q = 0; q < LSAP_U; ++q) {
            const int j = j0 + q * T;
            const int jj = j < C ? j : j0;
            vj[q] = (fl[q] & 2) ? w.v[jj] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < LSAP_U; ++q) {
            const int j = j0 + q * T;
            if (j >= C) continue;
            const double r = ((minVal + (double)cf[q]) - ui) - vj[q];
            const int st = (fl[q] & 1) ? -(C - j) : (C - j);             // it = C-1-j  ->  it+1 = C-j
            w.colstate[j] = st;
            const double s = (r < LSAP_INF) ? r : LSAP_INF;
            w.spc[j] = s;
            if (s < LSAP_INF) {
              Cand c2;
              c2.val = s; c2.st = st; c2.j = j;
              best = better(best, c2);
            }
          }
        }
        first = false;
      } else {
        for (int j0 = tid; j0 < C; j0 += LSAP_U * T) {
          float cf[LSAP_U];
          double vj[LSAP_U], sp[LSAP_U];
          int stv[LSAP_U], fl[LSAP_U];
#pragma unroll
          for (int q = 0; q < LSAP_U; ++q) {
            const int j = j0 + q * T;
            const int jj = j < C ? j : j0;
            stv[q] = j < C ? w.colstate[jj] : 0;
            fl[q] = w.flags[jj];
            cf[q] = crow[ident ? jj : w.freelist[jj]];
            sp[q] = w.spc[jj];
          }
#pragma unroll
          for (int q = 0; q < LSAP_U; ++q) {
            const int j = j0 + q * T;
            const int jj = j < C ? j : j0;
            vj[q] = (fl[q] & 2) ? w.v[jj] : 0.0;
          }
#pragma unroll
          for (int q = 0; q < LSAP_U; ++q) {
            const int j = j0 + q * T;
            const int st = stv[q];
            if (st == 0) continue;
            double s = sp[q];
            const double r = ((minVal + (double)cf[q]) - ui) - vj[q];
            if (r < s) { w.path[j] = i; w.pathstamp[j] = cur; w.spc[j] = r; s = r; }
            if (s < LSAP_INF) {
              Cand c2;
              c2.val = s; c2.st = st; c2.j = j;
              best = better(best, c2);
            }
          }
        }
      }
      best = cx.reduce(best);
      if (tid == 0) {
        if (best.st == 0) {
          cx.bc.err = 1;
          cx.bc.j = -1; cx.bc.next_i = -1; cx.bc.minVal = LSAP_INF;
        } else {
          const int j = best.j;
          const int idx = (best.st > 0 ? best.st : -best.st) - 1;       // position of j in `remaining`
          w.colstate[j] = 0;
          w.sc_list[nsc] = j;
          const int last = nrem - 1;                                       // swap-with-last removal
          const int jm = (w.remstamp[last] == cur) ? w.remaining[last] : (C - 1 - last);
          if (jm != j) {
            w.remaining[idx] = jm;
            w.remstamp[idx] = cur;
            const int sm = w.colstate[jm];
            w.colstate[jm] = sm > 0 ? (idx + 1) : -(idx + 1);
          }
          cx.bc.err = 0;
          cx.bc.minVal = best.val;
          cx.bc.j = j;
          cx.bc.next_i = best.st > 0 ? -1 : w.row4col[j];
        }
      }
      cx.sync();
      if (cx.bc.err) return cx.bc.err;
      minVal = cx.bc.minVal;
      ++nsc; --nrem;
      if (cx.bc.next_i < 0) sink = cx.bc.j; else i = cx.bc.next_i;
      // (the next loop iteration's first shared read of bc happens after the next barrier inside reduce())
    }
    // dual variables: u[cur] += minVal; every scanned column j: d = minVal - spc[j], v[j] -= d and, when j is assigned,
    // u[row4col[j]] += d  (scipy: u[i] += minVal - spc[col4row[i]] over the visited rows; col4row[row4col[j]] == j)
    if (tid == 0) w.u[cur] += minVal;
    for (int k = tid; k < nsc; k += T) {
      const int j = w.sc_list[k];
      const double d = minVal - w.spc[j];
      w.v[j] -= d;
      w.flags[j] |= 2;                 // distinct j per k: no two threads touch the same byte
      if (k < nsc - 1) w.u[w.row4col[j]] += d;
    }
    cx.sync();
    if (tid == 0) {                    // augment along the path (<= cur+1 hops; the guard only bounds a corrupted path)
      int j = sink, hops = 0;
      w.flags[sink] |= 1;               // (the dual update above is complete: barrier)
      for (;;) {
        const int r = (w.pathstamp[j] == cur) ? w.path[j] : cur;
        w.row4col[j] = r;
        const int t = w.col4row[r];
        w.col4row[r] = j;
        j = t;
        if (r == cur) break;
        if (++hops > R || j < 0) { cx.bc.err = 3; break; }
      }
    }
    cx.sync();
    if (cx.bc.err) return cx.bc.err;
  }
  return 0;
}

// hungarian_assigner.py:229-270 for one image: cost [N][n] fp32 -> out[(row_idx ? row_idx[p] : p)] = g+1 for matched proposals
// (out pre-zeroed by the caller).  status: 0 ok, 1 infeasible.  Must be entered by the whole CTA.
template <class CTX>
LSAP_FN int hungarian_v2_image(CTX& cx, const float* cost, int N, int n, int topk_k, const Ws& w, const int32_t* row_idx,
                               int64_t* out) {
  const int tid = cx.tid(), T = cx.nthreads();
  if (N <= 0 || n <= 0) return 0;
  for (int p = tid; p < N; p += T) w.freelist[p] = p;
  cx.sync();
  int nfree = N;
  for (int round = 0; round < topk_k; ++round) {
    if (topk_k > 1 && nfree < n) break;                 // `cost_new.shape[0] // num_gts != 0`
    const bool transposed = n < nfree;                  // scipy: transpose iff more rows than columns
    const int R = transposed ? n : nfree, C = transposed ? nfree : n;
    { const int rc = solve(cx, cost, w, N, n, R, C, transposed); if (rc) return rc; }
    for (int i = tid; i < R; i += T) {
      const int p = transposed ? w.freelist[w.col4row[i]] : w.freelist[i];
      const int g = transposed ? i : w.col4row[i];
      out[row_idx ? row_idx[p] : p] = (int64_t)g + 1;
    }
    cx.sync();
    if (!transposed) { nfree = 0; continue; }           // every free proposal got a GT
    for (int i = tid; i < R; i += T) w.freelist[w.col4row[i]] = -1;
    cx.sync();
    // ordered compaction of the free list by the whole CTA: a thread counts the survivors of its contiguous segment, a block prefix sum
    // gives its output offset, the survivors go to `sc_list` (free between solves) and back.  (The first version walked the list with
    // ONE warp, 32 entries per trip with a store between two loads: 525 dependent L2 round trips = 0.4 ms per round at 16 800 proposals.)
    {
      const int seg = (nfree + T - 1) / T;
      const int e0 = tid * seg < nfree ? tid * seg : nfree;
      const int e1 = e0 + seg < nfree ? e0 + seg : nfree;
      int cnt = 0;
      for (int e = e0; e < e1; ++e) cnt += (w.freelist[e] >= 0) ? 1 : 0;
      int total = 0;
      int o = cx.exscan(cnt, &total);
      for (int e = e0; e < e1; ++e) {
        const int val = w.freelist[e];
        if (val >= 0) w.sc_list[o++] = val;
      }
      cx.sync();
      for (int e = tid; e < total; e += T) w.freelist[e] = w.sc_list[e];
      cx.sync();
      nfree = total;
    }
  }
  return 0;
}

}  // namespace ptb_lsap
