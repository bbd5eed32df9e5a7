// HungarianAssignerV2's matching with the column scan of every Dijkstra step split over an 8-CTA thread-block cluster (round 2).
//
// Why: ncu of the one-CTA-per-image kernel (lsap_core.cuh, profiles/r02_ncu_hungarian_v2.json) shows ~980 instructions per thread and
// Dijkstra step for the 17 columns a thread owns, issue slots 51 % busy on the ONE SM an image runs on and the rest of the time spent at
// the block barrier: the step is bound by one SM's instruction issue, not by memory.  Here an image is solved by a cluster of CL_N CTAs:
//   * CTA r owns the contiguous column range [r*Cc, (r+1)*Cc) and keeps ITS part of the per-step read-write column state (spc fp64,
//     colstate int32, flags) in its own shared memory; a thread scans <= 5 columns, all global loads of a step in flight at once;
//   * per step: block arg-best -> every CTA stores its candidate into the leader's shared memory (st.shared::cluster) -> cluster
//     barrier A -> the leader's thread 0 picks the winner of the 8 (the same total order: identical to the sequential scan) and runs the
//     bookkeeping of lsap_core.cuh's solve() (swap-with-last removal, next row), reaching into the owners' shared memory for the two
//     column states it touches, then stores the broadcast record (minVal, column, next row) into every CTA -> cluster barrier B;
//   * dual update, augmentation, free-list compaction and the output scatter are the leader's (a handful of entries each); the
//     per-round gather of the cost matrix is shared by all CTAs.
//   * v2 of this file: the first cluster version kept v, path, row4col, remaining, u, col4row in the L2-resident workspace like the
//     one-CTA kernel and was NOT faster (10.3 vs 9.6 ms per 16 images): an augmentation is a chain of ~10 dependent L2 round trips in
//     the serial sections (remstamp -> remaining, row4col, the dual update's read-modify-writes, the path walk), not instruction
//     issue.  Now every per-column array lives in its owner's shared memory (33 B per column; the leader reaches it with
//     ld/st.shared::cluster, ~215 cycles instead of an L2 trip), `remaining` + stamps (uint16) and the row arrays u / col4row are in
//     the leader's shared memory; the only global traffic of a Dijkstra step is the cost row.
// Arrays that cross CTAs through GLOBAL memory (free list, gathered matrix) are read with ld.global.cg: an SM's L1 is not coherent
// with another SM's stores; the cluster barriers (arrive.release / wait.acquire) order them.
// The algorithm, its fp64 operation order and its tie rule are those of lsap_core.cuh (read its header first): assignments are
// bit-identical to scipy and to the single-CTA kernel (tests/test_lsap.py runs both).
#pragma once
#include "lsap_core.cuh"

namespace ptb_lsap {

constexpr int CL_N = 8;            // CTAs per image (the portable cluster size limit)
constexpr int CL_T = 512;          // threads per CTA
constexpr int CL_COLS = 2200;      // columns per CTA: 8 x 2200 = 17 600 = the single-CTA kernel's shared-memory limit
constexpr int CL_U = (CL_COLS + CL_T - 1) / CL_T;     // 5 columns per thread, one group

constexpr int CL_ROWS = 2048;      // rows (min(N, n)) the leader keeps u / col4row for in shared memory; larger problems use the one-CTA kernel
constexpr int CL_MAXC = CL_N * CL_COLS;
constexpr int CL_SC = 256;         // scanned-column list entries cached in shared memory (the rest goes through the workspace)

struct ClBcast {                   // written by the leader into every CTA
  double minVal;
  int j, next_i;                   // chosen column; row to continue from or -1 when j is a sink
  int err, pad;
  double ui;                       // u[next_i]
};

struct ClShared {                  // dynamic shared memory of every CTA (168 KB)
  // per-column state of the columns this CTA owns
  double spc[CL_COLS];
  double v[CL_COLS];
  int32_t colstate[CL_COLS];
  int32_t row4col[CL_COLS];
  int32_t path[CL_COLS];
  int32_t pathstamp[CL_COLS];
  uint8_t flags[CL_COLS + 8];
  // used in the leader only
  double u[CL_ROWS];
  int32_t col4row[CL_ROWS];
  uint16_t remaining[CL_MAXC];
  uint16_t remstamp[CL_MAXC];     // 0xFFFF = no entry of the current augmentation
  int32_t sc[CL_SC];
  Cand slot[CL_N];                 // the CTAs' candidates of the current step
  // every CTA
  Cand part[CL_T / 32];
  int scan[33];
  ClBcast bc;
};

__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a pointer into this CTA's shared memory) in CTA `rank`
__device__ __forceinline__ uint32_t cl_map(const void* p, uint32_t rank) {
  uint32_t r;
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_st_u64(uint32_t a, unsigned long long v) { asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void cl_st_u32(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void cl_st_u8(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t cl_ld_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared::cluster.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cl_ld_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long cl_ld_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared::cluster.b64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}

__device__ __forceinline__ Cand cl_warp_reduce(Cand c) {
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

// column j of the current round lives in CTA j / Cc at local index j % Cc
struct ClCols {
  ClShared* S;
  int Cc;
  __device__ __forceinline__ uint32_t a_colstate(int j) const { return cl_map(&S->colstate[j % Cc], (uint32_t)(j / Cc)); }
  __device__ __forceinline__ uint32_t a_spc(int j) const { return cl_map(&S->spc[j % Cc], (uint32_t)(j / Cc)); }
  __device__ __forceinline__ uint32_t a_v(int j) const { return cl_map(&S->v[j % Cc], (uint32_t)(j / Cc)); }
  __device__ __forceinline__ uint32_t a_flags(int j) const { return cl_map(&S->flags[j % Cc], (uint32_t)(j / Cc)); }
  __device__ __forceinline__ uint32_t a_row4col(int j) const { return cl_map(&S->row4col[j % Cc], (uint32_t)(j / Cc)); }
  __device__ __forceinline__ uint32_t a_path(int j) const { return cl_map(&S->path[j % Cc], (uint32_t)(j / Cc)); }
  __device__ __forceinline__ uint32_t a_pathstamp(int j) const { return cl_map(&S->pathstamp[j % Cc], (uint32_t)(j / Cc)); }
};
__device__ __forceinline__ void cl_bcast(ClShared& S, const ClBcast& b) {
#pragma unroll
  for (int r = 0; r < CL_N; ++r) {
    const uint32_t a = cl_map(&S.bc, (uint32_t)r);
    cl_st_u64(a, (unsigned long long)__double_as_longlong(b.minVal));
    cl_st_u64(a + 8u, ((unsigned long long)(uint32_t)b.next_i << 32) | (uint32_t)b.j);
    cl_st_u32(a + 16u, (uint32_t)b.err);
    cl_st_u64(a + 24u, (unsigned long long)__double_as_longlong(b.ui));
  }
}

// One linear_sum_assignment of R rows x C columns (R <= C, R <= CL_ROWS, C <= CL_MAXC), entered by every thread of every CTA of the
// cluster.  Returns 0 / 1 (infeasible) / 3 (internal), the same value in every CTA.  S.col4row[R] (leader) out.
__device__ __forceinline__ int solve_cl(ClShared& S, const uint32_t rank, const float* cost, const float* Tm, const Ws& w, int N, int n, int R,
                                        int C, bool transposed) {
  const int tid = threadIdx.x;
  constexpr int T = CL_T;
  const bool leader = rank == 0;
  ClCols cols;
  cols.S = &S;
  cols.Cc = (C + CL_N - 1) / CL_N;                       // <= CL_COLS (checked by the host)
  const int c0 = min(C, (int)rank * cols.Cc), c1 = min(C, c0 + cols.Cc);
  if (leader) {
    for (int i = tid; i < R; i += T) { S.u[i] = 0.0; S.col4row[i] = -1; }
    for (int j = tid; j < C; j += T) S.remstamp[j] = 0xFFFFu;
    if (tid == 0) S.bc.err = 0;
  }
  for (int j = c0 + tid; j < c1; j += T) {
    const int jl = j - c0;
    S.v[jl] = 0.0; S.row4col[jl] = -1; S.pathstamp[jl] = -1; S.flags[jl] = 0;
  }
  __syncthreads();
  cl_sync();
  for (int cur = 0; cur < R; ++cur) {
    int i = cur, nrem = C, nsc = 0, sink = -1;
    double minVal = 0.0, ui = 0.0;                       // u[cur] is still 0: a row's dual only changes once the row is assigned
    bool first = true;
    while (sink < 0) {
      const float* crow = transposed ? Tm + (size_t)i * (size_t)N : cost + (size_t)__ldcg(&w.freelist[i]) * (size_t)n;
      Cand best;
      best.val = 0.0; best.st = 0; best.j = -1;
      if (c0 < c1) {
        float cf[CL_U];
#pragma unroll
        for (int q = 0; q < CL_U; ++q) {
          const int j = c0 + tid + q * T;
          cf[q] = __ldcg(crow + (j < c1 ? j : c0));
        }
#pragma unroll
        for (int q = 0; q < CL_U; ++q) {
          const int j = c0 + tid + q * T;
          if (j >= c1) continue;
          const int jl = j - c0;
          const double r = ((minVal + (double)cf[q]) - ui) - S.v[jl];        // v[j] is exactly 0 for a column never scanned
          if (first) {
            const int st = (S.flags[jl] & 1) ? -(C - j) : (C - j);         // it = C-1-j  ->  it+1 = C-j
            S.colstate[jl] = st;
            const double s = (r < LSAP_INF) ? r : LSAP_INF;
            S.spc[jl] = s;
            if (s < LSAP_INF) {
              Cand c2;
              c2.val = s; c2.st = st; c2.j = j;
              best = better(best, c2);
            }
          } else {
            const int st = S.colstate[jl];
            if (st == 0) continue;
            double s = S.spc[jl];
            if (r < s) { S.path[jl] = i; S.pathstamp[jl] = cur; S.spc[jl] = r; s = r; }
            if (s < LSAP_INF) {
              Cand c2;
              c2.val = s; c2.st = st; c2.j = j;
              best = better(best, c2);
            }
          }
        }
      }
      first = false;
      // block arg-best -> the leader's slot of this CTA
      best = cl_warp_reduce(best);
      if ((tid & 31) == 0) S.part[tid >> 5] = best;
      __syncthreads();
      if (tid < 32) {
        Cand r2;
        r2.val = 0.0; r2.st = 0; r2.j = -1;
        if (tid < T / 32) r2 = S.part[tid];
        r2 = cl_warp_reduce(r2);
        if (tid == 0) {
          const uint32_t a = cl_map(&S.slot[rank], 0u);
          cl_st_u64(a, (unsigned long long)__double_as_longlong(r2.val));
          cl_st_u64(a + 8u, ((unsigned long long)(uint32_t)r2.j << 32) | (uint32_t)r2.st);
        }
      }
      cl_sync();                                               // A: every candidate of the step is in the leader
      if (leader && tid == 0) {
        Cand g = S.slot[0];
#pragma unroll
        for (int r = 1; r < CL_N; ++r) g = better(g, S.slot[r]);
        ClBcast b;
        b.pad = 0; b.ui = 0.0;
        if (g.st == 0) {
          b.err = 1; b.j = -1; b.next_i = -1; b.minVal = LSAP_INF;
        } else {
          const int j = g.j;
          const int idx = (g.st > 0 ? g.st : -g.st) - 1;                    // position of j in `remaining`
          cl_st_u32(cols.a_colstate(j), 0u);
          if (nsc < CL_SC) S.sc[nsc] = j; else w.sc_list[nsc] = j;
          const int last = nrem - 1;                                         // swap-with-last removal
          const int jm = (S.remstamp[last] == (uint16_t)cur) ? (int)S.remaining[last] : (C - 1 - last);
          if (jm != j) {
            S.remaining[idx] = (uint16_t)jm;
            S.remstamp[idx] = (uint16_t)cur;
            const uint32_t am = cols.a_colstate(jm);
            const int sm = (int)cl_ld_u32(am);
            cl_st_u32(am, (uint32_t)(sm > 0 ? (idx + 1) : -(idx + 1)));
          }
          b.err = 0; b.minVal = g.val; b.j = j;
          b.next_i = g.st > 0 ? -1 : (int)cl_ld_u32(cols.a_row4col(j));
          if (b.next_i >= 0) b.ui = S.u[b.next_i];
        }
        cl_bcast(S, b);
      }
      cl_sync();                                               // B: broadcast record and the two column-state updates are visible
      if (S.bc.err) return S.bc.err;
      minVal = S.bc.minVal;
      ++nsc; --nrem;
      if (S.bc.next_i < 0) sink = S.bc.j; else { i = S.bc.next_i; ui = S.bc.ui; }
    }
    // dual variables and augmentation (lsap_core.cuh): the leader's, a handful of entries each, every access shared memory (its own or,
    // through the cluster address space, the column owner's)
    if (leader) {
      if (tid == 0) S.u[cur] += minVal;
      for (int k = tid; k < nsc; k += T) {
        const int j = k < CL_SC ? S.sc[k] : w.sc_list[k];
        const double d = minVal - __longlong_as_double((long long)cl_ld_u64(cols.a_spc(j)));
        const uint32_t av = cols.a_v(j);
        cl_st_u64(av, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)cl_ld_u64(av)) - d));
        const uint32_t af = cols.a_flags(j);
        cl_st_u8(af, cl_ld_u8(af) | 2u);                  // distinct j per k: no two threads touch the same byte
        if (k < nsc - 1) S.u[(int)cl_ld_u32(cols.a_row4col(j))] += d;
      }
      __syncthreads();
      if (tid == 0) {                                    // augment along the path (<= cur+1 hops)
        int j = sink, hops = 0;
        const uint32_t af = cols.a_flags(sink);
        cl_st_u8(af, cl_ld_u8(af) | 1u);
        int err = 0;
        for (;;) {
          const int r = ((int)cl_ld_u32(cols.a_pathstamp(j)) == cur) ? (int)cl_ld_u32(cols.a_path(j)) : cur;
          cl_st_u32(cols.a_row4col(j), (uint32_t)r);
          const int t = S.col4row[r];
          S.col4row[r] = j;
          j = t;
          if (r == cur) break;
          if (++hops > R || j < 0) { err = 3; break; }
        }
        if (err) {
#pragma unroll
          for (int r = 0; r < CL_N; ++r) cl_st_u32(cl_map(&S.bc, (uint32_t)r) + 16u, (uint32_t)err);
        }
      }
    }
    __syncthreads();
    cl_sync();                                                 // C: duals, flags and the assignment are visible to the next scan
    if (S.bc.err) return S.bc.err;
  }
  return 0;
}

// hungarian_assigner.py:229-270 for one image on a cluster (see hungarian_v2_image in lsap_core.cuh)
__device__ __forceinline__ int hungarian_v2_image_cl(ClShared& S, const uint32_t rank, const float* cost, int N, int n, int topk_k, const Ws& w,
                                                     const int32_t* row_idx, int64_t* out) {
  const int tid = threadIdx.x;
  constexpr int T = CL_T;
  const bool leader = rank == 0;
  if (leader) for (int p = tid; p < N; p += T) w.freelist[p] = p;
  __syncthreads();
  cl_sync();
  int nfree = N;
  const float* Tm = w.T;
  for (int round = 0; round < topk_k; ++round) {
    if (topk_k > 1 && nfree < n) break;                 // `cost_new.shape[0] // num_gts != 0`
    const bool transposed = n < nfree;                  // scipy: transpose iff more rows than columns
    const int R = transposed ? n : nfree, C = transposed ? nfree : n;
    { const int rc = solve_cl(S, rank, cost, Tm, w, N, n, R, C, transposed); if (rc) return rc; }
    if (leader) {
      for (int i = tid; i < R; i += T) {
        const int p = transposed ? w.freelist[S.col4row[i]] : w.freelist[i];
        const int g = transposed ? i : S.col4row[i];
        out[row_idx ? row_idx[p] : p] = (int64_t)g + 1;
      }
      __syncthreads();
    }
    if (!transposed) { nfree = 0; continue; }           // every free proposal got a GT (uniform over the cluster)
    if (leader) {
      for (int i = tid; i < R; i += T) w.freelist[S.col4row[i]] = -1;
      __syncthreads();
      {                                                 // ordered compaction of the free list (see lsap_core.cuh): the leader's 512 threads
        const int seg = (nfree + T - 1) / T;
        const int e0 = min(nfree, tid * seg), e1 = min(nfree, e0 + seg);
        int cnt = 0;
        for (int e = e0; e < e1; ++e) cnt += (w.freelist[e] >= 0) ? 1 : 0;
        const int lane = tid & 31, wi = tid >> 5;
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        if (lane == 31) S.scan[wi] = inc;
        __syncthreads();
        if (wi == 0) {
          const int t = lane < T / 32 ? S.scan[lane] : 0;
          int ti = t;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += x;
          }
          S.scan[lane] = ti - t;
          if (lane == 31) S.scan[32] = ti;
        }
        __syncthreads();
        const int total = S.scan[32];
        int o = S.scan[wi] + inc - cnt;
        for (int e = e0; e < e1; ++e) {
          const int val = w.freelist[e];
          if (val >= 0) w.sc_list[o++] = val;
        }
        __syncthreads();
        for (int e = tid; e < total; e += T) w.freelist[e] = w.sc_list[e];
        if (tid == 0) {
#pragma unroll
          for (int r = 0; r < CL_N; ++r) cl_st_u32(cl_map(&S.bc, (uint32_t)r) + 8u, (uint32_t)total);     // bc.j = new nfree
        }
      }
    }
    __syncthreads();
    cl_sync();
    nfree = S.bc.j;
    // the next round's matrix, columns = the proposals still free: T2[g][e] = T[g][freelist[e]] — rows dealt to the CTAs; every Dijkstra
    // step of the round then reads its cost row directly (coalesced, no indirection)
    if (round + 1 < topk_k && n < nfree) {
      for (int e0 = tid; e0 < nfree; e0 += 8 * T) {
        int fe[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int e = e0 + q * T; fe[q] = __ldcg(&w.freelist[e < nfree ? e : e0]); }
        for (int g = (int)rank; g < n; g += CL_N) {
          const float* s0 = w.T + (size_t)g * (size_t)N;
          float* d0 = w.T2 + (size_t)g * (size_t)N;
          float a0[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) a0[q] = s0[fe[q]];
#pragma unroll
          for (int q = 0; q < 8; ++q) { const int e = e0 + q * T; if (e < nfree) d0[e] = a0[q]; }
        }
      }
      Tm = w.T2;
    }
    __syncthreads();
    cl_sync();
  }
  return 0;
}

}  // namespace ptb_lsap
