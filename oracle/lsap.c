/* TEST INFRASTRUCTURE — CPU oracle, never on the product path (only tests/, smoke() and bench.py's cpu_baseline may use it).
 *
 * Restatement of scipy.optimize.linear_sum_assignment (scipy 1.18.1 in this image; un-vendored third-party dependency of the
 * reference: TOV_mmdetection/mmdet/core/bbox/assigners/hungarian_assigner.py:10,236,257, unpinned in requirements/optional.txt:4).
 * Published algorithm: D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE T-AES 52(4), 2016 —
 * shortest augmenting paths with dual variables u, v; scipy's implementation (scipy/optimize/rectangular_lsap) adds
 *   - transpose when there are more rows than columns,
 *   - a `remaining` column list filled in REVERSE order and shrunk by swap-with-last,
 *   - the tie rule "among equal lowest reduced costs prefer a column that is still unassigned".
 * Pinned against scipy itself in tests/test_lsap.py (random fp32 costs, tie-heavy small-integer costs, both orientations).
 *
 * Two variants with identical results:
 *   lsap_solve(..., keyed=0): the sequential scan, statement by statement.
 *   lsap_solve(..., keyed=1): the selection written as the arg-max of a total order (value asc, then key2 desc with
 *       key2 = +(it+1) for unassigned columns and -(it+1) for assigned ones) — the form the CUDA kernel reduces in parallel
 *       (pointtinybenchmark_b200/csrc/lsap.cu).  The sequential scan picks, among the columns at the minimum, the LAST
 *       unassigned one in `remaining` order if any, else the FIRST one: exactly that arg-max.
 *
 * hungarian_v2(): the <= topk_k rounds of hungarian_assigner.py:229-270 on top.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cost: nr x nc row-major doubles, nr <= nc.  col4row out (nr).  returns 0 ok, -1 infeasible. */
static int lsap_core(int64_t nr, int64_t nc, const double *cost, int64_t *col4row, int keyed)
{
    if (nr <= 0 || nc <= 0 || nr > ((int64_t)1 << 40) || nc > ((int64_t)1 << 40)) return 0;
    const size_t znr = (size_t)nr, znc = (size_t)nc;
    double *u = calloc(znr, sizeof(double)), *v = calloc(znc, sizeof(double)), *spc = malloc(znc * sizeof(double));
    int64_t *path = malloc(znc * sizeof(int64_t)), *row4col = malloc(znc * sizeof(int64_t)), *remaining = malloc(znc * sizeof(int64_t));
    char *SR = malloc(znr), *SC = malloc(znc);
    int rc = 0;
    for (int64_t j = 0; j < nc; j++) { path[j] = -1; row4col[j] = -1; }
    for (int64_t i = 0; i < nr; i++) col4row[i] = -1;
    for (int64_t cur = 0; cur < nr && rc == 0; cur++) {
        double minVal = 0;
        int64_t i = cur, num_remaining = nc, sink = -1;
        for (int64_t it = 0; it < nc; it++) remaining[it] = nc - it - 1;
        memset(SR, 0, znr); memset(SC, 0, znc);
        for (int64_t j = 0; j < nc; j++) spc[j] = INFINITY;
        while (sink == -1) {
            int64_t index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            if (!keyed) {
                for (int64_t it = 0; it < num_remaining; it++) {
                    int64_t j = remaining[it];
                    double r = minVal + cost[i * nc + j] - u[i] - v[j];
                    if (r < spc[j]) { path[j] = i; spc[j] = r; }
                    if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
                }
            } else {
                int64_t bestkey = 0;          /* 0 = none */
                for (int64_t it = num_remaining - 1; it >= 0; it--) {      /* any visiting order gives the same arg-max */
                    int64_t j = remaining[it];
                    double r = minVal + cost[i * nc + j] - u[i] - v[j];
                    if (r < spc[j]) { path[j] = i; spc[j] = r; }
                    int64_t key2 = row4col[j] == -1 ? (it + 1) : -(it + 1);
                    if (spc[j] < lowest || (spc[j] == lowest && spc[j] < INFINITY && (bestkey == 0 || key2 > bestkey))) {
                        lowest = spc[j]; bestkey = key2;
                    }
                }
                index = bestkey == 0 ? -1 : (bestkey > 0 ? bestkey - 1 : -bestkey - 1);
            }
            minVal = lowest;
            if (minVal == INFINITY) { rc = -1; break; }
            int64_t j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc) break;
        u[cur] += minVal;
        for (int64_t r = 0; r < nr; r++) if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int64_t j = 0; j < nc; j++) if (SC[j]) v[j] -= minVal - spc[j];
        int64_t j = sink;
        for (;;) {
            int64_t r = path[j];
            row4col[j] = r;
            int64_t t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    free(u); free(v); free(spc); free(path); free(row4col); free(remaining); free(SR); free(SC);
    return rc;
}

/* scipy's wrapper: cost nr x nc (fp32, as the reference passes it; scipy converts to double), any orientation.
 * rows_out / cols_out (min(nr,nc)) = the (row_ind, col_ind) pair scipy returns (rows ascending).
 * returns 0, -1 infeasible, -2 NaN / -inf entries. */
int lsap_solve(int64_t nr, int64_t nc, const float *cost, int64_t *rows_out, int64_t *cols_out, int keyed)
{
    if (nr == 0 || nc == 0) return 0;
    for (int64_t e = 0; e < nr * nc; e++) if (cost[e] != cost[e] || cost[e] == -INFINITY) return -2;
    int transpose = nc < nr;
    int64_t R = transpose ? nc : nr, C = transpose ? nr : nc;
    double *m = malloc(sizeof(double) * nr * nc);
    for (int64_t i = 0; i < nr; i++)
        for (int64_t j = 0; j < nc; j++) {
            if (transpose) m[j * nr + i] = cost[i * nc + j]; else m[i * nc + j] = cost[i * nc + j];
        }
    int64_t *col4row = malloc(sizeof(int64_t) * R);
    int rc = lsap_core(R, C, m, col4row, keyed);
    if (rc == 0) {
        if (!transpose) {
            for (int64_t i = 0; i < R; i++) { rows_out[i] = i; cols_out[i] = col4row[i]; }
        } else {
            /* argsort of col4row (distinct values): counting placement */
            int64_t *owner = malloc(sizeof(int64_t) * C);
            for (int64_t j = 0; j < C; j++) owner[j] = -1;
            for (int64_t i = 0; i < R; i++) owner[col4row[i]] = i;
            int64_t k = 0;
            for (int64_t j = 0; j < C; j++) if (owner[j] >= 0) { rows_out[k] = j; cols_out[k] = owner[j]; k++; }
            free(owner);
        }
    }
    free(m); free(col4row);
    return rc;
}

/* hungarian_assigner.py:229-270: cost (N x n) fp32 -> assigned_gt_inds (N) (0 background, g+1 foreground). */
int hungarian_v2(int64_t N, int64_t n, const float *cost, int topk_k, int64_t *gt_inds, int keyed)
{
    for (int64_t p = 0; p < N; p++) gt_inds[p] = 0;
    if (N == 0 || n == 0) return 0;
    int64_t m = N < n ? N : n;
    int64_t *r = malloc(sizeof(int64_t) * (m + 1)), *c = malloc(sizeof(int64_t) * (m + 1));
    int rc = 0;
    if (topk_k == 1) {
        rc = lsap_solve(N, n, cost, r, c, keyed);
        if (rc == 0) for (int64_t k = 0; k < m; k++) gt_inds[r[k]] = c[k] + 1;
    } else {
        int64_t *index = malloc(sizeof(int64_t) * N);
        float *sub = malloc(sizeof(float) * N * n);
        char *taken = calloc(N, 1);
        int num = 0;
        for (;;) {
            int64_t nf = 0;
            for (int64_t p = 0; p < N; p++) if (!taken[p]) { memcpy(sub + nf * n, cost + p * n, sizeof(float) * n); index[nf++] = p; }
            if (nf / n == 0 || num + 1 > topk_k) break;
            num++;
            rc = lsap_solve(nf, n, sub, r, c, keyed);
            if (rc) break;
            int64_t mm = nf < n ? nf : n;
            for (int64_t k = 0; k < mm; k++) { gt_inds[index[r[k]]] = c[k] + 1; taken[index[r[k]]] = 1; }
        }
        free(index); free(sub); free(taken);
    }
    free(r); free(c);
    return rc;
}
