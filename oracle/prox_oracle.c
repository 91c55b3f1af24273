/*
 * oracle/prox_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, float64).
 *
 * Plain-C restatement of the proximal operators that the reference reaches on
 * its hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (background-subtraction_b200) never does.
 *
 * What is restated, and where the reference calls it:
 *   - prox_flat_linf   <- spams.proximalFlat(..., regul='group-lasso-linf')
 *                         called at /root/reference/inexact_alm_lsd.py:71-79
 *   - prox_graph_linf  <- spams.proximalGraph(..., regul='graph')
 *                         called at /root/reference/inexact_alm_lsd.py:49-57
 *                         (and per frame at :60-68)
 *   - block_shrink_l2  <- block_shrinkage_operator,
 *                         /root/reference/group_sparse_RPCA.py:13-42
 *
 * SPAMS itself (Mairal et al., INRIA; C++/OpenMP; version unpinned in the
 * reference -- no requirements file) is an un-vendored third-party dependency
 * that is not installed here.  Its published definitions are restated:
 *   group-lasso-linf :  argmin_v 1/2||u-v||^2 + lam * sum_g ||v_g||_inf
 *                       (non-overlapping groups; per column)
 *   graph            :  argmin_v 1/2||u-v||^2 + lam * sum_g eta_g ||v_g||_inf
 *                       (overlapping groups given by groups_var; per column)
 * PARITY UNPINNED against the real SPAMS binary: the restatement is pinned
 * instead against (i) an independent generic QP solve of the primal problem
 * on small images (tests/test_oracle.py) and (ii) the Moreau identity
 * v = u - Proj_{dual ball}(u).
 *
 * Closed form used: the prox of lam*||.||_inf is  v = u - Proj_{l1-ball(lam)}(u)
 * (Moreau decomposition; the l1 ball is the dual-norm ball of l_inf).
 * For overlapping groups the dual problem
 *     min_{xi^g}  1/2 || u - sum_g xi^g ||^2 ,  supp(xi^g) in g, ||xi^g||_1 <= lam*eta_g
 * is solved by cyclic block-coordinate descent (each block update is an exact
 * l1-ball projection), v = u - sum_g xi^g.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXG 64 /* largest group size handled on the stack */

static int cmp_desc(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x < y) - (x > y);
}

/* Euclidean projection of x[0..k) onto the l1 ball of radius z (z >= 0).
 * Sort-based (Held/Wolfe/Crowder; Duchi et al. 2008). out may alias x. */
static void l1ball_project(const double *x, int k, double z, double *out) {
    double a_stack[MAXG];
    double *a = (k <= MAXG) ? a_stack : (double *)malloc(sizeof(double) * (size_t)k);
    double s = 0.0;
    for (int i = 0; i < k; ++i) { a[i] = fabs(x[i]); s += a[i]; }
    if (s <= z) {
        if (out != x) memcpy(out, x, sizeof(double) * (size_t)k);
        if (a != a_stack) free(a);
        return;
    }
    qsort(a, (size_t)k, sizeof(double), cmp_desc);
    double cs = 0.0, theta = 0.0;
    for (int i = 0; i < k; ++i) {
        cs += a[i];
        double t = (cs - z) / (double)(i + 1);
        if (a[i] - t > 0.0) theta = t; else break;
    }
    for (int i = 0; i < k; ++i) {
        double m = fabs(x[i]) - theta;
        out[i] = (m > 0.0) ? copysign(m, x[i]) : 0.0;
    }
    if (a != a_stack) free(a);
}

/* ---- flat l_inf groups ---------------------------------------------------
 * U, V: m x n, Fortran order (column = frame).  groups: int32[m], 1-based ids,
 * 0 = pixel in no group (copied through).  Restates the call
 * /root/reference/inexact_alm_lsd.py:76-79. */
int prox_flat_linf(const double *U, long m, long n, const int *groups, double lam,
                   double *V, int nthreads) {
    int gmax = 0;
    for (long p = 0; p < m; ++p) { if (groups[p] < 0) return -1; if (groups[p] > gmax) gmax = groups[p]; }
    /* CSR of group -> pixel list */
    long *ptr = (long *)calloc((size_t)gmax + 2, sizeof(long));
    long *idx = (long *)malloc(sizeof(long) * (size_t)(m > 0 ? m : 1));
    for (long p = 0; p < m; ++p) ptr[groups[p] + 1]++;
    for (int g = 0; g <= gmax; ++g) ptr[g + 1] += ptr[g];
    long *fill = (long *)malloc(sizeof(long) * ((size_t)gmax + 1));
    memcpy(fill, ptr, sizeof(long) * ((size_t)gmax + 1));
    for (long p = 0; p < m; ++p) idx[fill[groups[p]]++] = p;
    free(fill);
    long maxsz = 0;
    for (int g = 1; g <= gmax; ++g) if (ptr[g + 1] - ptr[g] > maxsz) maxsz = ptr[g + 1] - ptr[g];
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)(2 * maxsz + 2));
        double *pr = buf + maxsz + 1;
#pragma omp for schedule(static)
        for (long f = 0; f < n; ++f) {
            const double *u = U + f * m;
            double *v = V + f * m;
            for (long q = ptr[0]; q < ptr[1]; ++q) v[idx[q]] = u[idx[q]]; /* id 0: untouched */
            for (int g = 1; g <= gmax; ++g) {
                long b = ptr[g], e = ptr[g + 1];
                int k = (int)(e - b);
                if (k == 0) continue;
                for (int i = 0; i < k; ++i) buf[i] = u[idx[b + i]];
                l1ball_project(buf, k, lam, pr);
                for (int i = 0; i < k; ++i) v[idx[b + i]] = buf[i] - pr[i];
            }
        }
        free(buf);
    }
    free(ptr); free(idx);
    return 0;
}

/* ---- overlapping l_inf groups ("graph") -----------------------------------
 * groups_var given in CSC form exactly as the reference builds it
 * (/root/reference/inexact_alm_lsd.py:31-43): column g lists the pixels of
 * group g (indices[indptr[g]..indptr[g+1]) ), eta[g] its weight.  No group
 * nesting (the reference's 'groups' matrix is empty, :28).
 * Cyclic BCD on the dual, per column, until the largest dual change in a sweep
 * is <= tol or max_sweeps is reached.  sweeps_out[f] = sweeps used (may be NULL). */
int prox_graph_linf(const double *U, long m, long n, const int *indptr, const int *indices,
                    long ngroups, const double *eta, double lam, double tol, int max_sweeps,
                    double *V, int *sweeps_out, int nthreads) {
    long nnz = indptr[ngroups];
    int maxsz = 0;
    for (long g = 0; g < ngroups; ++g) {
        int k = indptr[g + 1] - indptr[g];
        if (k > maxsz) maxsz = k;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *xi = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
        double *tot = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
        double *r = (double *)malloc(sizeof(double) * (size_t)(2 * maxsz + 2));
        double *pr = r + maxsz + 1;
#pragma omp for schedule(dynamic, 1)
        for (long f = 0; f < n; ++f) {
            const double *u = U + f * m;
            double *v = V + f * m;
            memset(xi, 0, sizeof(double) * (size_t)nnz);
            memset(tot, 0, sizeof(double) * (size_t)m);
            int sw = 0;
            for (; sw < max_sweeps; ++sw) {
                double change = 0.0;
                for (long g = 0; g < ngroups; ++g) {
                    int b = indptr[g], k = indptr[g + 1] - b;
                    if (k == 0) continue;
                    for (int i = 0; i < k; ++i) {
                        int p = indices[b + i];
                        r[i] = u[p] - tot[p] + xi[b + i];
                    }
                    l1ball_project(r, k, lam * eta[g], pr);
                    for (int i = 0; i < k; ++i) {
                        int p = indices[b + i];
                        double dlt = pr[i] - xi[b + i];
                        if (fabs(dlt) > change) change = fabs(dlt);
                        tot[p] += dlt;
                        xi[b + i] = pr[i];
                    }
                }
                if (change <= tol) { ++sw; break; }
            }
            for (long p = 0; p < m; ++p) v[p] = u[p] - tot[p];
            if (sweeps_out) sweeps_out[f] = sw;
        }
        free(xi); free(tot); free(r);
    }
    return 0;
}

/* ---- per-frame l2 block shrinkage -----------------------------------------
 * Restates /root/reference/group_sparse_RPCA.py:13-42 with the blocks of one
 * frame given as a label map: labels[f*m + p] = 1-based block index of pixel p
 * in frame f, 0 = complement.  Blocks of a frame are disjoint in the reference
 * (they come from one connected-component label map,
 * /root/reference/motion_saliency_check.py:41-47), so a label map carries the
 * same information as the list of boolean masks.  lam_ptr/lam_val: CSR of the
 * per-frame lambda lists.  The complement of all blocks is one more group with
 * epsilon = non_block_lambda / mu (:37-40).  A zero-norm group gets factor 0
 * (max(1 - eps/0, 0) = max(-inf, 0) = 0, :35). */
int block_shrink_l2(const double *G, long m, long n, const int *labels, const int *lam_ptr,
                    const double *lam_val, double mu, double non_block_lambda, double *R) {
    for (long f = 0; f < n; ++f) {
        const double *g = G + f * m;
        double *r = R + f * m;
        const int *lab = labels + f * m;
        int nb = lam_ptr[f + 1] - lam_ptr[f];
        double *ss = (double *)calloc((size_t)nb + 1, sizeof(double));
        for (long p = 0; p < m; ++p) {
            int l = lab[p];
            if (l < 0 || l > nb) { free(ss); return -1; }
            ss[l] += g[p] * g[p];
        }
        double *fac = (double *)malloc(sizeof(double) * ((size_t)nb + 1));
        for (int l = 0; l <= nb; ++l) {
            double eps = (l == 0) ? non_block_lambda / mu : lam_val[lam_ptr[f] + l - 1] / mu;
            double nrm = sqrt(ss[l]);
            double t = (nrm > 0.0) ? 1.0 - eps / nrm : 0.0;
            fac[l] = t > 0.0 ? t : 0.0;
        }
        for (long p = 0; p < m; ++p) r[p] = fac[lab[p]] * g[p];
        free(ss); free(fac);
    }
    return 0;
}

/* exported for direct unit tests of the projection */
void l1ball_project_export(const double *x, int k, double z, double *out) { l1ball_project(x, k, z, out); }
