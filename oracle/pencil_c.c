/* pencil_c.c -- C/OpenMP restatement of oracle/pencil.py -- ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Same arithmetic, operation by operation, as the NumPy pencil oracle (which is pinned bit-exactly
 * against the real reference, tests/golden/make_golden.py); tests/test_oracle_c.py pins this file
 * against the NumPy oracle with array_equal.  It exists so that full-size configs (256^3 Chebyshev,
 * 1024^3 Laplacian sub-blocks) can be checked in seconds and so that bench.py's CPU arm can use
 * every host core.  Compile with -O2 -ffp-contract=off (no FMA contraction, no -ffast-math).
 *
 * fd:    findiff.FinDiff apply behind reference numgrids/diff.py:88, :259-262
 * dense: scipy sparse matvec of the Kronecker matrix, numgrids/diff.py:194-197 (ordered sum)
 * lap:   the idiom Diff(g,2,0)(f)+Diff(g,2,1)(f)+Diff(g,2,2)(f), reference tests/test_diff.py:29-39
 */
#include <stdint.h>
#include <stddef.h>

typedef struct {
    int n_center, n_side;
    const double *wc, *wf, *wb;
    const int32_t *oc, *of, *ob;
    double inv_h_pow;
} po_fd_tables;

static inline double fd_point(const double* p, int64_t stride, int64_t i, int64_t n, const po_fd_tables* T) {
    const int nb = T->n_center / 2;
    const double* w; const int32_t* o; int nt;
    if (i < nb) { w = T->wf; o = T->of; nt = T->n_side; }
    else if (i >= n - nb) { w = T->wb; o = T->ob; nt = T->n_side; }
    else { w = T->wc; o = T->oc; nt = T->n_center; }
    double acc = 0.0;
    for (int k = 0; k < nt; ++k) {
        const double prod = w[k] * p[(int64_t)o[k] * stride];
        acc = acc + prod;
    }
    return acc * T->inv_h_pow;
}

/* out = FinDiff(axis, h, order, acc)(in) [/ x along the axis when xdiv != NULL] on [outer][n][inner] */
void po_fd_apply(const double* in, double* out, int64_t outer, int64_t n, int64_t inner,
                 const po_fd_tables* T, const double* xdiv) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t o = 0; o < outer; ++o)
        for (int64_t i = 0; i < n; ++i) {
            const double* row = in + (o * n + i) * inner;
            double* dst = out + (o * n + i) * inner;
            for (int64_t c = 0; c < inner; ++c) {
                double d = fd_point(row + c, inner, i, n, T);
                if (xdiv) d = d / xdiv[i];
                dst[c] = d;
            }
        }
}

/* out_i = sum_j M[i][j] in_j along the middle axis of [outer][n][inner]; j ascending or descending;
 * accumulate from 0.0; product and sum rounded separately. */
void po_dense_apply(const double* in, double* out, int64_t outer, int64_t n, int64_t inner,
                    const double* M, int descending) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t o = 0; o < outer; ++o)
        for (int64_t i = 0; i < n; ++i) {
            double* dst = out + (o * n + i) * inner;
            for (int64_t c = 0; c < inner; ++c) dst[c] = 0.0;
            for (int64_t q = 0; q < n; ++q) {
                const int64_t j = descending ? n - 1 - q : q;
                const double m = M[i * n + j];
                const double* src = in + (o * n + j) * inner;
                for (int64_t c = 0; c < inner; ++c) {
                    const double prod = m * src[c];
                    dst[c] = dst[c] + prod;
                }
            }
        }
}

/* out = (d0 + d1) + d2 with d_a = FD second derivative along axis a of a [n0][n1][n2] array */
void po_fdlap3(const double* in, double* out, int64_t n0, int64_t n1, int64_t n2,
               const po_fd_tables* T0, const po_fd_tables* T1, const po_fd_tables* T2) {
    const int64_t s0 = n1 * n2, s1 = n2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t i = 0; i < n0; ++i)
        for (int64_t j = 0; j < n1; ++j) {
            const double* row = in + i * s0 + j * s1;
            double* dst = out + i * s0 + j * s1;
            for (int64_t k = 0; k < n2; ++k) {
                const double d0 = fd_point(row + k, s0, i, n0, T0);
                const double d1 = fd_point(row + k, s1, j, n1, T1);
                const double d2 = fd_point(row + k, 1, k, n2, T2);
                dst[k] = (d0 + d1) + d2;
            }
        }
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU arm asks for the host's cores explicitly */
void po_set_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int po_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
