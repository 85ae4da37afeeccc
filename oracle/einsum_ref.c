/* TEST INFRASTRUCTURE ONLY — plain-C loop-nest restatement of pairwise einsum.
 *
 * Independent second opinion for the NumPy oracle (oracle/muscle_oracle.py): the same explicit
 * loop nest the reference's own tests use as ground truth for hyperindex contractions
 * (/root/reference/test/integration/omeinsum.jl:225-236, test/integration/cuda.jl:169-180):
 *     C[inds_c] += A[inds_a] * B[inds_b]   over every label.
 * Dense column-major operands (Julia Array layout), complex interleaved (re,im).
 * O(prod of all extents) — small cases only.  Built by oracle/build_oracle.py with gcc.
 */
#include <complex.h>
#include <stdint.h>
#include <string.h>

#define MAXL 32

#define DEFINE_EINSUM(NAME, T)                                                                   \
    int NAME(int nlabels, const int64_t *extent, /* per label id 0..nlabels-1 */                 \
             int na, const int32_t *ma, const T *A, int nb, const int32_t *mb, const T *B,       \
             int nc, const int32_t *mc, T *C)                                                    \
    {                                                                                            \
        if (nlabels > MAXL) return 1;                                                            \
        int64_t sa[MAXL] = {0}, sb[MAXL] = {0}, sc[MAXL] = {0};                                  \
        int64_t s = 1;                                                                           \
        for (int i = 0; i < na; i++) { sa[ma[i]] += s; s *= extent[ma[i]]; }                     \
        s = 1;                                                                                   \
        for (int i = 0; i < nb; i++) { sb[mb[i]] += s; s *= extent[mb[i]]; }                     \
        s = 1;                                                                                   \
        for (int i = 0; i < nc; i++) { sc[mc[i]] += s; s *= extent[mc[i]]; }                     \
        int64_t csize = s;                                                                       \
        for (int64_t i = 0; i < csize; i++) C[i] = 0;                                            \
        int64_t idx[MAXL] = {0};                                                                 \
        int64_t total = 1;                                                                       \
        for (int l = 0; l < nlabels; l++) total *= extent[l];                                    \
        int64_t oa = 0, ob = 0, oc = 0;                                                          \
        for (int64_t it = 0; it < total; it++) {                                                 \
            C[oc] += A[oa] * B[ob];                                                              \
            for (int l = 0; l < nlabels; l++) { /* odometer */                                   \
                idx[l]++; oa += sa[l]; ob += sb[l]; oc += sc[l];                                 \
                if (idx[l] < extent[l]) break;                                                   \
                oa -= sa[l] * extent[l]; ob -= sb[l] * extent[l]; oc -= sc[l] * extent[l];       \
                idx[l] = 0;                                                                      \
            }                                                                                    \
        }                                                                                        \
        return 0;                                                                                \
    }

DEFINE_EINSUM(oracle_einsum_f32, float)
DEFINE_EINSUM(oracle_einsum_f64, double)
DEFINE_EINSUM(oracle_einsum_c64, float complex)
DEFINE_EINSUM(oracle_einsum_c128, double complex)
