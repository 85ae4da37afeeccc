// Thin SVD on the device for the callers of the hot path: `tensor_svd_thin`
// (src/Operations/tensor_svd.jl:100-124: permutedims -> reshape -> LinearAlgebra.svd -> tensorify) and through it
// `simple_update` (src/Operations/simple_update.jl:35-82). The reference calls LAPACK (host) or cuTensorNet
// `gateSplit!` (ext/MusclecuTensorNetExt.jl); this is a hand-written one-sided Jacobi (Hestenes) SVD:
//
//   G <- A (m x n, m >= n; a wide matrix is factorised through its conjugate transpose), V <- I.
//   Sweep: every column pair (p, q) once, in round-robin tournament order: n/2 disjoint pairs per step, one WARP per
//   pair, all steps of a sweep separated by a grid-wide barrier (cooperative launch, every CTA resident).
//   Pair update: alpha = |g_p|^2, beta = |g_q|^2, gamma = g_p^H g_q; if |gamma| > tol sqrt(alpha beta) the plane
//   rotation that makes the two columns orthogonal is applied to G and V.
//   Converged when a sweep rotates nothing: sigma_j = |g_j|, u_j = g_j / sigma_j, sorted descending by a rank count.
//
// One-sided Jacobi is backward stable and computes small singular values to high RELATIVE accuracy; columns stay
// contiguous, so every access is a coalesced lane-strided walk down one or two columns (L2 resident at tensor-network
// bond sizes). Cost per sweep: 3 m n^2 complex flops, n - 1 grid barriers.
//
// Measured (ncu, 128 x 128 ComplexF64): ~11 sweeps to converge, 4.3 us per tournament step, and the step is an
// instruction-latency chain inside one warp (720 warp instructions per step, issue slots 19 % busy), not a barrier or
// a memory problem: a single-cluster variant with W = [G; V] in distributed shared memory and cluster barriers was
// built, verified and measured SLOWER (7.3 vs 5.4 ms at 128^2, 24.8 vs 12.8 ms at 256^2 ComplexF32: 64 warps instead of
// a warp per pair) and was dropped. The next lever is more lanes per pair (a CTA per pair with a shared-memory
// reduction) or a blocked Jacobi whose inner products run on the DMMA gather-GEMM.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace mb200 {

namespace {

template <typename T> struct SvdTraits;
template <> struct SvdTraits<float> { using R = float; static constexpr bool cplx = false; };
template <> struct SvdTraits<double> { using R = double; static constexpr bool cplx = false; };
template <> struct SvdTraits<float2> { using R = float; static constexpr bool cplx = true; };
template <> struct SvdTraits<double2> { using R = double; static constexpr bool cplx = true; };

__device__ __forceinline__ float conj_(float a) { return a; }
__device__ __forceinline__ double conj_(double a) { return a; }
__device__ __forceinline__ float2 conj_(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ double2 conj_(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ float abs2(float a) { return a * a; }
__device__ __forceinline__ double abs2(double a) { return a * a; }
__device__ __forceinline__ float abs2(float2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double abs2(double2 a) { return a.x * a.x + a.y * a.y; }
// conj(a) * b accumulated into (re, im)
__device__ __forceinline__ void cdot(float a, float b, float &re, float &) { re = fmaf(a, b, re); }
__device__ __forceinline__ void cdot(double a, double b, double &re, double &) { re = fma(a, b, re); }
__device__ __forceinline__ void cdot(float2 a, float2 b, float &re, float &im) {
    re = fmaf(a.x, b.x, re); re = fmaf(a.y, b.y, re);
    im = fmaf(a.x, b.y, im); im = fmaf(-a.y, b.x, im);
}
__device__ __forceinline__ void cdot(double2 a, double2 b, double &re, double &im) {
    re = fma(a.x, b.x, re); re = fma(a.y, b.y, re);
    im = fma(a.x, b.y, im); im = fma(-a.y, b.x, im);
}
// x' = c x - s conj(ph) y ; y' = s ph x + c y      (ph = gamma / |gamma|, real case: ph = +-1)
template <typename R> __device__ __forceinline__ void rot(R &x, R &y, R c, R s, R phr, R) {
    const R nx = c * x - s * phr * y, ny = s * phr * x + c * y;
    x = nx; y = ny;
}
__device__ __forceinline__ void rot(float2 &x, float2 &y, float c, float s, float phr, float phi) {
    // conj(ph) y = (phr y.x + phi y.y, phr y.y - phi y.x) ; ph x = (phr x.x - phi x.y, phr x.y + phi x.x)
    const float2 nx = make_float2(c * x.x - s * (phr * y.x + phi * y.y), c * x.y - s * (phr * y.y - phi * y.x));
    const float2 ny = make_float2(s * (phr * x.x - phi * x.y) + c * y.x, s * (phr * x.y + phi * x.x) + c * y.y);
    x = nx; y = ny;
}
__device__ __forceinline__ void rot(double2 &x, double2 &y, double c, double s, double phr, double phi) {
    const double2 nx = make_double2(c * x.x - s * (phr * y.x + phi * y.y), c * x.y - s * (phr * y.y - phi * y.x));
    const double2 ny = make_double2(s * (phr * x.x - phi * x.y) + c * y.x, s * (phr * x.y + phi * x.x) + c * y.y);
    x = nx; y = ny;
}
template <typename T, typename R> __device__ __forceinline__ T scale(T a, R f);
template <> __device__ __forceinline__ float scale(float a, float f) { return a * f; }
template <> __device__ __forceinline__ double scale(double a, double f) { return a * f; }
template <> __device__ __forceinline__ float2 scale(float2 a, float f) { return make_float2(a.x * f, a.y * f); }
template <> __device__ __forceinline__ double2 scale(double2 a, double f) { return make_double2(a.x * f, a.y * f); }
template <typename T> __device__ __forceinline__ T one_of();
template <> __device__ __forceinline__ float one_of<float>() { return 1.f; }
template <> __device__ __forceinline__ double one_of<double>() { return 1.0; }
template <> __device__ __forceinline__ float2 one_of<float2>() { return make_float2(1.f, 0.f); }
template <> __device__ __forceinline__ double2 one_of<double2>() { return make_double2(1.0, 0.0); }
template <typename T> __device__ __forceinline__ T zero_of_();
template <> __device__ __forceinline__ float zero_of_<float>() { return 0.f; }
template <> __device__ __forceinline__ double zero_of_<double>() { return 0.0; }
template <> __device__ __forceinline__ float2 zero_of_<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero_of_<double2>() { return make_double2(0.0, 0.0); }

// 1 / sqrt(x): the FP64 rsqrt is accurate to the last bits; rsqrtf is a 2-ulp approximation that makes c^2 + s^2 drift
// from 1 over thousands of rotations (ComplexF32 singular values 1e-5 -> 5e-5), so Float32 takes the rounded form
__device__ __forceinline__ double inv_sqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float inv_sqrt(float x) { return 1.0f / sqrtf(x); }

__device__ __forceinline__ float max_abs_part(float a) { return fabsf(a); }
__device__ __forceinline__ double max_abs_part(double a) { return fabs(a); }
__device__ __forceinline__ float max_abs_part(float2 a) { return fmaxf(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ double max_abs_part(double2 a) { return fmax(fabs(a.x), fabs(a.y)); }
template <typename T, typename R> __device__ __forceinline__ T make_cplx(R re, R im);
template <> __device__ __forceinline__ float make_cplx<float, float>(float re, float) { return re; }
template <> __device__ __forceinline__ double make_cplx<double, double>(double re, double) { return re; }
template <> __device__ __forceinline__ float2 make_cplx<float2, float>(float re, float im) { return make_float2(re, im); }
template <> __device__ __forceinline__ double2 make_cplx<double2, double>(double re, double im) { return make_double2(re, im); }
// a - c * u
__device__ __forceinline__ float sub_mul(float a, float c, float u) { return fmaf(-c, u, a); }
__device__ __forceinline__ double sub_mul(double a, double c, double u) { return fma(-c, u, a); }
__device__ __forceinline__ float2 sub_mul(float2 a, float2 c, float2 u) {
    return make_float2(a.x - (c.x * u.x - c.y * u.y), a.y - (c.x * u.y + c.y * u.x));
}
__device__ __forceinline__ double2 sub_mul(double2 a, double2 c, double2 u) {
    return make_double2(a.x - (c.x * u.x - c.y * u.y), a.y - (c.x * u.y + c.y * u.x));
}

template <typename R> __device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct SvdParams {
    const void *A;      // rows x cols, dense column-major
    void *G, *V;        // work: G (m x n), V (n x n)      [m >= n after the optional conjugate transpose]
    void *U, *Vt;       // outputs: U (rows x k), Vt (cols x k) = conj(right singular vectors), k = min(rows, cols)
    void *S;            // k singular values (real), descending
    int *counters;      // [0] rotations of the current sweep, [1] sweeps done, [2] scaling exponent + bias, [3] converged,
                        // [4] completed (null) columns; [5..7] scratch (completion norm bits)
    int rows, cols, m, n, transposed;
    double tol;
    int max_sweeps;
    int warps_per_pair;   // 1, 2, 4 or 8 warps of one CTA cooperate on a column pair (long columns: more lanes per pair)
};

template <typename T>
__global__ void __launch_bounds__(256) jacobi_svd_kernel(const __grid_constant__ SvdParams p) {
    using R = typename SvdTraits<T>::R;
    cg::grid_group grid = cg::this_grid();
    const int m = p.m, n = p.n;
    const T *A = reinterpret_cast<const T *>(p.A);
    T *G = reinterpret_cast<T *>(p.G), *V = reinterpret_cast<T *>(p.V);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int64_t warp = tid >> 5, nwarps = nthreads >> 5;

    // ---- pre-scale: alpha, beta and gamma^2 are 2nd and 4th powers of the data, so a matrix whose entries sit near 1e-11 or
    // 1e+9 in Float32 (1e-77 / 1e+77 in Float64) would under- or overflow them and every rotation would be skipped. The working
    // copy is A * 2^-e with e = exponent of max |a_ij| (an exact scaling); the singular values are scaled back at the end.
    constexpr int EBIAS = 4096;
    if (tid < 8) p.counters[tid] = 0;
    grid.sync();
    {
        int emax = -EBIAS;
        for (int64_t e = tid; e < (int64_t)m * n; e += nthreads) {
            const R a = max_abs_part(A[e]);
            if (a > (R)0 && a == a) emax = max(emax, (int)ilogb(a));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
        if (lane == 0 && emax > -EBIAS) atomicMax(&p.counters[2], emax + EBIAS);
    }
    grid.sync();
    const int escale = p.counters[2] > 0 ? p.counters[2] - EBIAS : 0;
    const R down = (R)ldexp(1.0, -escale), up = (R)ldexp(1.0, escale);
    // ---- init: G = A or A^H (scaled), V = I
    for (int64_t e = tid; e < (int64_t)m * n; e += nthreads) {
        const int i = (int)(e % m), j = (int)(e / m);
        G[e] = scale<T, R>(p.transposed ? conj_(A[(int64_t)i * p.rows + j]) : A[e], down);
    }
    for (int64_t e = tid; e < (int64_t)n * n; e += nthreads) V[e] = (e % n == e / n) ? one_of<T>() : zero_of_<T>();
    grid.sync();

    const int np = (n + 1) & ~1;          // players of the tournament (one dummy when n is odd)
    const R tol = (R)p.tol;
    bool converged = n <= 1;
    // A column pair is handled by a GROUP of W warps of one CTA (W = 1: the original warp-per-pair form). Round 1 measured the
    // step as a latency chain inside one warp with 3.5 warps per SM at n = 1024; W warps per pair put W times more loads in flight.
    const int W = p.warps_per_pair;
    const int wic = threadIdx.x >> 5, gic = wic / W, wig = wic % W;
    const int64_t group = (int64_t)blockIdx.x * (8 / W) + gic, ngroups = (int64_t)gridDim.x * (8 / W);
    const int glane = wig * 32 + lane, gsize = W * 32;
    __shared__ R red[2][8][4];
    for (int sweep = 0; sweep < p.max_sweeps && n > 1; sweep++) {
        for (int r = 0; r < np - 1; r++) {
            int it = 0;
            for (int64_t k = group; k < np / 2; k += ngroups, it++) {
                int a, b;
                if (k == 0) { a = np - 1; b = r; }
                else { a = (int)((r + k) % (np - 1)); b = (int)((r - k + (np - 1)) % (np - 1)); }
                const int pc = min(a, b), qc = max(a, b);
                if (qc >= n) continue;    // the dummy player sits out (uniform over the group)
                T *gp = G + (int64_t)pc * m, *gq = G + (int64_t)qc * m;
                R alpha = 0, beta = 0, gre = 0, gim = 0;
                for (int i = glane; i < m; i += gsize) {
                    const T x = gp[i], y = gq[i];
                    alpha += abs2(x); beta += abs2(y);
                    cdot(x, y, gre, gim);
                }
                alpha = warp_sum(alpha); beta = warp_sum(beta); gre = warp_sum(gre); gim = warp_sum(gim);
                if (W > 1) {              // the W warps' partial sums, added in a fixed order by every thread (identical results)
                    if (lane == 0) { red[it & 1][wic][0] = alpha; red[it & 1][wic][1] = beta; red[it & 1][wic][2] = gre; red[it & 1][wic][3] = gim; }
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + gic), "r"(gsize) : "memory");
                    alpha = beta = gre = gim = 0;
                    for (int w = 0; w < W; w++) {
                        alpha += red[it & 1][gic * W + w][0]; beta += red[it & 1][gic * W + w][1];
                        gre += red[it & 1][gic * W + w][2]; gim += red[it & 1][gic * W + w][3];
                    }
                }
                const R g2 = gre * gre + gim * gim;
                if (!(g2 > tol * tol * alpha * beta) || g2 == (R)0) continue;
                // four slow operations (rsqrt, sqrt, div, rsqrt) instead of eight: FP64 sqrt / div are ~40-instruction
                // sequences and this scalar chain is a third of a step's instructions (ncu: 720 per warp per step)
                const R inv_g = inv_sqrt(g2);
                const R zeta = (R)0.5 * (beta - alpha) * inv_g;
                const R t = (zeta >= 0 ? (R)1 : (R)-1) / (fabs(zeta) + sqrt((R)1 + zeta * zeta));
                const R c = inv_sqrt((R)1 + t * t), s = c * t;
                const R phr = gre * inv_g, phi = gim * inv_g;
                for (int i = glane; i < m; i += gsize) {
                    T x = gp[i], y = gq[i];
                    rot(x, y, c, s, phr, phi);
                    gp[i] = x; gq[i] = y;
                }
                T *vp = V + (int64_t)pc * n, *vq = V + (int64_t)qc * n;
                for (int i = glane; i < n; i += gsize) {
                    T x = vp[i], y = vq[i];
                    rot(x, y, c, s, phr, phi);
                    vp[i] = x; vq[i] = y;
                }
                if (glane == 0) atomicAdd(&p.counters[0], 1);
            }
            grid.sync();
        }
        const int rotated = *(volatile int *)&p.counters[0];
        grid.sync();
        if (tid == 0) { p.counters[0] = 0; p.counters[1] = sweep + 1; }
        grid.sync();
        if (rotated == 0) { converged = true; break; }
    }
    if (tid == 0) p.counters[3] = converged ? 1 : 0;   // read back by mb200_svd_last_info

    // ---- singular values: column norms, first in column order
    R *S = reinterpret_cast<R *>(p.S);
    for (int64_t j = warp; j < n; j += nwarps) {
        const T *g = G + (int64_t)j * m;
        R a = 0;
        for (int i = lane; i < m; i += 32) a += abs2(g[i]);
        a = warp_sum(a);
        if (lane == 0) S[j] = sqrt(a);
    }
    grid.sync();
    // rank of column j among the norms (descending, ties by index), then scatter the normalised vectors
    T *U = reinterpret_cast<T *>(p.U), *Vt = reinterpret_cast<T *>(p.Vt);
    for (int64_t j = warp; j < n; j += nwarps) {
        const R sj = S[j];
        int rank = 0;
        for (int i = lane; i < n; i += 32) {
            const R si = S[i];
            rank += (si > sj || (si == sj && i < j)) ? 1 : 0;
        }
        rank = (int)warp_sum((R)rank);
        const R inv = sj > (R)0 ? (R)1 / sj : (R)0;
        const T *g = G + (int64_t)j * m, *v = V + (int64_t)j * n;
        // m >= n: left vectors are the normalised columns of G (m long), right vectors the columns of V (n long).
        // transposed: A = V' S U'^H, so U <- V' and the right vectors <- U'.
        T *left = p.transposed ? Vt : U;     // receives the m-long vectors
        T *right = p.transposed ? U : Vt;    // receives the n-long vectors
        for (int i = lane; i < m; i += 32) {
            const T u = scale<T, R>(g[i], inv);
            left[(int64_t)rank * m + i] = p.transposed ? conj_(u) : u;
        }
        for (int i = lane; i < n; i += 32) right[(int64_t)rank * n + i] = p.transposed ? v[i] : conj_(v[i]);
    }
    grid.sync();
    // S sorted: every thread recomputes the rank of its entry (n is small) into a register, then writes after a barrier
    R mine = 0;
    int myrank = -1;
    if (tid < n) {
        mine = S[tid];
        int rank = 0;
        for (int i = 0; i < n; i++) {
            const R si = S[i];
            rank += (si > mine || (si == mine && i < (int)tid)) ? 1 : 0;
        }
        myrank = rank;
    }
    grid.sync();
    if (myrank >= 0) S[myrank] = mine;
    grid.sync();

    // ---- null columns. A column whose norm is zero (or lost in the subnormal range) was written as a zero vector above, so U
    // (or Vt) would not be isometric for rank-deficient input - product states, low-entanglement bonds - while LAPACK, which
    // tensor_svd_thin / simple_update rely on (src/Operations/tensor_svd.jl:113), always returns orthonormal vectors. They are
    // the LAST columns (S is sorted): each is replaced by a pseudo-random vector orthogonalised against all the columns before
    // it (classical Gram-Schmidt, three passes: coefficients by one warp per column, update by one thread per row).
    {
        T *left = p.transposed ? reinterpret_cast<T *>(p.Vt) : reinterpret_cast<T *>(p.U);
        const R tiny = sizeof(R) == 4 ? (R)2e-18 : (R)2.4e-153;   // ~16 sqrt(smallest normal): below it alpha, beta lose all bits
        int first_null = n;
        for (int j = n - 1; j >= 0 && S[j] < tiny; j--) first_null = j;
        T *coef = V;                                   // V (n x n work) is free now: its first n entries hold the coefficients
        for (int d = first_null; d < n; d++) {
            T *v = left + (int64_t)d * m;
            for (int64_t i = tid; i < m; i += nthreads) {
                uint32_t hsh = (uint32_t)i * 2654435761u ^ ((uint32_t)d * 40503u + 0x9E3779B9u);
                hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
                v[i] = scale<T, R>(one_of<T>(), (R)((int)(hsh & 0xFFFF) - 32768) / (R)32768 + (R)(i == d % m ? 2 : 0));
            }
            grid.sync();
            for (int pass = 0; pass < 3; pass++) {
                for (int64_t c = warp; c < d; c += nwarps) {
                    const T *u = left + c * m;
                    R re = 0, im = 0;
                    for (int i = lane; i < m; i += 32) cdot(u[i], v[i], re, im);
                    re = warp_sum(re); im = warp_sum(im);
                    if (lane == 0) coef[c] = make_cplx<T, R>(re, im);
                }
                grid.sync();
                for (int64_t i = tid; i < m; i += nthreads) {
                    T acc = v[i];
                    for (int c = 0; c < d; c++) acc = sub_mul(acc, coef[c], left[(int64_t)c * m + i]);
                    v[i] = acc;
                }
                grid.sync();
            }
            // normalise (every warp recomputes the norm: no extra barrier for a broadcast)
            R a = 0;
            for (int i = lane; i < m; i += 32) a += abs2(v[i]);
            a = warp_sum(a);
            const R inv = a > (R)0 ? inv_sqrt(a) : (R)0;
            grid.sync();
            for (int64_t i = tid; i < m; i += nthreads) v[i] = scale<T, R>(v[i], inv);
            grid.sync();
        }
        if (tid == 0) p.counters[4] = n - first_null;
    }
    // singular values back to A's scale
    for (int64_t j = tid; j < n; j += nthreads) S[j] *= up;
}


// ---- thin QR (Householder) --------------------------------------------------------------------------------
// `tensor_qr_thin` (src/Operations/tensor_qr.jl:57-79: permutedims -> reshape -> LinearAlgebra.qr -> Matrix(F.Q),
// Matrix(F.R)). A (m x n) = Q (m x k) R (k x n), k = min(m, n). Cooperative kernel, one warp per column:
//   for j < k:  warp 0 builds the reflector H_j = I - tau v v^H from W[j:, j] (v stored in place, R_jj kept aside);
//               grid barrier; every trailing column c > j gets W[j:, c] -= tau v (v^H W[j:, c]); grid barrier
//   Q = H_0 ... H_{k-1} [I; 0]: reflectors applied backwards to the columns c >= j that can be non-trivial.
// Reflectors are Hermitian (real tau = 2 / v^H v), so R's diagonal is -e^{i arg x_0} |x|: QR is unique up to that
// diagonal phase, which is all the reference's tests pin (Q R = A, Q isometric).
struct QrParams {
    const void *A;
    void *W;        // m x n work copy
    void *Q, *Rm;   // m x k, k x n
    void *rd;       // k diagonal entries of R
    void *tau;      // k reals
    int m, n, k;
};

template <typename T>
__global__ void __launch_bounds__(256) householder_qr_kernel(const __grid_constant__ QrParams p) {
    using R = typename SvdTraits<T>::R;
    cg::grid_group grid = cg::this_grid();
    const int m = p.m, n = p.n, k = p.k;
    const T *A = reinterpret_cast<const T *>(p.A);
    T *W = reinterpret_cast<T *>(p.W), *Q = reinterpret_cast<T *>(p.Q), *Rm = reinterpret_cast<T *>(p.Rm);
    T *rd = reinterpret_cast<T *>(p.rd);
    R *tau = reinterpret_cast<R *>(p.tau);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int64_t warp = tid >> 5, nwarps = nthreads >> 5;
    for (int64_t e = tid; e < (int64_t)m * n; e += nthreads) W[e] = A[e];
    for (int64_t e = tid; e < (int64_t)m * k; e += nthreads) Q[e] = (e % m == e / m) ? one_of<T>() : zero_of_<T>();
    grid.sync();
    for (int j = 0; j < k; j++) {
        if (warp == 0) {   // reflector from x = W[j:, j]
            T *x = W + (int64_t)j * m;
            R nrm2 = 0;
            for (int i = j + lane; i < m; i += 32) nrm2 += abs2(x[i]);
            nrm2 = warp_sum(nrm2);
            const T x0 = x[j];
            const R ax0 = sqrt(abs2(x0)), nrm = sqrt(nrm2);
            if (lane == 0) {
                if (nrm > (R)0) {
                    // e^{i theta} = x0 / |x0| (1 when x0 == 0); v = x + e^{i theta} |x| e_1; H x = -e^{i theta} |x| e_1
                    const T ph = ax0 > (R)0 ? scale<T, R>(x0, (R)1 / ax0) : one_of<T>();
                    tau[j] = (R)1 / (nrm * (nrm + ax0));
                    rd[j] = scale<T, R>(ph, -nrm);
                    x[j] = scale<T, R>(ph, ax0 + nrm);
                } else {
                    tau[j] = (R)0;
                    rd[j] = zero_of_<T>();
                }
            }
        }
        grid.sync();
        const R tj = tau[j];
        const T *v = W + (int64_t)j * m;
        if (tj != (R)0) {
            for (int64_t c = j + 1 + warp; c < n; c += nwarps) {
                T *w = W + c * m;
                R re = 0, im = 0;
                for (int i = j + lane; i < m; i += 32) cdot(v[i], w[i], re, im);
                re = warp_sum(re) * tj; im = warp_sum(im) * tj;
                for (int i = j + lane; i < m; i += 32) {   // w -= tau (v^H w) v
                    T vi = v[i], wi = w[i];
                    if constexpr (SvdTraits<T>::cplx) { wi.x -= vi.x * re - vi.y * im; wi.y -= vi.x * im + vi.y * re; }
                    else wi -= vi * re;
                    w[i] = wi;
                }
            }
        }
        grid.sync();
    }
    // R: upper triangle of W with the saved diagonal
    for (int64_t e = tid; e < (int64_t)k * n; e += nthreads) {
        const int i = (int)(e % k), c = (int)(e / k);
        Rm[e] = i < c ? W[(int64_t)c * m + i] : (i == c ? rd[i] : zero_of_<T>());
    }
    // Q = H_0 ... H_{k-1} [I; 0]
    for (int j = k - 1; j >= 0; j--) {
        const R tj = tau[j];
        const T *v = W + (int64_t)j * m;
        if (tj != (R)0) {
            for (int64_t c = j + warp; c < k; c += nwarps) {
                T *q = Q + c * m;
                R re = 0, im = 0;
                for (int i = j + lane; i < m; i += 32) cdot(v[i], q[i], re, im);
                re = warp_sum(re) * tj; im = warp_sum(im) * tj;
                for (int i = j + lane; i < m; i += 32) {
                    T vi = v[i], qi = q[i];
                    if constexpr (SvdTraits<T>::cplx) { qi.x -= vi.x * re - vi.y * im; qi.y -= vi.x * im + vi.y * re; }
                    else qi -= vi * re;
                    q[i] = qi;
                }
            }
        }
        grid.sync();
    }
}

}  // namespace

cudaError_t launch_svd(int dtype, const void *A, int rows, int cols, void *U, void *S, void *Vt, void *work_G, void *work_V,
                       int *counters, double tol, int max_sweeps, cudaStream_t s) {
    SvdParams p{};
    p.A = A; p.G = work_G; p.V = work_V; p.U = U; p.Vt = Vt; p.S = S; p.counters = counters;
    p.rows = rows; p.cols = cols;
    p.transposed = rows < cols ? 1 : 0;
    p.m = std::max(rows, cols); p.n = std::min(rows, cols);
    p.tol = tol; p.max_sweeps = max_sweeps;
    const void *fn;
    switch (dtype) {
        case MB200_F32: fn = (const void *)jacobi_svd_kernel<float>; break;
        case MB200_F64: fn = (const void *)jacobi_svd_kernel<double>; break;
        case MB200_C64: fn = (const void *)jacobi_svd_kernel<float2>; break;
        default: fn = (const void *)jacobi_svd_kernel<double2>; break;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // W warps per column pair (W | 8, a CTA holds 8 / W pairs): as many as keep every pair of a tournament step resident at once
    // and give every lane at least one row; never more CTAs than can be resident (cooperative launch). MB200_SVD_WARPS pins W.
    const int64_t pairs = ((int64_t)p.n + 1) / 2;
    const int64_t resident = (int64_t)sms * std::max(1, per_sm);
    static const int forced = [] { const char *e = getenv("MB200_SVD_WARPS"); return e ? atoi(e) : 0; }();
    int W = 1;
    while (W < 8 && (pairs * (2 * W) + 7) / 8 <= resident && p.m >= 64 * W) W *= 2;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) W = forced;
    p.warps_per_pair = W;
    const int64_t want = std::max<int64_t>(1, (pairs * W + 7) / 8);
    const int grid = (int)std::min<int64_t>(want, resident);
    void *args[] = {(void *)&p};
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, s);
}

}  // namespace mb200

namespace mb200 {

cudaError_t launch_qr(int dtype, const void *A, int m, int n, void *Q, void *Rm, void *work_W, void *work_rd, void *work_tau,
                      cudaStream_t s) {
    QrParams p{};
    p.A = A; p.W = work_W; p.Q = Q; p.Rm = Rm; p.rd = work_rd; p.tau = work_tau;
    p.m = m; p.n = n; p.k = std::min(m, n);
    const void *fn;
    switch (dtype) {
        case MB200_F32: fn = (const void *)householder_qr_kernel<float>; break;
        case MB200_F64: fn = (const void *)householder_qr_kernel<double>; break;
        case MB200_C64: fn = (const void *)householder_qr_kernel<float2>; break;
        default: fn = (const void *)householder_qr_kernel<double2>; break;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = std::max<int64_t>(1, ((int64_t)std::max(n, p.k) + 7) / 8);   // one warp per column
    const int grid = (int)std::min<int64_t>(want, (int64_t)sms * std::max(1, std::min(per_sm, 2)));
    void *args[] = {(void *)&p};
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, s);
}

}  // namespace mb200
