// The two single-operand / broadcasting members of Muscle's einsum family that sit next to binary_einsum
// (SURVEY §8f row 2), as HBM-bound streaming kernels for sm_100a:
//
//   unary_einsum  y[out] = sum over the labels of x missing from y, repeated labels of x read on the diagonal
//                 (src/Operations/unary_einsum.jl:26-36 -> OMEinsum `einsum!`, ext/MuscleOMEinsumExt.jl:25-38):
//                 axis sums, traces, diagonals, plain permutations.
//   hadamard      c = a .* broadcast(b), inds(b) a subset of inds(a), c laid out like a
//                 (src/Operations/hadamard.jl:42-77).
//
// Roofline for both: HBM. Algorithmic bytes: unary = |x| + |y|; hadamard = |a| + |b| + |c|.
#include <algorithm>

#include "kernels.cuh"

namespace mb200 {

namespace {

template <typename T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ float zero_of<float>() { return 0.f; }
template <> __device__ __forceinline__ double zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ float2 zero_of<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero_of<double2>() { return make_double2(0.0, 0.0); }

__device__ __forceinline__ void acc_add(float &a, float v) { a += v; }
__device__ __forceinline__ void acc_add(double &a, double v) { a += v; }
__device__ __forceinline__ void acc_add(float2 &a, float2 v) { a.x += v.x; a.y += v.y; }
__device__ __forceinline__ void acc_add(double2 &a, double2 v) { a.x += v.x; a.y += v.y; }

__device__ __forceinline__ float shfl_x(float v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double shfl_x(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ float2 shfl_x(float2 v, int o) { return make_float2(shfl_x(v.x, o), shfl_x(v.y, o)); }
__device__ __forceinline__ double2 shfl_x(double2 v, int o) { return make_double2(shfl_x(v.x, o), shfl_x(v.y, o)); }

__device__ __forceinline__ float mul(float a, float b) { return a * b; }
__device__ __forceinline__ double mul(double a, double b) { return a * b; }
__device__ __forceinline__ float2 mul(float2 a, float2 b) {
    return make_float2(fmaf(-a.y, b.y, a.x * b.x), fmaf(a.y, b.x, a.x * b.y));
}
__device__ __forceinline__ double2 mul(double2 a, double2 b) {
    return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.y, b.x, a.x * b.y));
}

// ---- unary_einsum ------------------------------------------------------------------------------------
// TPO threads cooperate on one output element. TPO = 1: one thread per output, consecutive threads walk y's
// memory order (coalesced when x's unit-stride mode is kept). TPO = 32: one warp per output, lanes stride over the
// fastest summed mode (coalesced when x's unit-stride mode is summed), shuffle reduction.
// gridDim.y > 1 slices the summed range (few outputs, long sums: sum of all elements, traces, column sums of a
// tall matrix): slice s writes part[s * total_c + output], unary_fold_kernel adds the slices in order - the result
// does not depend on scheduling (no atomics).
template <typename T, int TPO>
__global__ void __launch_bounds__(256, 3) unary_kernel(const __grid_constant__ UnaryParams p, int nsplit, const T *__restrict__ X,
                                                    T *__restrict__ Y, T *__restrict__ part) {
    const int lane = threadIdx.x % TPO;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / TPO;
    const int64_t k0 = p.nk > 0 ? p.k_ext[0] : 1;
    const int64_t sx0 = p.nk > 0 ? p.k_sx[0] : 0;
    const int64_t outer = p.nk > 0 ? p.total_k / k0 : 1;
    const bool cut_outer = outer >= nsplit;                 // slices cut the outer loop when it is long enough, else the inner one
    const int64_t per = ((cut_outer ? outer : k0) + nsplit - 1) / nsplit;
    // one group of TPO threads per (output, slice) pair; outputs fastest so neighbouring groups stay coalesced
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / TPO; g < p.total_c * nsplit; g += ngroups) {
        const int64_t slice = g / p.total_c, id = g - slice * p.total_c;
        int64_t ko_b = 0, ko_e = outer, j_b = 0, j_e = k0;
        if (nsplit > 1) {
            if (cut_outer) { ko_b = min(outer, per * slice); ko_e = min(outer, ko_b + per); }
            else { j_b = min(k0, per * slice); j_e = min(k0, j_b + per); }
        }
        int64_t r = id, ox = 0, oy = 0;
        for (int i = 0; i < p.nc; i++) {
            const int64_t e = p.c_ext[i], d = r % e;
            r /= e;
            ox += d * p.c_sx[i];
            oy += d * p.c_sy[i];
        }
        T acc0 = zero_of<T>(), acc1 = zero_of<T>(), acc2 = zero_of<T>(), acc3 = zero_of<T>();
        for (int64_t ko = ko_b; ko < ko_e; ko++) {
            int64_t q = ko, kx = ox;
            for (int i = 1; i < p.nk; i++) {
                const int64_t e = p.k_ext[i], d = q % e;
                q /= e;
                kx += d * p.k_sx[i];
            }
            const T *xp = X + kx;
            int64_t j = j_b + lane;
            // eight loads in flight per thread (ncu r01: 3 CTAs per SM x 4 loads left 49 KB in flight per SM, 60 % of HBM on the
            // 268 MB column sums); the summation tree is fixed, so results do not depend on scheduling
            for (; j + 7 * TPO < j_e; j += 8 * TPO) {
                const T v0 = xp[j * sx0], v1 = xp[(j + TPO) * sx0], v2 = xp[(j + 2 * TPO) * sx0], v3 = xp[(j + 3 * TPO) * sx0];
                const T v4 = xp[(j + 4 * TPO) * sx0], v5 = xp[(j + 5 * TPO) * sx0], v6 = xp[(j + 6 * TPO) * sx0], v7 = xp[(j + 7 * TPO) * sx0];
                acc_add(acc0, v0); acc_add(acc1, v1); acc_add(acc2, v2); acc_add(acc3, v3);
                acc_add(acc0, v4); acc_add(acc1, v5); acc_add(acc2, v6); acc_add(acc3, v7);
            }
            for (; j + 3 * TPO < j_e; j += 4 * TPO) {
                const T v0 = xp[j * sx0], v1 = xp[(j + TPO) * sx0], v2 = xp[(j + 2 * TPO) * sx0], v3 = xp[(j + 3 * TPO) * sx0];
                acc_add(acc0, v0); acc_add(acc1, v1); acc_add(acc2, v2); acc_add(acc3, v3);
            }
            for (; j < j_e; j += TPO) acc_add(acc0, xp[j * sx0]);
        }
        acc_add(acc0, acc1); acc_add(acc2, acc3); acc_add(acc0, acc2);
        if (TPO > 1) {
#pragma unroll
            for (int o = TPO / 2; o > 0; o >>= 1) acc_add(acc0, shfl_x(acc0, o));
        }
        if (lane == 0) {
            if (nsplit > 1) part[g] = acc0;
            else Y[oy] = acc0;
        }
    }
}

// part[slice * total_c + output] -> Y: one warp per output, lanes stride over the slices, fixed shuffle tree
// (deterministic; a sum of 4096 partials takes 128 steps instead of 4096).
template <typename T>
__global__ void __launch_bounds__(256) unary_fold_kernel(const __grid_constant__ UnaryParams p, const T *__restrict__ part,
                                                         int nsplit, T *__restrict__ Y) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; id < p.total_c; id += nwarps) {
        int64_t r = id, oy = 0;
        for (int i = 0; i < p.nc; i++) {
            const int64_t e = p.c_ext[i];
            oy += (r % e) * p.c_sy[i];
            r /= e;
        }
        T v = zero_of<T>();
        for (int s = lane; s < nsplit; s += 32) acc_add(v, part[(int64_t)s * p.total_c + id]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc_add(v, shfl_x(v, o));
        if (lane == 0) Y[oy] = v;
    }
}

// ---- hadamard ---------------------------------------------------------------------------------------
// One thread per 16 bytes of a / c (VEC elements along a's unit-stride mode). BMODE: 0 = b is broadcast along that
// mode (one b element per vector), 1 = b is contiguous along it (16-byte load of b too), 2 = anything else.
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <typename T, int VEC, int BMODE, typename IDX>
__global__ void __launch_bounds__(256) hadamard_kernel(const __grid_constant__ HadamardParams p, const T *__restrict__ A,
                                                       const T *__restrict__ B, T *__restrict__ C) {
    using P = Pack<T, VEC>;
    const IDX nvec = (IDX)(p.total / VEC);
    const IDX stride = (IDX)gridDim.x * blockDim.x;
    const IDX e0 = (IDX)(p.ext[0] / VEC);   // vectors along mode 0
    for (IDX id = (IDX)blockIdx.x * blockDim.x + threadIdx.x; id < nvec; id += stride) {
        IDX r = id / e0;
        int64_t ob = (int64_t)(id - r * e0) * VEC * p.sb[0];
        for (int i = 1; i < p.n; i++) {
            const IDX e = (IDX)p.ext[i], q = r / e;
            ob += (int64_t)(r - q * e) * p.sb[i];
            r = q;
        }
        const P a = reinterpret_cast<const P *>(A)[id];
        P c;
        if constexpr (BMODE == 0) {
            const T b = B[ob];
#pragma unroll
            for (int j = 0; j < VEC; j++) c.v[j] = mul(a.v[j], b);
        } else if constexpr (BMODE == 1) {
            const P b = *reinterpret_cast<const P *>(B + ob);
#pragma unroll
            for (int j = 0; j < VEC; j++) c.v[j] = mul(a.v[j], b.v[j]);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; j++) c.v[j] = mul(a.v[j], B[ob + j * p.sb[0]]);
        }
        reinterpret_cast<P *>(C)[id] = c;
    }
}

inline int grid_for(int64_t n, int threads, int cap = 148 * 16) {
    int64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

inline bool unary_warp_form(const UnaryParams &p) { return p.nk > 0 && p.k_sx[0] == 1 && p.k_ext[0] >= 16; }

template <typename T>
cudaError_t launch_unary_t(const UnaryParams &p, const T *X, T *Y, T *part, int nsplit, cudaStream_t s) {
    const bool warp = unary_warp_form(p);
    const int grid = grid_for(p.total_c * nsplit * (warp ? 32 : 1), 256);
    if (warp) unary_kernel<T, 32><<<grid, 256, 0, s>>>(p, nsplit, X, Y, part);
    else unary_kernel<T, 1><<<grid, 256, 0, s>>>(p, nsplit, X, Y, part);
    if (nsplit > 1) unary_fold_kernel<T><<<grid_for(p.total_c * 32, 256), 256, 0, s>>>(p, part, nsplit, Y);
    return cudaGetLastError();
}

template <typename T, int VEC>
cudaError_t launch_hadamard_t(const HadamardParams &p, const void *A, const void *B, void *C, bool vec_ok, cudaStream_t s) {
    const T *a = (const T *)A, *b = (const T *)B;
    T *c = (T *)C;
    const bool small = p.total < ((int64_t)1 << 31);
#define MB200_HAD(V, M)                                                                          \
    do {                                                                                          \
        const int g = grid_for(p.total / (V), 256);                                               \
        if (small) hadamard_kernel<T, V, M, unsigned><<<g, 256, 0, s>>>(p, a, b, c);              \
        else hadamard_kernel<T, V, M, int64_t><<<g, 256, 0, s>>>(p, a, b, c);                     \
    } while (0)
    if (vec_ok && VEC > 1) {
        if (p.sb[0] == 0) MB200_HAD(VEC, 0);
        else if (p.sb[0] == 1 && p.b_vec_aligned) MB200_HAD(VEC, 1);
        else MB200_HAD(VEC, 2);
    } else if (p.sb[0] == 0) {
        MB200_HAD(1, 0);
    } else {
        MB200_HAD(1, 2);
    }
#undef MB200_HAD
    return cudaGetLastError();
}

}  // namespace

int unary_nsplit(const UnaryParams &p) {
    if (p.nk == 0 || p.total_k < 4096) return 1;
    const int tpo = unary_warp_form(p) ? 32 : 1;
    const int64_t fill = 148 * 2048;                                  // resident threads of the whole GPU
    if (p.total_c * tpo >= fill / 2) return 1;                        // enough threads already: skip the fold pass
    const int64_t want = (fill + p.total_c * tpo - 1) / (p.total_c * tpo);
    const int64_t k0 = p.k_ext[0], outer = p.total_k / k0;
    const int64_t cap = std::max<int64_t>(outer, k0 / (16 * tpo));    // slices cut the outer loop, or the inner one
    return (int)std::max<int64_t>(1, std::min<int64_t>(std::min(want, cap), 4096));
}

cudaError_t launch_unary(int dtype, const UnaryParams &p, const void *X, void *Y, void *part, int nsplit, cudaStream_t s) {
    if (p.total_c <= 0) return cudaSuccess;
    switch (dtype) {
        case MB200_F32: return launch_unary_t<float>(p, (const float *)X, (float *)Y, (float *)part, nsplit, s);
        case MB200_F64: return launch_unary_t<double>(p, (const double *)X, (double *)Y, (double *)part, nsplit, s);
        case MB200_C64: return launch_unary_t<float2>(p, (const float2 *)X, (float2 *)Y, (float2 *)part, nsplit, s);
        default: return launch_unary_t<double2>(p, (const double2 *)X, (double2 *)Y, (double2 *)part, nsplit, s);
    }
}

cudaError_t launch_hadamard(int dtype, const HadamardParams &p, const void *A, const void *B, void *C, cudaStream_t s) {
    if (p.total <= 0) return cudaSuccess;
    const int vec = (int)(16 / dtype_size(dtype));
    const bool vec_ok = p.n > 0 && p.ext[0] % vec == 0 && ((((uintptr_t)A) | ((uintptr_t)C)) & 15) == 0;
    switch (dtype) {
        case MB200_F32: return launch_hadamard_t<float, 4>(p, A, B, C, vec_ok, s);
        case MB200_F64: return launch_hadamard_t<double, 2>(p, A, B, C, vec_ok, s);
        case MB200_C64: return launch_hadamard_t<float2, 2>(p, A, B, C, vec_ok, s);
        default: return launch_hadamard_t<double2, 1>(p, A, B, C, vec_ok, s);
    }
}

}  // namespace mb200
