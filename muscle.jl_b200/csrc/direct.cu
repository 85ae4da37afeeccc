// K5 — single-kernel direct contraction, plus the two tiny helper kernels (offset tables, dtype
// promotion). sm_100a.
//
// The direct kernel is the launch-bound fast path: one launch, no workspace, no offset tables.
// Every contraction in the reference's own test-suite (test/unit/operations/binary_einsum.jl) is
// this size. It is also the general-shape path for outer products / scaling (K <= 2), where the
// op is a pure streaming kernel, and for empty or zero-extent cases.
//
// One thread per output element (grid-stride): decode the element's digits over C's walk modes
// (left, right, batch) to get base offsets in A, B and C, then run the summed modes with the
// fastest one as a plain strided inner loop.
#include <algorithm>

#include "kernels.cuh"

namespace mb200 {

namespace {

__device__ __forceinline__ void cfma(float &acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ __forceinline__ void cfma(double &acc, double a, double b) { acc = fma(a, b, acc); }
__device__ __forceinline__ void cfma(float2 &acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfma(double2 &acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}
template <typename T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ float zero_of<float>() { return 0.f; }
template <> __device__ __forceinline__ double zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ float2 zero_of<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ double2 zero_of<double2>() { return make_double2(0.0, 0.0); }

template <typename T>
__global__ void __launch_bounds__(256) direct_kernel(const __grid_constant__ DirectParams p,
                                                     const T *__restrict__ A, const T *__restrict__ B,
                                                     T *__restrict__ C) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t k0 = p.nk > 0 ? p.k_ext[0] : 1;
    const int64_t sa0 = p.nk > 0 ? p.k_sa[0] : 0;
    const int64_t sb0 = p.nk > 0 ? p.k_sb[0] : 0;
    const int64_t outer = p.nk > 0 ? (k0 > 0 ? p.total_k / k0 : 0) : 1;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < p.total_c; id += stride) {
        int64_t r = id, oa = 0, ob = 0, oc = 0;
        for (int i = 0; i < p.nc; i++) {
            int64_t e = p.c_ext[i];
            int64_t d = r % e;
            r /= e;
            oa += d * p.c_sa[i];
            ob += d * p.c_sb[i];
            oc += d * p.c_sc[i];
        }
        T acc = zero_of<T>();
        if (p.total_k > 0) {
            for (int64_t ko = 0; ko < outer; ko++) {
                int64_t q = ko, ka = oa, kb = ob;
                for (int i = 1; i < p.nk; i++) {
                    int64_t e = p.k_ext[i];
                    int64_t d = q % e;
                    q /= e;
                    ka += d * p.k_sa[i];
                    kb += d * p.k_sb[i];
                }
                for (int64_t j = 0; j < k0; j++) cfma(acc, A[ka + j * sa0], B[kb + j * sb0]);
            }
        }
        C[oc] = acc;
    }
}

// Few outputs, long sums (inner products, norms <psi|psi>, isometry checks): one thread per output would serialise
// the whole sum. Here blockIdx.y cuts the summed range into `nsplit` slices; each CTA reduces its slice of one
// output with a strided loop + warp shuffles and adds the partial with atomics (C is zeroed first by zero_kernel).
template <typename T> struct Scalar;
template <> struct Scalar<float> { using type = float; static constexpr int N = 1; };
template <> struct Scalar<double> { using type = double; static constexpr int N = 1; };
template <> struct Scalar<float2> { using type = float; static constexpr int N = 2; };
template <> struct Scalar<double2> { using type = double; static constexpr int N = 2; };
__device__ __forceinline__ float comp(const float &v, int) { return v; }
__device__ __forceinline__ double comp(const double &v, int) { return v; }
__device__ __forceinline__ float comp(const float2 &v, int i) { return i ? v.y : v.x; }
__device__ __forceinline__ double comp(const double2 &v, int i) { return i ? v.y : v.x; }

template <typename T>
__global__ void __launch_bounds__(256) zero_kernel(const __grid_constant__ DirectParams p, T *__restrict__ C) {
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= p.total_c) return;
    int64_t r = id, oc = 0;
    for (int i = 0; i < p.nc; i++) { int64_t e = p.c_ext[i]; oc += (r % e) * p.c_sc[i]; r /= e; }
    C[oc] = zero_of<T>();
}

template <typename T>
__global__ void __launch_bounds__(256) direct_splitk_kernel(const __grid_constant__ DirectParams p, const T *__restrict__ A,
                                                            const T *__restrict__ B, T *__restrict__ C) {
    using S = typename Scalar<T>::type;
    __shared__ S red[8][2];
    const int64_t id = blockIdx.x;   // one output element per blockIdx.x
    int64_t r = id, oa = 0, ob = 0, oc = 0;
    for (int i = 0; i < p.nc; i++) {
        int64_t e = p.c_ext[i], d = r % e;
        r /= e;
        oa += d * p.c_sa[i]; ob += d * p.c_sb[i]; oc += d * p.c_sc[i];
    }
    const int64_t per = (p.total_k + gridDim.y - 1) / gridDim.y;
    const int64_t k_begin = (int64_t)blockIdx.y * per, k_end = min(p.total_k, k_begin + per);
    T acc = zero_of<T>();
    for (int64_t k = k_begin + threadIdx.x; k < k_end; k += blockDim.x) {
        int64_t q = k, ka = oa, kb = ob;
        for (int i = 0; i < p.nk; i++) {
            int64_t e = p.k_ext[i], d = q % e;
            q /= e;
            ka += d * p.k_sa[i]; kb += d * p.k_sb[i];
        }
        cfma(acc, A[ka], B[kb]);
    }
    S part[2] = {comp(acc, 0), Scalar<T>::N == 2 ? comp(acc, 1) : S(0)};
#pragma unroll
    for (int c = 0; c < Scalar<T>::N; c++) {
        S v = part[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < Scalar<T>::N) {
        S v = 0;
        for (int w = 0; w < 8; w++) v += red[w][threadIdx.x];
        atomicAdd(reinterpret_cast<S *>(C + oc) + threadIdx.x, v);
    }
}

// "apply": one thread per row of the big operand. The operator (J x K <= 8 x 8) sits in shared memory, the row's K
// inputs are loaded once, all J outputs are produced from registers; offsets come from a digit decode over the
// MERGED big-free modes (2-4 modes for a gate on an n-qubit state), so there are no tables and every byte of the
// big tensor is read once and written once: a streaming kernel.
template <typename T, int KP, int JP>
__global__ void __launch_bounds__(256) apply_kernel(const __grid_constant__ ApplyParams p, const T *__restrict__ X,
                                                    const T *__restrict__ S, T *__restrict__ C) {
    // KP / JP: K and J rounded up to 2, 4 or 8 at compile time (operator zero-padded) so the inner loops unroll
    // to exactly the work needed; rows are decoded with 32-bit divisions (total_big < 2^31 is checked on the host)
    __shared__ T sS[JP][KP];
    if (threadIdx.x < JP * KP) {
        const int j = threadIdx.x / KP, k = threadIdx.x % KP;
        sS[j][k] = (j < p.J && k < p.K) ? S[p.js[j] + p.ks[k]] : zero_of<T>();
    }
    __syncthreads();
    const unsigned total = (unsigned)p.total_big, stride = gridDim.x * blockDim.x;
    for (unsigned r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += stride) {
        unsigned q = r;
        int64_t ox = 0, oc = 0;
        for (int i = 0; i < p.nbig; i++) {
            const unsigned e = (unsigned)p.big_ext[i], qq = q / e, d = q - qq * e;
            q = qq;
            ox += (int64_t)d * p.big_sx[i];
            oc += (int64_t)d * p.big_sc[i];
        }
        T x[KP];
#pragma unroll
        for (int k = 0; k < KP; k++) x[k] = (k < p.K) ? X[ox + p.kx[k]] : zero_of<T>();
#pragma unroll
        for (int j = 0; j < JP; j++) {
            T acc = zero_of<T>();
#pragma unroll
            for (int k = 0; k < KP; k++) cfma(acc, x[k], sS[j][k]);
            if (j < p.J) C[oc + p.jc[j]] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) table_kernel(int64_t *__restrict__ out, int64_t size,
                                                    const __grid_constant__ TableSpec spec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < size; x += stride) {
        int64_t r = x, off = 0;
        for (int i = 0; i < spec.n; i++) {
            int64_t e = spec.ext[i];
            off += (r % e) * spec.stride[i];
            r /= e;
        }
        out[x] = off;
    }
}

template <typename TD, typename TS> __device__ __forceinline__ TD conv(TS v);
template <> __device__ __forceinline__ double conv<double, float>(float v) { return (double)v; }
template <> __device__ __forceinline__ float2 conv<float2, float>(float v) { return make_float2(v, 0.f); }
template <> __device__ __forceinline__ double2 conv<double2, float>(float v) { return make_double2((double)v, 0.0); }
template <> __device__ __forceinline__ double2 conv<double2, double>(double v) { return make_double2(v, 0.0); }
template <> __device__ __forceinline__ double2 conv<double2, float2>(float2 v) { return make_double2((double)v.x, (double)v.y); }

template <typename TD, typename TS>
__global__ void __launch_bounds__(256) convert_kernel(TD *__restrict__ dst, const TS *__restrict__ src,
                                                      int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = conv<TD, TS>(src[i]);
}

// out[i] = sum_s in[s * n + i] over plain reals (a sum does not care about re/im interleaving)
template <typename T>
__global__ void __launch_bounds__(256) reduce_slots_kernel(T *__restrict__ out, const T *__restrict__ in, int64_t n, int nslots) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T acc = in[i];
        for (int s = 1; s < nslots; s++) {
            T v = in[(int64_t)s * n + i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
    }
}

// the same on scalars, for slabs whose byte count is not a multiple of 32 (small blocks of the blocked-tensor bridge)
template <typename T>
__global__ void __launch_bounds__(256) reduce_slots_scalar_kernel(T *__restrict__ out, const T *__restrict__ in, int64_t n, int nslots) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T acc = in[i];
        for (int s = 1; s < nslots; s++) acc += in[(int64_t)s * n + i];
        out[i] = acc;
    }
}

// Cross-rank barrier of the fused reduce-scatter, without a collective: after its contraction every rank stores the
// call's epoch into ITS entry of every rank's flag array (peer mappings, system scope); the owner's slot-sum kernel
// spins until all nslots entries carry the epoch. The GEMM epilogue ended with __threadfence_system() and the kernel
// boundary orders this store after it, so a rank that acquires the flag also sees the slot data.
__global__ void signal_peers_kernel(ScatterDesc flags, int epoch) {
    const int r = threadIdx.x;
    if (r < flags.nranks) {
        __threadfence_system();
        int *f = reinterpret_cast<int *>(flags.peer[r]) + flags.rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
}

template <typename T>
__global__ void __launch_bounds__(256) reduce_slots_wait_kernel(T *__restrict__ out, const T *in, int64_t n, int nslots,
                                                                const int *flags, int epoch) {
    if (threadIdx.x < nslots) {
        const long long t0 = clock64();
        int f;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(f) : "l"(flags + threadIdx.x) : "memory");
            if (clock64() - t0 > 20000000000LL) __trap();   // a rank that never arrives must not hang the GPU
        } while (f - epoch < 0);
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T acc = in[i];   // first touched after the acquire above: never a stale L1 line
        for (int s = 1; s < nslots; s++) {
            T v = in[(int64_t)s * n + i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
    }
}

inline int grid_for(int64_t n, int threads, int cap = 148 * 16) {
    int64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

}  // namespace

cudaError_t launch_build_table(int64_t *out, int64_t size, const TableSpec &spec, cudaStream_t s) {
    table_kernel<<<grid_for(size, 256), 256, 0, s>>>(out, size, spec);
    return cudaGetLastError();
}

cudaError_t launch_direct(int dtype, const DirectParams &p, const void *A, const void *B, void *C,
                          cudaStream_t s) {
    if (p.total_c <= 0) return cudaSuccess;
    // few outputs, long sums: split the summed range over CTAs (two launches: zero C, then reduce + atomics)
    if (p.total_c <= 2048 && p.total_k >= 8192) {
        int64_t want = (148 * 8 + p.total_c - 1) / p.total_c;               // ~8 CTAs per SM in total
        int64_t cap = p.total_k / 2048;                                      // at least 2048 terms per CTA
        int nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(want, cap), 65535));
        dim3 grid((unsigned)p.total_c, (unsigned)nsplit);
        int gz = grid_for(p.total_c, 256);
#define MB200_SPLITK(T)                                                                       \
        zero_kernel<T><<<gz, 256, 0, s>>>(p, (T *)C);                                         \
        direct_splitk_kernel<T><<<grid, 256, 0, s>>>(p, (const T *)A, (const T *)B, (T *)C)
        switch (dtype) {
            case MB200_F32: MB200_SPLITK(float); break;
            case MB200_F64: MB200_SPLITK(double); break;
            case MB200_C64: MB200_SPLITK(float2); break;
            default: MB200_SPLITK(double2); break;
        }
#undef MB200_SPLITK
        return cudaGetLastError();
    }
    int g = grid_for(p.total_c, 256);
    switch (dtype) {
        case MB200_F32: direct_kernel<float><<<g, 256, 0, s>>>(p, (const float *)A, (const float *)B, (float *)C); break;
        case MB200_F64: direct_kernel<double><<<g, 256, 0, s>>>(p, (const double *)A, (const double *)B, (double *)C); break;
        case MB200_C64: direct_kernel<float2><<<g, 256, 0, s>>>(p, (const float2 *)A, (const float2 *)B, (float2 *)C); break;
        default: direct_kernel<double2><<<g, 256, 0, s>>>(p, (const double2 *)A, (const double2 *)B, (double2 *)C); break;
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_apply_t(const ApplyParams &p, const void *X, const void *S, void *C, cudaStream_t s) {
    const int g = grid_for(p.total_big, 256, 148 * 8);
    const int kp = p.K <= 2 ? 2 : (p.K <= 4 ? 4 : 8), jp = p.J <= 2 ? 2 : (p.J <= 4 ? 4 : 8);
#define MB200_APPLY(KP, JP) apply_kernel<T, KP, JP><<<g, 256, 0, s>>>(p, (const T *)X, (const T *)S, (T *)C)
    if (kp == 2 && jp == 2) MB200_APPLY(2, 2);
    else if (kp == 2 && jp == 4) MB200_APPLY(2, 4);
    else if (kp == 2) MB200_APPLY(2, 8);
    else if (kp == 4 && jp == 2) MB200_APPLY(4, 2);
    else if (kp == 4 && jp == 4) MB200_APPLY(4, 4);
    else if (kp == 4) MB200_APPLY(4, 8);
    else if (jp == 2) MB200_APPLY(8, 2);
    else if (jp == 4) MB200_APPLY(8, 4);
    else MB200_APPLY(8, 8);
#undef MB200_APPLY
    return cudaGetLastError();
}

cudaError_t launch_apply(int dtype, const ApplyParams &p, const void *X, const void *S, void *C, cudaStream_t s) {
    if (p.total_big <= 0) return cudaSuccess;
    if (p.total_big >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    switch (dtype) {
        case MB200_F32: return launch_apply_t<float>(p, X, S, C, s);
        case MB200_F64: return launch_apply_t<double>(p, X, S, C, s);
        case MB200_C64: return launch_apply_t<float2>(p, X, S, C, s);
        default: return launch_apply_t<double2>(p, X, S, C, s);
    }
}

cudaError_t launch_reduce_slots(int dtype, void *out, const void *staging, int64_t slab_elems, int nslots, cudaStream_t s) {
    const int64_t bytes = slab_elems * (int64_t)dtype_size(dtype);
    if (bytes <= 0) return cudaSuccess;
    if (bytes % 32 != 0 || (((uintptr_t)out | (uintptr_t)staging) & 31)) {   // odd slab size / unaligned base: scalar form
        if (dtype_is_double(dtype)) {
            const int64_t n = bytes / 8;
            reduce_slots_scalar_kernel<double><<<grid_for(n, 256), 256, 0, s>>>((double *)out, (const double *)staging, n, nslots);
        } else {
            const int64_t n = bytes / 4;
            reduce_slots_scalar_kernel<float><<<grid_for(n, 256), 256, 0, s>>>((float *)out, (const float *)staging, n, nslots);
        }
        return cudaGetLastError();
    }
    if (dtype_is_double(dtype)) {
        const int64_t n = bytes / 32;   // double4
        reduce_slots_kernel<double4><<<grid_for(n, 256), 256, 0, s>>>((double4 *)out, (const double4 *)staging, n, nslots);
    } else {
        const int64_t n = bytes / 16;   // float4
        reduce_slots_kernel<float4><<<grid_for(n, 256), 256, 0, s>>>((float4 *)out, (const float4 *)staging, n, nslots);
    }
    return cudaGetLastError();
}

cudaError_t launch_signal_peers(const ScatterDesc &flags, int epoch, cudaStream_t s) {
    signal_peers_kernel<<<1, 32, 0, s>>>(flags, epoch);
    return cudaGetLastError();
}

cudaError_t launch_reduce_slots_wait(int dtype, void *out, const void *staging, int64_t slab_elems, int nslots, const int *flags,
                                     int epoch, cudaStream_t s) {
    const int64_t bytes = slab_elems * (int64_t)dtype_size(dtype);
    if (bytes <= 0) return cudaSuccess;
    if (bytes % 32 != 0 || nslots > 32) return cudaErrorInvalidValue;
    if (dtype_is_double(dtype)) {
        const int64_t n = bytes / 32;   // double4
        reduce_slots_wait_kernel<double4><<<grid_for(n, 256, 148 * 8), 256, 0, s>>>((double4 *)out, (const double4 *)staging, n, nslots, flags, epoch);
    } else {
        const int64_t n = bytes / 16;   // float4
        reduce_slots_wait_kernel<float4><<<grid_for(n, 256, 148 * 8), 256, 0, s>>>((float4 *)out, (const float4 *)staging, n, nslots, flags, epoch);
    }
    return cudaGetLastError();
}

cudaError_t launch_convert(int dd, void *dst, int ds, const void *src, int64_t n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    int g = grid_for(n, 256);
#define MB200_CONV(TD, TS) convert_kernel<TD, TS><<<g, 256, 0, s>>>((TD *)dst, (const TS *)src, n)
    if (dd == MB200_F64 && ds == MB200_F32) MB200_CONV(double, float);
    else if (dd == MB200_C64 && ds == MB200_F32) MB200_CONV(float2, float);
    else if (dd == MB200_C128 && ds == MB200_F32) MB200_CONV(double2, float);
    else if (dd == MB200_C128 && ds == MB200_F64) MB200_CONV(double2, double);
    else if (dd == MB200_C128 && ds == MB200_C64) MB200_CONV(double2, float2);
    else return cudaErrorInvalidValue;
#undef MB200_CONV
    return cudaGetLastError();
}

}  // namespace mb200
