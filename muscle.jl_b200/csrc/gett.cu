// K2 — gather-GEMM ("GETT": TTGT with the transposes fused into the GEMM's loads and stores).
//
//   C[rowC[m] + colC[n] + batC[l]] = sum_k A[rowA[m] + kA[k] + batA[l]] * B[colB[n] + kB[k] + batB[l]]
//
// The reference's BackendBase does TTGT as four passes (src/Operations/binary_einsum.jl:89-95:
// permutedims(A), permutedims(B), gemm, permutedims(C)). Here the three permutes never touch HBM:
// operand tiles are gathered straight from the tensors' native layouts with element-granular
// cp.async (16 B per ComplexF64 element = half a 32 B sector, so a gather costs at most 2x sector
// over-fetch from L2 and nothing extra from HBM), and the epilogue scatters C directly in the
// requested index order. Offsets come from small per-plan int64 tables (rows / columns / k / batch).
//
// Compute cores
//   CoreZ : ComplexF64, FP64 tensor cores  — mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), complex product
//           as the 4M real decomposition on interleaved (re,im) fragments: one LDS.128 yields both the
//           real and the imaginary A (or B) fragment; -Im(A) is a sign-bit flip on the ALU pipe.
//   CoreD : Float64, DMMA.
//   CoreC / CoreS : ComplexF32 / Float32 on FFMA (used for shapes the tcgen05 path does not take).
//
// Pipeline: STAGES-deep cp.async ring over k-blocks of BK, one __syncthreads per k-block.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "kernels.cuh"

namespace mb200 {

namespace {

// ------------------------------------------------------------------------------------------------
// zero-fill form: copies BYTES when valid, writes BYTES of zeros otherwise (no branch)
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void *smem, const void *gmem, bool valid) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    int src = valid ? BYTES : 0;
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(src));
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem), "r"(src));
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gmem), "r"(src));
}
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem));
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double neg_bits(double x) {  // sign flip on the integer pipe
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}

// ------------------------------------------------------------------------------------------------
// ComplexF64 on DMMA (4M). Warp tile WM x WN complex, m8n8k4 fragments.
template <int BM_, int BN_, int WM_, int WN_, int BK_, int STAGES_>
struct CoreZ {
    using Elem = double2;
    static constexpr int KSTEP = 1;   // k values per smem element
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = BK_, STAGES = STAGES_;
    static constexpr int NWARPS = (BM / WM) * (BN / WN), NTHREADS = 32 * NWARPS;
    static constexpr int LDA = BM + 2, LDB = BN + 2;  // == 2 (mod 8) in 16 B units: conflict-free LDS.128
    static constexpr int MT = WM / 8, NT = WN / 8;
    struct Acc { double re[MT][NT][2], im[MT][NT][2]; };
    __device__ static void init(Acc &a) {
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) a.re[i][j][0] = a.re[i][j][1] = a.im[i][j][0] = a.im[i][j][1] = 0.0;
    }
    static constexpr int SLOTS = (BK / 4) * 4;  // hook calls per k-block
    template <class Hook>
    __device__ static void compute(Acc &acc, const Elem *sa, const Elem *sb, int warp, int lane, Hook &&hook) {
        const int wm = (warp % (BM / WM)) * WM, wn = (warp / (BM / WM)) * WN;
        const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double2 af[MT], bf[NT];
            double nai[MT];
#pragma unroll
            for (int i = 0; i < MT; i++) {
                af[i] = sa[(kk * 4 + fk) * LDA + wm + i * 8 + fr];
                nai[i] = neg_bits(af[i].y);
            }
#pragma unroll
            for (int j = 0; j < NT; j++) bf[j] = sb[(kk * 4 + fk) * LDB + wn + j * 8 + fr];
            // four passes of MT*NT independent DMMAs: consecutive DMMAs never share an accumulator
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(acc.re[i][j][0], acc.re[i][j][1], af[i].x, bf[j].x);
            hook(kk * 4 + 0);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(acc.im[i][j][0], acc.im[i][j][1], af[i].x, bf[j].y);
            hook(kk * 4 + 1);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(acc.re[i][j][0], acc.re[i][j][1], nai[i], bf[j].y);
            hook(kk * 4 + 2);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(acc.im[i][j][0], acc.im[i][j][1], af[i].y, bf[j].x);
            hook(kk * 4 + 3);
        }
    }
    __device__ static void store(const Acc &acc, Elem *C, const int64_t *sRowC, const int64_t *sColC,
                                 int64_t cb, int mrem, int nrem, int warp, int lane, const ScatterDesc &sc) {
        const int wm = (warp % (BM / WM)) * WM, wn = (warp / (BM / WM)) * WN;
        const int fr = lane >> 2, fc = (lane & 3) * 2;
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int c = wn + j * 8 + fc + h;
                if (c >= nrem) continue;
                int64_t co = sColC[c] + cb;
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    int r = wm + i * 8 + fr;
                    if (r < mrem) *scatter_ptr(sc, C, sRowC[r] + co) = make_double2(acc.re[i][j][h], acc.im[i][j][h]);
                }
            }
        if (sc.nranks) __threadfence_system();   // peer stores must be visible before the cross-rank barrier
    }
};

// Float64 on DMMA.
template <int BM_, int BN_, int WM_, int WN_, int BK_, int STAGES_>
struct CoreD {
    using Elem = double;
    static constexpr int KSTEP = 1;   // k values per smem element
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = BK_, STAGES = STAGES_;
    static constexpr int NWARPS = (BM / WM) * (BN / WN), NTHREADS = 32 * NWARPS;
    static constexpr int LDA = BM + 4, LDB = BN + 4;  // == 4 (mod 16) in 8 B units: conflict-free LDS.64
    static constexpr int MT = WM / 8, NT = WN / 8;
    struct Acc { double c[MT][NT][2]; };
    __device__ static void init(Acc &a) {
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) a.c[i][j][0] = a.c[i][j][1] = 0.0;
    }
    static constexpr int SLOTS = (BK / 4) * MT;
    template <class Hook>
    __device__ static void compute(Acc &acc, const Elem *sa, const Elem *sb, int warp, int lane, Hook &&hook) {
        const int wm = (warp % (BM / WM)) * WM, wn = (warp / (BM / WM)) * WN;
        const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; i++) af[i] = sa[(kk * 4 + fk) * LDA + wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < NT; j++) bf[j] = sb[(kk * 4 + fk) * LDB + wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < MT; i++) {
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(acc.c[i][j][0], acc.c[i][j][1], af[i], bf[j]);
                hook(kk * MT + i);
            }
        }
    }
    __device__ static void store(const Acc &acc, Elem *C, const int64_t *sRowC, const int64_t *sColC,
                                 int64_t cb, int mrem, int nrem, int warp, int lane, const ScatterDesc &) {
        const int wm = (warp % (BM / WM)) * WM, wn = (warp / (BM / WM)) * WN;
        const int fr = lane >> 2, fc = (lane & 3) * 2;
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int c = wn + j * 8 + fc + h;
                if (c >= nrem) continue;
                int64_t co = sColC[c] + cb;
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    int r = wm + i * 8 + fr;
                    if (r < mrem) C[sRowC[r] + co] = acc.c[i][j][h];
                }
            }
    }
};

// Float64 on DMMA for operands that are both K-major with the unit-stride summed mode of even extent: an smem element is a PAIR
// of consecutive k (16 bytes, like a ComplexF64 element), so one cp.async moves two k and one LDS.128 feeds two DMMAs
// (x with x, y with y: CoreZ without the imaginary passes). Half the gather instructions and fragment loads per flop of CoreD
// and twice the DMMAs per CTA barrier. BK counts pairs; GettParams::K and the k-offset tables are walked in pairs (KSTEP = 2).
template <int BM_, int BN_, int WM_, int WN_, int BK_, int STAGES_>
struct CoreD2 {
    using Elem = double2;
    static constexpr int KSTEP = 2;
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, BK = BK_, STAGES = STAGES_;
    static constexpr int NWARPS = (BM / WM) * (BN / WN), NTHREADS = 32 * NWARPS;
    static constexpr int LDA = BM + 2, LDB = BN + 2;  // == 2 (mod 8) in 16 B units: conflict-free LDS.128
    static constexpr int MT = WM / 8, NT = WN / 8;
    struct Acc { double c[MT][NT][2]; };
    __device__ static void init(Acc &a) {
#pragma unroll
        for (int i = 0; i < MT; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) a.c[i][j][0] = a.c[i][j][1] = 0.0;
    }
    static constexpr int SLOTS = (BK / 4) * 2;
    template <class Hook>
    __device__ static void compute(Acc &acc, const Elem *sa, const Elem *sb, int warp, int lane, Hook &&hook) {
        const int wm = (warp % (BM / WM)) * WM, wn = (warp / (BM / WM)) * WN;
        const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
        for (int kk = 0; kk < BK / 4; kk++) {
            double2 af[MT], bf[NT];
#pragma unroll
            for (int i = 0; i < MT; i++) af[i] = sa[(kk * 4 + fk) * LDA + wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < NT; j++) bf[j] = sb[(kk * 4 + fk) * LDB + wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(acc.c[i][j][0], acc.c[i][j][1], af[i].x, bf[j].x);   // even k of the pairs
            hook(kk * 2 + 0);
#pragma unroll
            for (int i = 0; i < MT; i++)
#pragma unroll
                for (int j = 0; j < NT; j++) dmma(acc.c[i][j][0], acc.c[i][j][1], af[i].y, bf[j].y);   // odd k
            hook(kk * 2 + 1);
        }
    }
    __device__ static void store(const Acc &acc, Elem *Cv, const int64_t *sRowC, const int64_t *sColC,
                                 int64_t cb, int mrem, int nrem, int warp, int lane, const ScatterDesc &) {
        double *C = reinterpret_cast<double *>(Cv);   // C offsets are in doubles
        const int wm = (warp % (BM / WM)) * WM, wn = (warp / (BM / WM)) * WN;
        const int fr = lane >> 2, fc = (lane & 3) * 2;
#pragma unroll
        for (int j = 0; j < NT; j++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int c = wn + j * 8 + fc + h;
                if (c >= nrem) continue;
                int64_t co = sColC[c] + cb;
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    int r = wm + i * 8 + fr;
                    if (r < mrem) C[sRowC[r] + co] = acc.c[i][j][h];
                }
            }
    }
};

// ComplexF32 / Float32 on FFMA. 256 threads as 16 x 16; thread (tx, ty) owns rows tx + 16 i and
// columns ty + 16 j, so a warp's smem reads are 16 consecutive elements (A) or 2 broadcasts (B).
template <typename E, int BM_, int BN_, int BK_, int STAGES_>
struct CoreF {
    using Elem = E;
    static constexpr int KSTEP = 1;
    static constexpr int BM = BM_, BN = BN_, BK = BK_, STAGES = STAGES_;
    static constexpr int NTHREADS = 256;
    static constexpr int LDA = BM, LDB = BN;
    static constexpr int TM = BM / 16, TN = BN / 16;
    struct Acc { E c[TM][TN]; };
    __device__ static void zero(float &v) { v = 0.f; }
    __device__ static void zero(float2 &v) { v = make_float2(0.f, 0.f); }
    __device__ static void mac(float &c, float a, float b) { c = fmaf(a, b, c); }
    __device__ static void mac(float2 &c, float2 a, float2 b) {
        c.x = fmaf(a.x, b.x, c.x);
        c.y = fmaf(a.x, b.y, c.y);
        c.x = fmaf(-a.y, b.y, c.x);
        c.y = fmaf(a.y, b.x, c.y);
    }
    __device__ static void init(Acc &a) {
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
            for (int j = 0; j < TN; j++) zero(a.c[i][j]);
    }
    static constexpr int SLOTS = BK;
    template <class Hook>
    __device__ static void compute(Acc &acc, const Elem *sa, const Elem *sb, int warp, int lane, Hook &&hook) {
        const int tid = warp * 32 + lane, tx = tid & 15, ty = tid >> 4;
#pragma unroll
        for (int k = 0; k < BK; k++) {
            E af[TM], bf[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) af[i] = sa[k * LDA + tx + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; j++) bf[j] = sb[k * LDB + ty + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) mac(acc.c[i][j], af[i], bf[j]);
            hook(k);
        }
    }
    __device__ static void store(const Acc &acc, Elem *C, const int64_t *sRowC, const int64_t *sColC,
                                 int64_t cb, int mrem, int nrem, int warp, int lane, const ScatterDesc &) {
        const int tid = warp * 32 + lane, tx = tid & 15, ty = tid >> 4;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int c = ty + 16 * j;
            if (c >= nrem) continue;
            int64_t co = sColC[c] + cb;
#pragma unroll
            for (int i = 0; i < TM; i++) {
                int r = tx + 16 * i;
                if (r < mrem) C[sRowC[r] + co] = acc.c[i][j];
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
template <class Core>
constexpr size_t gett_smem_bytes() {
    return (size_t)Core::STAGES * Core::BK * (Core::LDA + Core::LDB) * sizeof(typename Core::Elem) +
           (size_t)(Core::BM + Core::BN) * 2 * sizeof(int64_t) + (size_t)4 * Core::BK * sizeof(int64_t);
}

template <class Core>
constexpr size_t gett_persistent_smem_bytes() {   // two sets of offset tables
    return gett_smem_bytes<Core>() + (size_t)(Core::BM + Core::BN) * 2 * sizeof(int64_t);
}

// Tile order: groups of GROUP_M row tiles, walked column by column inside a group, so the CTAs of a
// wave share a few A row-panels (kept in the 126 MB L2) while B column-panels stream through.
constexpr int GROUP_M = 8;

// Split-K (gridDim.y > 1): a contraction with too few tiles to fill the 148 SMs runs every tile's k-range in
// gridDim.y slices; slice s of tile t writes its partial tile to ws[(s * ntiles + t) * BM * BN ...] (tile-linear,
// coalesced) and splitk_reduce_kernel adds the slices in order and scatters to C through the tables (deterministic).
template <class Core>
__global__ void __launch_bounds__(Core::NTHREADS, Core::NTHREADS <= 128 ? 4 : 1) gett_kernel(const __grid_constant__ GettParams p,
                                                                                            typename Core::Elem *ws, int64_t tile0) {
    using E = typename Core::Elem;
    constexpr int BM = Core::BM, BN = Core::BN, BK = Core::BK, S = Core::STAGES;
    constexpr int LDA = Core::LDA, LDB = Core::LDB, NT = Core::NTHREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *sA = reinterpret_cast<E *>(smem_raw);
    E *sB = sA + (size_t)S * BK * LDA;
    int64_t *sRowA = reinterpret_cast<int64_t *>(sB + (size_t)S * BK * LDB);
    int64_t *sColB = sRowA + BM;
    int64_t *sRowC = sColB + BN;
    int64_t *sColC = sRowC + BM;
    int64_t *sK = sColC + BN;   // [2][2][BK]: k-offsets of A and B for the k-block whose gather is issued next

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
    const int64_t tiles = tiles_m * tiles_n;
    const int64_t bid = tile0 + blockIdx.x;   // tile0 > 0: the split tail launch of a ragged last wave
    const int64_t l = bid / tiles;
    const int64_t t = bid % tiles;
    // grouped rasterisation
    const int64_t per_group = GROUP_M * tiles_n;
    const int64_t g = t / per_group;
    const int64_t gm0 = g * GROUP_M;
    const int64_t gsz = (tiles_m - gm0) < GROUP_M ? (tiles_m - gm0) : GROUP_M;
    const int64_t tm = gm0 + (t % per_group) % gsz;
    const int64_t tn = (t % per_group) / gsz;
    const int64_t m0 = tm * BM, n0 = tn * BN;
    const int mrem = (int)((p.M - m0) < BM ? (p.M - m0) : BM);
    const int nrem = (int)((p.N - n0) < BN ? (p.N - n0) : BN);

    const bool split = gridDim.y > 1;
    for (int i = tid; i < BM; i += NT) {
        bool v = i < mrem;
        sRowA[i] = v ? p.rowA[m0 + i] : 0;
        sRowC[i] = split ? i : (v ? p.rowC[m0 + i] : 0);
    }
    for (int i = tid; i < BN; i += NT) {
        bool v = i < nrem;
        sColB[i] = v ? p.colB[n0 + i] : 0;
        sColC[i] = split ? (int64_t)i * BM : (v ? p.colC[n0 + i] : 0);
    }
    const int64_t ab = p.batA[l], bb = p.batB[l];
    const int64_t cb = split ? ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (BM * BN) : p.batC[l];
    __syncthreads();

    constexpr int64_t OFFB = (int64_t)sizeof(E) / Core::KSTEP;   // bytes per table offset unit (one scalar element of the tensor)
    const char *gA = reinterpret_cast<const char *>(p.A) + ab * OFFB;
    const char *gB = reinterpret_cast<const char *>(p.B) + bb * OFFB;
    // this slice's k-blocks [kb0, kb0 + KB); k beyond the slice (or beyond K) is zero-filled
    const int64_t KB_all = (p.K + BK - 1) / BK;
    const int64_t kb0 = KB_all * blockIdx.y / gridDim.y, kb1 = KB_all * (blockIdx.y + 1) / gridDim.y;
    const int64_t KB = kb1 - kb0;
    const int64_t k_limit = min(p.K, kb1 * BK);

    // One "unit" = one element-granular cp.async. Units [0, UA) belong to A, [UA, UA+UB) to B. A unit's (row, k)
    // assignment is fixed per thread, so its global row pointer and smem slot are hoisted out of the k loop; the
    // only per-k-block input is the k-offset, read from shared memory (staged one block ahead, pre-scaled to
    // bytes, -1 = past the end of K). Invalid elements are zero-filled by the cp.async itself: no branches.
    constexpr int UA = (BM * BK + NT - 1) / NT, UB = (BN * BK + NT - 1) / NT, UNITS = UA + UB;
    const char *u_ptr[UNITS];   // global row base (bytes); nullptr = row outside the tile / unit unused
    int u_dst[UNITS];           // smem byte offset inside a stage (A and B share one stage block)
    int u_k[UNITS];             // index into the staged k-offsets: k for A, BK + k for B
    constexpr int STAGE_A = BK * LDA * (int)sizeof(E), STAGE_B = BK * LDB * (int)sizeof(E);
#pragma unroll
    for (int u = 0; u < UNITS; u++) {
        if (u < UA) {
            const int i = tid + u * NT;
            int m, k;
            if (p.a_kmajor) { k = i % BK; m = i / BK; } else { m = i % BM; k = i / BM; }
            const bool ok = i < BM * BK && m < mrem;
            u_ptr[u] = ok ? gA + sRowA[m] * OFFB : nullptr;
            u_dst[u] = (i < BM * BK) ? (k * LDA + m) * (int)sizeof(E) : -1;
            u_k[u] = k;
        } else {
            const int i = tid + (u - UA) * NT;
            int n, k;
            if (p.b_kmajor) { k = i % BK; n = i / BK; } else { n = i % BN; k = i / BN; }
            const bool ok = i < BN * BK && n < nrem;
            u_ptr[u] = ok ? gB + sColB[n] * OFFB : nullptr;
            u_dst[u] = (i < BN * BK) ? (k * LDB + n) * (int)sizeof(E) : -1;
            u_k[u] = BK + k;
        }
    }
    const unsigned sA_u32 = (unsigned)__cvta_generic_to_shared(sA), sB_u32 = (unsigned)__cvta_generic_to_shared(sB);
    auto issue_unit = [&](int stage, int u, const int64_t *koff) {
        if ((u < UA ? (BM * BK) % NT : (BN * BK) % NT) != 0 && u_dst[u] < 0) return;   // only partial-tile variants
        const int64_t ko = koff[u_k[u]];
        const bool v = u_ptr[u] != nullptr && ko >= 0;
        const char *src = v ? u_ptr[u] + ko : reinterpret_cast<const char *>(p.A);
        const unsigned dst = (u < UA ? sA_u32 + stage * STAGE_A : sB_u32 + stage * STAGE_B) + u_dst[u];
        const int bytes = v ? (int)sizeof(E) : 0;
        if constexpr (sizeof(E) == 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes));
        else if constexpr (sizeof(E) == 8)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes));
        else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes));
    };

    typename Core::Acc acc;
    Core::init(acc);

    // k-offsets are staged through shared memory one k-block ahead of their use (pre-scaled to bytes, -1 past K):
    // the gather units read them with a short LDS instead of stalling the in-order MMA warps on a global load.
    auto stage_koff = [&](int64_t kblock, int slot) {
        if (tid < 2 * BK) {
            const int k = tid % BK;
            const int64_t kg = (kb0 + kblock) * BK + k;
            int64_t v = -1;
            if (kg < k_limit) v = ((tid < BK) ? p.kA[kg * Core::KSTEP] : p.kB[kg * Core::KSTEP]) * OFFB;
            sK[slot * 2 * BK + tid] = v;
        }
    };
    // prologue: stages 0..S-2
    for (int s = 0; s < S - 1; s++) {
        stage_koff(s, 0);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < UNITS; u++) issue_unit(s, u, sK);
        cp_async_commit();
        __syncthreads();
    }
    stage_koff(S - 1, 0);   // offsets of the first k-block gathered inside the main loop
    for (int64_t kb = 0; kb < KB; kb++) {
        cp_async_wait<S - 2>();
        __syncthreads();     // stage kb landed; sK[kb & 1] (written one iteration ago / in the prologue) visible
        const int64_t nk = kb + S - 1;
        const int nstage = (int)(nk % S);
        const int st = (int)(kb % S);
        const int64_t *koff = sK + (kb & 1) * 2 * BK;
        stage_koff(nk + 1, (int)((kb + 1) & 1));   // for the next iteration's gather
        // the next stage's gather is spread over the MMA stream: slot s issues units [s*UNITS/SLOTS, ...)
        Core::compute(acc, sA + (size_t)st * BK * LDA, sB + (size_t)st * BK * LDB, warp, lane, [&](int slot) {
            constexpr int SLOTS = Core::SLOTS;
            constexpr int PER = (UNITS + SLOTS - 1) / SLOTS;
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int u = slot * PER + q;
                if (u < UNITS) issue_unit(nstage, u, koff);
            }
            if (slot == SLOTS - 1) cp_async_commit();
        });
    }
    cp_async_wait<0>();
    if (split) {
        ScatterDesc none{};
        Core::store(acc, ws, sRowC, sColC, cb, BM, BN, warp, lane, none);   // whole tile: rows / columns past the edge are zeros
    } else {
        Core::store(acc, reinterpret_cast<E *>(p.C), sRowC, sColC, cb, mrem, nrem, warp, lane, p.sc);
    }
}

// Persistent form of gett_kernel for contractions with several tiles per SM: CTA b walks tiles b, b + gridDim.x, ... and the
// (tile, k-block) pairs form ONE software pipeline — during the last k-block of a tile the gather units already fetch the first
// k-block of the next tile (their row pointers are rebuilt from the next tile's offset tables, which arrived by cp.async many
// k-blocks earlier), and the epilogue stores of a tile run while that gather is in flight. What a tile no longer pays: the CTA
// launch, the table loads, the exposed first gather (98 KB per CTA) and the drain before the stores. Pays off for short sums
// (K <= 512: 8192 x 8192 x 256 30.6 -> 32.1 TFLOP/s); for K >= 1024 the plain kernel is ~1 % faster and stays in use.
// No split-K, no scatter epilogue.
template <class Core>
__global__ void __launch_bounds__(Core::NTHREADS, 1) gett_persistent_kernel(const __grid_constant__ GettParams p, int64_t ntiles_run) {
    using E = typename Core::Elem;
    constexpr int BM = Core::BM, BN = Core::BN, BK = Core::BK, S = Core::STAGES;
    constexpr int LDA = Core::LDA, LDB = Core::LDB, NT = Core::NTHREADS;
    static_assert(S == 2, "the cross-tile pipeline below assumes a prefetch distance of one k-block");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *sA = reinterpret_cast<E *>(smem_raw);
    E *sB = sA + (size_t)S * BK * LDA;
    int64_t *sTab = reinterpret_cast<int64_t *>(sB + (size_t)S * BK * LDB);   // [2][rowA BM | colB BN | rowC BM | colC BN]
    constexpr int TABN = 2 * (BM + BN);
    int64_t *sK = sTab + 2 * TABN;                                            // [2][2][BK]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
    const int64_t tiles = tiles_m * tiles_n;
    const int64_t KB = (p.K + BK - 1) / BK;

    struct Tile { int64_t m0, n0, l; int mrem, nrem; };
    auto tile_of = [&](int64_t bid) {
        Tile t;
        t.l = bid / tiles;
        const int64_t r = bid % tiles;
        const int64_t per_group = GROUP_M * tiles_n, g = r / per_group, gm0 = g * GROUP_M;
        const int64_t gsz = (tiles_m - gm0) < GROUP_M ? (tiles_m - gm0) : GROUP_M;
        t.m0 = (gm0 + (r % per_group) % gsz) * BM;
        t.n0 = ((r % per_group) / gsz) * BN;
        t.mrem = (int)((p.M - t.m0) < BM ? (p.M - t.m0) : BM);
        t.nrem = (int)((p.N - t.n0) < BN ? (p.N - t.n0) : BN);
        return t;
    };
    // offset tables of a tile -> smem set `buf`, by 8-byte cp.async (rows / columns past the edge read entry 0: never used)
    auto issue_tables = [&](const Tile &t, int buf) {
        int64_t *d = sTab + buf * TABN;
        for (int i = tid; i < BM; i += NT) {
            const int64_t m = t.m0 + (i < t.mrem ? i : 0);
            cp_async<8>(d + i, p.rowA + m);
            cp_async<8>(d + BM + BN + i, p.rowC + m);
        }
        for (int i = tid; i < BN; i += NT) {
            const int64_t n = t.n0 + (i < t.nrem ? i : 0);
            cp_async<8>(d + BM + i, p.colB + n);
            cp_async<8>(d + 2 * BM + BN + i, p.colC + n);
        }
    };

    constexpr int UA = (BM * BK + NT - 1) / NT, UB = (BN * BK + NT - 1) / NT, UNITS = UA + UB;
    static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "full tiles of units only");
    const char *u_ptr[UNITS];   // global row base (bytes) of the tile being GATHERED; nullptr = row outside the tile / no tile
    int u_dst[UNITS], u_k[UNITS];
    constexpr int STAGE_A = BK * LDA * (int)sizeof(E), STAGE_B = BK * LDB * (int)sizeof(E);
#pragma unroll
    for (int u = 0; u < UNITS; u++) {   // the (row, k) slot of a unit never changes
        const int i = tid + (u < UA ? u : u - UA) * NT;
        int r, k;
        if (u < UA) { if (p.a_kmajor) { k = i % BK; r = i / BK; } else { r = i % BM; k = i / BM; } u_dst[u] = (k * LDA + r) * (int)sizeof(E); u_k[u] = k; }
        else { if (p.b_kmajor) { k = i % BK; r = i / BK; } else { r = i % BN; k = i / BN; } u_dst[u] = (k * LDB + r) * (int)sizeof(E); u_k[u] = BK + k; }
    }
    auto build_units = [&](const Tile &t, int buf, bool exists) {
        const int64_t *d = sTab + buf * TABN;
        const E *gA = reinterpret_cast<const E *>(p.A) + (exists ? p.batA[t.l] : 0);
        const E *gB = reinterpret_cast<const E *>(p.B) + (exists ? p.batB[t.l] : 0);
#pragma unroll
        for (int u = 0; u < UNITS; u++) {
            const int i = tid + (u < UA ? u : u - UA) * NT;
            if (u < UA) {
                const int m = p.a_kmajor ? i / BK : i % BM;
                u_ptr[u] = (exists && m < t.mrem) ? reinterpret_cast<const char *>(gA + d[m]) : nullptr;
            } else {
                const int n = p.b_kmajor ? i / BK : i % BN;
                u_ptr[u] = (exists && n < t.nrem) ? reinterpret_cast<const char *>(gB + d[BM + n]) : nullptr;
            }
        }
    };
    const unsigned sA_u32 = (unsigned)__cvta_generic_to_shared(sA), sB_u32 = (unsigned)__cvta_generic_to_shared(sB);
    auto issue_unit = [&](int stage, int u, const int64_t *koff) {
        const int64_t ko = koff[u_k[u]];
        const bool v = u_ptr[u] != nullptr && ko >= 0;
        const char *src = v ? u_ptr[u] + ko : reinterpret_cast<const char *>(p.A);
        const unsigned dst = (u < UA ? sA_u32 + stage * STAGE_A : sB_u32 + stage * STAGE_B) + u_dst[u];
        const int bytes = v ? (int)sizeof(E) : 0;
        static_assert(sizeof(E) == 16, "ComplexF64 tiles");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes));
    };
    // byte offsets of k-block `kblock` (of any tile: they depend on k only) -> sK slot
    auto stage_koff = [&](int64_t kblock, int slot) {
        if (tid < 2 * BK) {
            const int64_t kg = kblock * BK + tid % BK;
            int64_t v = -1;
            if (kg < p.K) v = ((tid < BK) ? p.kA[kg] : p.kB[kg]) * (int64_t)sizeof(E);
            sK[slot * 2 * BK + tid] = v;
        }
    };

    int64_t bid = blockIdx.x;
    if (bid >= ntiles_run) return;
    Tile cur = tile_of(bid);
    issue_tables(cur, 0);
    cp_async_commit();
    stage_koff(0, 0);
    cp_async_wait<0>();
    __syncthreads();
    build_units(cur, 0, true);
#pragma unroll
    for (int u = 0; u < UNITS; u++) issue_unit(0, u, sK);
    cp_async_commit();
    __syncthreads();
    stage_koff(1 % KB, 1);   // offsets of global block 1 (gathered during block 0)

    typename Core::Acc acc;
    Core::init(acc);
    int buf = 0;
    int64_t g = 0;   // global k-block counter of this CTA: stage g & 1 is computed, stage (g + 1) & 1 is gathered
    for (;;) {
        const int64_t nbid = bid + gridDim.x;
        const bool has_next = nbid < ntiles_run;
        const Tile nxt = has_next ? tile_of(nbid) : cur;
        for (int64_t kb = 0; kb < KB; kb++, g++) {
            cp_async_wait<0>();
            __syncthreads();     // block g landed; sK[(g + 1) & 1] (written one iteration ago) visible; previous epilogue finished
            if (kb == 0 && has_next) issue_tables(nxt, buf ^ 1);                  // lands long before the last block of this tile
            if (kb == KB - 1) build_units(nxt, buf ^ 1, has_next);                // from here on the gather belongs to the next tile
            const int64_t *koff = sK + ((g + 1) & 1) * 2 * BK;
            stage_koff((kb + 2) % KB, (int)(g & 1));                              // block g + 2, for the next iteration's gather
            const int st = (int)(g & 1), nstage = st ^ 1;
            Core::compute(acc, sA + (size_t)st * BK * LDA, sB + (size_t)st * BK * LDB, warp, lane, [&](int slot) {
                constexpr int SLOTS = Core::SLOTS;
                constexpr int PER = (UNITS + SLOTS - 1) / SLOTS;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    const int u = slot * PER + q;
                    if (u < UNITS) issue_unit(nstage, u, koff);
                }
                if (slot == SLOTS - 1) cp_async_commit();
            });
        }
        // epilogue of this tile; the first k-block of the next one is in flight
        const int64_t *d = sTab + buf * TABN;
        Core::store(acc, reinterpret_cast<E *>(p.C), d + BM + BN, d + 2 * BM + BN, p.batC[cur.l], cur.mrem, cur.nrem, warp, lane, p.sc);
        if (!has_next) break;
        Core::init(acc);
        cur = nxt;
        bid = nbid;
        buf ^= 1;
    }
    cp_async_wait<0>();
}

template <typename E> __device__ __forceinline__ E add_e(E a, E b);
template <> __device__ __forceinline__ float add_e(float a, float b) { return a + b; }
template <> __device__ __forceinline__ double add_e(double a, double b) { return a + b; }
template <> __device__ __forceinline__ float2 add_e(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
template <> __device__ __forceinline__ double2 add_e(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// C[rowC[m] + colC[n] + batC[l]] = sum over slices of ws[(s * ntiles + tile) * BM * BN + c * BM + r]
template <typename E, int BM, int BN>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const __grid_constant__ GettParams p, const E *__restrict__ ws, int nsplit,
                                                            int64_t ntiles, int64_t tile0) {
    // one thread per output element (tile-linear index: consecutive threads read consecutive workspace elements and,
    // rows being C's fastest mode, write consecutive C elements); slices added in order, four loads in flight
    const int64_t tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
    const int64_t tiles = tiles_m * tiles_n;
    const int64_t total = ntiles * (BM * BN), stride = (int64_t)gridDim.x * blockDim.x, slice = ntiles * (BM * BN);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t tile = idx / (BM * BN);
        const int e = (int)(idx - tile * (BM * BN));
        const int r = e % BM, c = e / BM;
        const int64_t l = (tile0 + tile) / tiles, t = (tile0 + tile) % tiles;
        const int64_t per_group = GROUP_M * tiles_n, g = t / per_group, gm0 = g * GROUP_M;
        const int64_t gsz = (tiles_m - gm0) < GROUP_M ? (tiles_m - gm0) : GROUP_M;
        const int64_t m0 = (gm0 + (t % per_group) % gsz) * BM, n0 = ((t % per_group) / gsz) * BN;
        if (m0 + r >= p.M || n0 + c >= p.N) continue;
        const E *w = ws + idx;
        E acc = w[0];
        int s = 1;
        for (; s + 3 < nsplit; s += 4) {
            const E v0 = w[(int64_t)s * slice], v1 = w[(int64_t)(s + 1) * slice], v2 = w[(int64_t)(s + 2) * slice], v3 = w[(int64_t)(s + 3) * slice];
            acc = add_e(add_e(add_e(add_e(acc, v0), v1), v2), v3);
        }
        for (; s < nsplit; s++) acc = add_e(acc, w[(int64_t)s * slice]);
        reinterpret_cast<E *>(p.C)[p.rowC[m0 + r] + p.colC[n0 + c] + p.batC[l]] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming variant for "a small operator applied to a huge tensor" (N <= BN, K <= 32, no batch): the MPS-MPO
// middle step (1 048 576 x 16 x 16) and gate application. HBM-bound (AI ~ 4 flop/B), so the structure is a copy
// kernel with a DMMA in the middle: persistent CTAs, the whole B operand resident in shared memory, and a software
// pipeline over row tiles — while tile i is multiplied and stored, the elements of tile i+1 and the row-offset
// tables of tile i+2 are in flight (cp.async), so no tile ever exposes its load latency.
template <class Core>
__global__ void __launch_bounds__(Core::NTHREADS, 4) stream_kernel(const __grid_constant__ GettParams p, int kpad) {
    using E = typename Core::Elem;
    constexpr int BM = Core::BM, BN = Core::BN, LDA = Core::LDA, LDB = Core::LDB, NT = Core::NTHREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *sA = reinterpret_cast<E *>(smem_raw);                       // [2][kpad][LDA]
    E *sB = sA + (size_t)2 * kpad * LDA;                           // [kpad][LDB]
    int64_t *sRowA = reinterpret_cast<int64_t *>(sB + (size_t)kpad * LDB);   // [3][BM]
    int64_t *sRowC = sRowA + 3 * BM;                               // [3][BM]
    int64_t *sColC = sRowC + 3 * BM;                               // [BN]
    int64_t *sKA = sColC + BN;                                     // [kpad] byte offsets, -1 past K
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (p.M + BM - 1) / BM;
    const int nrem = (int)p.N;
    const E *gA = reinterpret_cast<const E *>(p.A) + p.batA[0];
    const E *gB = reinterpret_cast<const E *>(p.B) + p.batB[0];
    const int64_t cb = p.batC[0];

    // resident operand and tables
    for (int i = tid; i < kpad * BN; i += NT) {
        const int n = i % BN, k = i / BN;
        E v{};
        if (n < p.N && k < p.K) v = gB[p.colB[n] + p.kB[k]];
        sB[k * LDB + n] = v;
    }
    for (int i = tid; i < BN; i += NT) sColC[i] = i < p.N ? p.colC[i] : 0;
    for (int i = tid; i < kpad; i += NT) sKA[i] = i < p.K ? p.kA[i] * (int64_t)sizeof(E) : -1;

    auto issue_tables = [&](int64_t tile, int slot) {      // 8-byte cp.async of the tile's row offsets
        const int64_t m0 = tile * BM;
        for (int i = tid; i < BM; i += NT) {
            const bool v = m0 + i < p.M;
            const unsigned da = (unsigned)__cvta_generic_to_shared(sRowA + slot * BM + i);
            const unsigned dc = (unsigned)__cvta_generic_to_shared(sRowC + slot * BM + i);
            const int bytes = v ? 8 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(da), "l"(p.rowA + (v ? m0 + i : 0)), "r"(bytes));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dc), "l"(p.rowC + (v ? m0 + i : 0)), "r"(bytes));
        }
    };
    auto issue_elements = [&](int64_t tile, int stage, int slot) {
        const int mrem = (int)min((int64_t)BM, p.M - tile * BM);
        const int64_t *rows = sRowA + slot * BM;
        E *dstb = sA + (size_t)stage * kpad * LDA;
        for (int i = tid; i < BM * kpad; i += NT) {
            int m, k;
            if (p.a_kmajor) { k = i % kpad; m = i / kpad; } else { m = i % BM; k = i / BM; }
            const int64_t ko = sKA[k];
            const bool v = m < mrem && ko >= 0;
            const char *src = reinterpret_cast<const char *>(gA) + (v ? rows[m] * (int64_t)sizeof(E) + ko : 0);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(dstb + k * LDA + m);
            const int bytes = v ? (int)sizeof(E) : 0;
            if constexpr (sizeof(E) == 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes));
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes));
        }
    };

    int64_t t0 = blockIdx.x;
    if (t0 >= ntiles) return;
    issue_tables(t0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();                                        // tables(t0), sB, sKA, sColC visible
    issue_elements(t0, 0, 0);
    if (t0 + gridDim.x < ntiles) issue_tables(t0 + gridDim.x, 1);
    cp_async_commit();
    int it = 0;
    for (int64_t tile = t0; tile < ntiles; tile += gridDim.x, it++) {
        cp_async_wait<0>();
        __syncthreads();                                    // elements(tile) and tables(next) landed; previous tile's readers done
        const int64_t nxt = tile + gridDim.x, nxt2 = nxt + gridDim.x;
        if (nxt < ntiles) issue_elements(nxt, (it + 1) & 1, (it + 1) % 3);
        if (nxt2 < ntiles) issue_tables(nxt2, (it + 2) % 3);
        cp_async_commit();
        typename Core::Acc acc;
        Core::init(acc);
        const E *a = sA + (size_t)(it & 1) * kpad * LDA;
        for (int kb = 0; kb < kpad; kb += Core::BK)
            Core::compute(acc, a + (size_t)kb * LDA, sB + (size_t)kb * LDB, warp, lane, [](int) {});
        const int mrem = (int)min((int64_t)BM, p.M - tile * BM);
        Core::store(acc, reinterpret_cast<E *>(p.C), sRowC + (it % 3) * BM, sColC, cb, mrem, nrem, warp, lane, p.sc);
    }
    cp_async_wait<0>();
}

template <class Core>
cudaError_t launch_stream(const GettParams &p, cudaStream_t s) {
    using E = typename Core::Elem;
    const int kpad = (int)((p.K + Core::BK - 1) / Core::BK) * Core::BK;
    const size_t smem = ((size_t)2 * kpad * Core::LDA + (size_t)kpad * Core::LDB) * sizeof(E) +
                        (size_t)(6 * Core::BM + Core::BN + kpad) * sizeof(int64_t);
    const int64_t ntiles = (p.M + Core::BM - 1) / Core::BM;
    const int64_t grid = std::min<int64_t>(ntiles, 148 * 5);
    stream_kernel<Core><<<(unsigned)grid, Core::NTHREADS, smem, s>>>(p, kpad);
    return cudaGetLastError();
}
template <class Core>
cudaError_t configure_stream() {
    return cudaFuncSetAttribute(stream_kernel<Core>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

template <class Core>
cudaError_t launch(const GettParams &p, cudaStream_t s) {
    const int64_t tiles_m = (p.M + Core::BM - 1) / Core::BM, tiles_n = (p.N + Core::BN - 1) / Core::BN;
    const int64_t grid = tiles_m * tiles_n * p.L;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    using E = typename Core::Elem;
    using RE = typename std::conditional<Core::KSTEP == 2, double, E>::type;   // element type of C and of the split-K partial tiles
    // split-K when the tiles cannot fill the GPU: slices = how many times the tile set fits into the resident CTA
    // slots, at most one slice per k-block (a k-block of work per CTA is the granularity of the pipeline)
    static const int sk_mode = [] { const char *e = getenv("MB200_SPLITK"); return e ? atoi(e) : 1; }();
    const int64_t KB = (p.K + Core::BK - 1) / Core::BK;
    const int64_t slots = 148 * (Core::NTHREADS <= 128 ? 4 : 1);
    int64_t nsplit = 1;
    if (sk_mode && p.sc.nranks == 0 && grid * 4 <= slots * 3 && KB >= 2) nsplit = std::min<int64_t>(KB, std::max<int64_t>(1, slots / grid));
    if (nsplit > 1) {
        E *ws = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&ws, (size_t)nsplit * grid * Core::BM * Core::BN * sizeof(E), s);
        if (e != cudaSuccess) return e;
        gett_kernel<Core><<<dim3((unsigned)grid, (unsigned)nsplit), Core::NTHREADS, gett_smem_bytes<Core>(), s>>>(p, ws, 0);
        splitk_reduce_kernel<RE, Core::BM, Core::BN><<<(unsigned)std::min<int64_t>((grid * Core::BM * Core::BN + 255) / 256, 148 * 16), 256, 0, s>>>(p, (const RE *)ws, (int)nsplit, grid, 0);
        e = cudaGetLastError();
        cudaFreeAsync(ws, s);
        return e;
    }
    // Ragged last wave: the tiles beyond the last full wave (at most half a wave of them) run as a second launch with their
    // k-range split so that they fill the SMs once more for 1 / nsplit of a tile time (512 tiles on 148 SMs: 4 -> 3.5 tile times).
    const int64_t tail = grid % slots, full = grid - tail;
    const bool split_tail = sk_mode == 1 && p.sc.nranks == 0 && full > 0 && tail > 0 && tail * 2 <= slots && KB >= 8;   // MB200_SPLITK=2: A/B without it
    const int64_t run = split_tail ? full : grid;   // tiles of the main launch
    // main launch: persistent (cross-tile pipeline) for the ComplexF64 main tile when every SM gets several tiles
    static const int persist_mode = [] { const char *e = getenv("MB200_PERSIST"); return e ? atoi(e) : 1; }();
    bool launched = false;
    if constexpr (Core::STAGES == 2 && Core::KSTEP == 1 && sizeof(E) == 16 && (Core::BM * Core::BK) % Core::NTHREADS == 0 &&
                  (Core::BN * Core::BK) % Core::NTHREADS == 0 && (Core::NTHREADS > 128)) {
        // measured (tools/ab_tail.py): 8192 x 8192 x 256 30.6 -> 32.1 TFLOP/s, but K >= 1024 loses ~1 % (the extra control flow in the
        // k loop costs more than the hidden prologue / epilogue gains once a tile runs for 32+ k-blocks) -> short-K shapes only
        if (persist_mode && p.sc.nranks == 0 && KB >= 4 && (KB <= 16 || persist_mode == 2) && run >= 2 * 148) {
            gett_persistent_kernel<Core><<<148, Core::NTHREADS, gett_persistent_smem_bytes<Core>(), s>>>(p, run);
            launched = true;
        }
    }
    if (!launched) gett_kernel<Core><<<(unsigned)run, Core::NTHREADS, gett_smem_bytes<Core>(), s>>>(p, nullptr, 0);
    if (split_tail) {
        const int64_t tsplit = std::min<int64_t>(KB / 2, slots / tail);
        E *ws = nullptr;
        cudaError_t e = cudaMallocAsync((void **)&ws, (size_t)tsplit * tail * Core::BM * Core::BN * sizeof(E), s);
        if (e != cudaSuccess) return e;
        gett_kernel<Core><<<dim3((unsigned)tail, (unsigned)tsplit), Core::NTHREADS, gett_smem_bytes<Core>(), s>>>(p, ws, full);
        splitk_reduce_kernel<RE, Core::BM, Core::BN><<<(unsigned)std::min<int64_t>((tail * Core::BM * Core::BN + 255) / 256, 148 * 16), 256, 0, s>>>(p, (const RE *)ws, (int)tsplit, tail, full);
        e = cudaGetLastError();
        cudaFreeAsync(ws, s);
        return e;
    }
    return cudaGetLastError();
}
template <class Core>
cudaError_t configure() {
    if constexpr (Core::STAGES == 2 && Core::KSTEP == 1 && sizeof(typename Core::Elem) == 16 && (Core::BM * Core::BK) % Core::NTHREADS == 0 &&
                  (Core::BN * Core::BK) % Core::NTHREADS == 0 && (Core::NTHREADS > 128)) {
        cudaError_t e = cudaFuncSetAttribute(gett_persistent_kernel<Core>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)gett_persistent_smem_bytes<Core>());
        if (e != cudaSuccess) return e;
    }
    return cudaFuncSetAttribute(gett_kernel<Core>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)gett_smem_bytes<Core>());
}

// tile menus ---------------------------------------------------------------------------------------
using Z_128x64 = CoreZ<128, 64, 32, 32, 32, 2>;  // main ComplexF64 tile: 8 warps x (32 x 32); long k-blocks: one CTA barrier per 32 k
using Z_128x64k8 = CoreZ<128, 64, 32, 32, 8, 4>; // short K (<= 16): do not zero-fill a 32-deep k-block
using Z_128x16 = CoreZ<128, 16, 16, 16, 8, 4>;   // skinny N
using Z_128x16s = CoreZ<128, 16, 32, 16, 8, 2>;  // skinny N and short K (MPS-MPO middle step: N = K = 16): 4 warps, 2 stages -> 4-5 CTAs/SM
using Z_64x16t = CoreZ<64, 16, 16, 16, 8, 2>;    // streaming kernel tile: 4 warps x (16 x 16), ~42 KB smem -> 5 CTAs/SM
using Z_16x128 = CoreZ<16, 128, 16, 16, 8, 4>;   // skinny M
using D_128x128 = CoreD<128, 128, 64, 32, 16, 3>; // Float64: 8 warps x (64 x 32); BK = 32 x 2 stages measured slower (28.1 vs 31.5 TFLOP/s at 8192^3)
using D2_128x128 = CoreD2<128, 128, 64, 32, 16, 3>; // Float64, K-major pairs: 8 warps x (64 x 32), 32 k per k-block
using D_128x16 = CoreD<128, 16, 16, 16, 8, 4>;
using D_16x128 = CoreD<16, 128, 16, 16, 8, 4>;
using C_128x64 = CoreF<float2, 128, 64, 16, 3>;  // ComplexF32 FFMA: thread tile 8 x 4
using C_128x16 = CoreF<float2, 128, 16, 8, 4>;
using C_16x128 = CoreF<float2, 16, 128, 8, 4>;
using S_128x128 = CoreF<float, 128, 128, 16, 3>; // Float32 FFMA: thread tile 8 x 8
using S_128x16 = CoreF<float, 128, 16, 8, 4>;
using S_16x128 = CoreF<float, 16, 128, 8, 4>;

}  // namespace

cudaError_t gett_configure() {
    cudaError_t e;
#define MB200_CFG(C) if ((e = configure<C>()) != cudaSuccess) return e
    if ((e = configure_stream<Z_64x16t>()) != cudaSuccess) return e;
    MB200_CFG(Z_128x64); MB200_CFG(Z_128x64k8); MB200_CFG(Z_128x16); MB200_CFG(Z_128x16s); MB200_CFG(Z_16x128);
    MB200_CFG(D_128x128); MB200_CFG(D2_128x128); MB200_CFG(D_128x16); MB200_CFG(D_16x128);
    MB200_CFG(C_128x64); MB200_CFG(C_128x16); MB200_CFG(C_16x128);
    MB200_CFG(S_128x128); MB200_CFG(S_128x16); MB200_CFG(S_16x128);
#undef MB200_CFG
    return cudaSuccess;
}

cudaError_t launch_gett_f64(int dtype, const GettParams &p, cudaStream_t s) {
    if (dtype == MB200_C128) {
        if (p.N <= 16 && p.K <= 32 && p.L == 1 && p.M >= 4096) return launch_stream<Z_64x16t>(p, s);   // streaming, HBM-bound
        if (p.N <= 16 && p.M > 16) return p.K <= 32 ? launch<Z_128x16s>(p, s) : launch<Z_128x16>(p, s);
        if (p.M <= 16 && p.N > 16) return launch<Z_16x128>(p, s);
        return p.K <= 16 ? launch<Z_128x64k8>(p, s) : launch<Z_128x64>(p, s);
    }
    if (p.N <= 16 && p.M > 16) return launch<D_128x16>(p, s);
    if (p.M <= 16 && p.N > 16) return launch<D_16x128>(p, s);
    static const int pairs_mode = [] { const char *e = getenv("MB200_F64_PAIRS"); return e ? atoi(e) : 1; }();
    if (pairs_mode && p.k_pairs && p.K >= 64 && p.K % 2 == 0 && ((((uintptr_t)p.A) | ((uintptr_t)p.B)) & 15) == 0) {
        GettParams q = p;
        q.K = p.K / 2;   // k is walked in pairs
        return launch<D2_128x128>(q, s);   // 8192^3: 31.4 -> 32.9 TFLOP/s; a 128 x 64 / BK = 32 pair tile measured slower (31.0)
    }
    return launch<D_128x128>(p, s);
}

cudaError_t launch_simt_f32(int dtype, const GettParams &p, cudaStream_t s) {
    if (dtype == MB200_C64) {
        if (p.N <= 16 && p.M > 16) return launch<C_128x16>(p, s);
        if (p.M <= 16 && p.N > 16) return launch<C_16x128>(p, s);
        return launch<C_128x64>(p, s);
    }
    if (p.N <= 16 && p.M > 16) return launch<S_128x16>(p, s);
    if (p.M <= 16 && p.N > 16) return launch<S_16x128>(p, s);
    return launch<S_128x128>(p, s);
}

}  // namespace mb200
