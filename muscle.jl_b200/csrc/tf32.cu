// K3 — ComplexF32 / Float32 GEMM on 5th-gen tensor cores: tcgen05.mma with accumulators in TMEM, operands fed by TMA,
// split-operand error compensation, 4M complex decomposition, fused permuting epilogue.
//
// Operand format (written by the K1 pack pass with the split writers, permute.cu): a K-major matrix [rows][4*K] of
// 32-bit words where every group of 8 k-values is stored as four consecutive 32 B chunks
//     | re_hi(k0..k7) | re_x(k0..k7) | im_hi(k0..k7) | im_x(k0..k7) |        (128 B per row per group)
// One TMA box row is one 128 B swizzle-128B line; each 32 B chunk is exactly one MMA operand (K = 8 tf32 or K = 16 bf16),
// so the MMA descriptor for chunk p is the tile descriptor advanced by 32*p bytes.
//
// Two split schemes (Tf32Params::mixed):
//   * mixed TF32 + BF16 (default). x_hi = tf32_rn(x); the x chunk holds bf16 PAIRS: (bf16(x), bf16(x - x_hi)) for the row
//     operand and (bf16(y - y_hi), bf16(y)) for the column operand. x*y = x_hi*y_hi (one kind::tf32 MMA, K = 8) +
//     [x*y_lo + x_lo*y] (ONE kind::f16 bf16 MMA, K = 16, same tensor cycles and smem bytes as a K = 8 tf32 MMA) + O(2^-22):
//     2 MMAs per real product instead of 3. The cross terms are ~2^-11 of the product, so bf16's 2^-9 rounding leaves
//     ~5e-7 relative error (simulated in fp64; measured with the TMEM accumulation: see DESIGN 3.4).
//     Per group of 8 k the MMA warp issues 8 MMAs (2 x 4M) into two TMEM accumulators:
//         D_re += rh*rh' + rx.rx' - (ih*ih' + ix.ix')      (minus = a_negate in the idesc)
//         D_im += rh*ih' + rx.ix' +  ih*rh' + ix.rx'
//   * 3xTF32 (MB200_SPLIT_SCHEME=3xtf32): x_hi = x & 0xFFFFE000, the x chunk is the fp32 remainder x_lo; 12 MMAs:
//         D_re += rh*rh' + rh*rl' + rl*rh' - (ih*ih' + ih*il' + il*ih')
//         D_im += rh*ih' + rh*il' + rl*ih' +  ih*rh' + ih*rl' + il*rh'
// Effective ceiling in complex-flop terms: TF32 dense peak / 2 (mixed: 8 tf32-equivalent MACs per complex MAC = 8 flops)
// or / 3 (3xTF32).
//
// Float32 (REAL) operands use the same 128 B lines with two 8-k groups per line: | hi(k0..7) | x(k0..7) | hi(k8..15) |
// x(k8..15) |, 4 (mixed) or 6 (3xTF32) MMAs per line into ONE accumulator, and a 128 x 256 tile (N = 256 keeps the smem
// operand traffic per MMA at 12 KB / 128 cycles, under the 128 B/clk limit).
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..9 = epilogue (tcgen05.ld 32 lanes x BN/2 columns each -> running sums in registers ->
// C[rowC[m] + colC[n] + batC[l]]). Pipelines: full/empty mbarriers over a 6-stage smem ring (TMA <-> MMA),
// tmem_full/tmem_empty over two TMEM chunk buffers (MMA <-> epilogue). Persistent: one CTA per SM walks the
// 128 x BN output tiles.
//
// Measured limits (config 3): the path runs at the 1000 W board power cap (tools/power_probe.py: 982 W, sw_power_cap,
// SM clock 1.67-1.70 GHz), so sustained throughput is set by energy per flop. ncu on the 1-CTA kernel: tensor sub-pipes 61.6 %
// active, shared-memory data pipe 89 % busy (MMA operand reads 59.6 % + TMA fill 29.8 %: at N = 128 a K=8 tf32 MMA reads 8 KB of
// operands per 64 tensor cycles = all 128 B/clk of shared memory; forcing BN = 64 makes the GEMM 1.77x slower) -> CTA pairs.
// Tried and not kept: six N = 256 MMAs over concatenated [re;im] column-operand planes (25 % fewer smem bytes, GEMM 2.21 -> 2.13 ms
// but the pack and DRAM traffic grow by the same amount); same-kind grouping of the mixed scheme's MMAs (slower than interleaved);
// stream-K (balanced (tile, 128-k chunk) unit ranges per CTA, tiles cut between two CTAs finished through a flag-ordered workspace
// hand-over; verified bit-exact): for 256 tiles on 148 SMs (one 2048^3 batch of config 3) it gained 2.6 % instead of the 13 % the
// wave arithmetic promises, and contiguous unit ranges lost up to 18 % elsewhere (CTAs walk distant tiles concurrently, the packed
// panels stop hitting in L2) - under a power cap SMs idle in a ragged last wave let the busy ones clock higher.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <type_traits>

#include "kernels.cuh"

namespace mb200 {

namespace {

constexpr int TBM = 128;          // tile rows = TMEM lanes
constexpr int TTHREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int GROUP_BYTES = 128;  // bytes per row per smem line: one 8-k group (complex) or two (real)
constexpr int CHUNK_K = 128;      // k per TMEM accumulation chunk (two-level accumulation)

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// CTA-pair form: executed by both CTAs of the pair, each into its own shared memory; the transaction bytes are
// counted on the LEADER CTA's mbarrier (peer bit of the shared::cluster address cleared, as cute::SM100_TMA_2SM_LOAD does)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_pair(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
template <bool CTA2>
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    if constexpr (CTA2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
}
template <bool CTA2>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T. BF16 = false: kind::tf32 (K = 8); true: kind::f16 with bf16 operands (K = 16).
// CTA2: issued by the leader CTA of a pair for both tensor cores (M = 256: 128 rows per CTA, each CTA holds N / 2 rows of B).
template <bool CTA2, bool BF16>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CTA2 && BF16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if constexpr (CTA2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if constexpr (BF16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive once all MMAs issued so far by this thread have completed; CTA2: on the barrier at this offset in BOTH CTAs
template <bool CTA2>
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    if constexpr (CTA2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (= 1, unused for swizzled K-major),
//   [32,46) stride byte offset >> 4 (= 1024 B between 8-row groups), [46,48) version = 1 (Blackwell),
//   [61,64) layout type = 2 (SWIZZLE_128B). Tiles are 1024 B aligned, so base_offset = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @4, a/b_format @7/@10 (TF32 = 2 for kind::tf32,
// BF16 = 1 for kind::f16), a_negate @13, a/b K-major (0) @15/@16, N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool neg_a, bool bf16 = false) {
    const uint32_t fmt = bf16 ? 1u : 2u;   // M = 256 only with cta_group::2
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((neg_a ? 1u : 0u) << 13) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Tf32Params {
    ScatterDesc sc;
    void *C;
    const int64_t *rowC, *colC, *batC;
    int64_t M, N, L;
    int64_t ntiles;
    int KG;   // number of smem lines along k: K / 8 (complex), ceil(K / 16) (real)
    // split-K (too few tiles for 148 SMs): work unit u = (tile u % ntiles, slice u / ntiles); slice s covers the smem
    // lines [KG * s / nsplit, KG * (s + 1) / nsplit) and leaves its partial tile in ws (tile-linear, row fastest);
    // tf32_splitk_reduce_kernel adds the slices in order and scatters through the C tables.
    int nsplit;
    void *ws;
    int mixed;   // 1: TF32 + BF16 operand format (8 / 4 MMAs per line), 0: 3xTF32 (12 / 6)
    // cross-GPU split-K (kernels.cuh DistDesc): partial sub-tiles go to ws (this rank's own workspace), then the unit's flag is
    // raised in its owner's flag array
    int dist_nranks, dist_rank, dist_epoch, dist_fence_all;
    int *dist_flags[MB200_MAX_PEERS];
    unsigned long long *timeline;
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int BN, bool REAL, bool CTA2>
struct Tf32Smem {
    static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;   // CTA pair: each CTA holds half of the column operand's rows
    static constexpr int A_BYTES = TBM * GROUP_BYTES;   // 16 KB
    static constexpr int B_BYTES = B_ROWS * GROUP_BYTES;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int TSTAGES = (192 * 1024) / STAGE;   // 6 x 32 KB, 4 x 48 KB; pair: 8 x 24 KB, 6 x 32 KB
    static constexpr int TOTAL = TSTAGES * STAGE + 1024 /*align*/ + 256 /*barriers*/ + 2 * BN * 8 /*colC, two tiles*/;
};

// Two-level accumulation. The tensor core adds into TMEM with truncation, so a long accumulation chain
// drifts linearly (measured: rel. error 1.45e-8 * K, i.e. 5.9e-5 at K = 4096). TMEM therefore only ever
// holds a chunk of CHUNK_GROUPS*8 = 128 k (double-buffered: the MMA warp fills buffer c&1 while the
// epilogue warps drain the other one); the running sums live in the epilogue warps' registers and are
// added with ordinary round-to-nearest FADDs.
// tile id -> (m0, n0, batch): grouped rasterisation, 8 row-tiles share a B column panel in L2
struct TileCoord { int m0, n0, l; };
template <int BN, int PM = TBM>   // PM = rows per tile: 128, or 256 for a CTA pair
__device__ __forceinline__ TileCoord tile_coord(const Tf32Params &p, int64_t tile) {
    const int64_t tiles_m = (p.M + PM - 1) / PM, tiles_n = (p.N + BN - 1) / BN;
    const int64_t per = tiles_m * tiles_n;
    const int64_t l = tile / per, t = tile % per;
    const int64_t gsz_full = 8, pg = gsz_full * tiles_n, g = t / pg, gm0 = g * gsz_full;
    const int64_t gsz = (tiles_m - gm0) < gsz_full ? (tiles_m - gm0) : gsz_full;
    const int64_t tm = gm0 + (t % pg) % gsz, tn = (t % pg) / gsz;
    return TileCoord{(int)(tm * PM), (int)(tn * BN), (int)l};
}

// Persistent: CTA b walks tiles b, b + gridDim.x, ... The three roles keep running counters (smem stage / TMEM
// chunk buffer and their mbarrier phases) across tiles, so the TMA producer prefetches the next tile's first
// stages while the MMA warp is still on the current tile's tail, and the MMA warp starts the next tile's first
// chunk while the epilogue warps are still storing the previous tile: per-tile prologue and epilogue are hidden
// (K = 512 slices: 0.362 -> 0.317 ms).
//
// CTA2 (cluster of two CTAs = one TPC, tcgen05 cta_group::2): the pair owns a 256 x BN tile. Each CTA loads its own 128
// rows of the row operand and HALF of the column operand's rows (BN / 2), the leader CTA (rank 0) issues M = 256 MMAs
// that drive both tensor cores, each CTA's accumulators sit in its own TMEM and are drained by its own epilogue warps.
// Why: the 1-CTA kernel is bound by the shared-memory data pipe — ncu (mixed scheme, config 3): MMA operand reads 59.6 %
// + TMA fill 29.8 % of the pipe's peak, tensor pipe 61.6 % active. With the pair every MMA reads 6 KB instead of 8 KB per
// CTA and a stage is 24 KB instead of 32 KB. Barriers: full[] lives in the leader (its expect_tx covers both CTAs'
// bytes; the peer's TMA completes on it through the cleared peer bit), empty[] and tmem_full[] exist in both CTAs and are
// signalled by multicast commits, tmem_empty[] lives in the leader and counts the epilogue warps of both CTAs.
template <int BN, bool REAL, bool CTA2>
__global__ void __launch_bounds__(TTHREADS + 32, 1)   // dist mode launches an 11th warp: the flag signaller
tf32_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ Tf32Params p) {
    using SM = Tf32Smem<BN, REAL, CTA2>;
    constexpr int TSTAGES = SM::TSTAGES;
    constexpr int NCTA = CTA2 ? 2 : 1;
    constexpr int PM = TBM * NCTA;                // rows per (pair) tile
    constexpr int CHUNK_GROUPS = CHUNK_K / (REAL ? 16 : 8);   // smem lines per TMEM chunk
    constexpr int HALF = BN / 2;                  // columns per epilogue thread
    constexpr uint32_t BUF_COLS = REAL ? BN : 2 * BN;   // D (real) or D_re | D_im
    constexpr uint32_t TMEM_COLS = 2 * BUF_COLS;  // two chunk buffers (256 or 512: powers of two)
    extern __shared__ unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(tiles + TSTAGES * SM::STAGE);
    uint64_t *empty = full + TSTAGES;
    uint64_t *tmem_full = empty + TSTAGES;        // [2]
    uint64_t *tmem_empty = tmem_full + 2;         // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    uint32_t *sig_posted = tmem_slot + 1;         // dist mode: units whose partial stores are complete (epilogue -> signal warp)
    int64_t *sColC = reinterpret_cast<int64_t *>(tiles + TSTAGES * SM::STAGE + 256);   // [2][BN], per tile parity

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.timeline && threadIdx.x == 0) atomicMin(&p.timeline[0], global_ns());
    const int64_t ntiles = p.ntiles, nunits = p.ntiles * p.nsplit;
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;          // 0 = leader
    const int64_t walker = CTA2 ? blockIdx.x / 2 : blockIdx.x;    // persistent walker id (CTA or CTA pair)
    const int64_t nwalkers = CTA2 ? gridDim.x / 2 : gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
        for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8 * NCTA); }
        *sig_posted = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc<CTA2>(tmem_slot, TMEM_COLS);
    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // pair: both CTAs' barriers and TMEM exist before any remote use
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer (every CTA: its rows of A, its share of B)
            uint32_t it = 0;   // running stage counter
            for (int64_t unit = walker; unit < nunits; unit += nwalkers) {
                const int64_t tile = unit % ntiles, sp = unit / ntiles;
                const TileCoord tc = tile_coord<BN, PM>(p, tile);
                const int kg0 = (int)((int64_t)p.KG * sp / p.nsplit), kg1 = (int)((int64_t)p.KG * (sp + 1) / p.nsplit);
                for (int kg = kg0; kg < kg1; kg++, it++) {
                    const int s = it % TSTAGES;
                    mbar_wait(&empty[s], ((it / TSTAGES) & 1) ^ 1);
                    unsigned char *a = tiles + s * SM::STAGE;
                    if constexpr (CTA2) {
                        if (rank == 0) mbar_expect_tx(&full[s], 2 * SM::STAGE);
                        tma_load_3d_pair(a, &mapA, &full[s], kg * 32, tc.m0 + (int)rank * TBM, tc.l);
                        tma_load_3d_pair(a + SM::A_BYTES, &mapB, &full[s], kg * 32, tc.n0 + (int)rank * SM::B_ROWS, tc.l);
                    } else {
                        mbar_expect_tx(&full[s], SM::STAGE);
                        tma_load_3d(a, &mapA, &full[s], kg * 32, tc.m0, tc.l);
                        tma_load_3d(a + SM::A_BYTES, &mapB, &full[s], kg * 32, tc.n0, tc.l);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {   // ---- MMA issuer (pair: the leader CTA only)
            constexpr uint32_t IDESC = make_idesc(PM, BN, false), IDESC_NEG = make_idesc(PM, BN, true);
            constexpr uint32_t IDESC_X = make_idesc(PM, BN, false, true), IDESC_XNEG = make_idesc(PM, BN, true, true);
            const bool mixed = p.mixed != 0;
            uint32_t it = 0, ch = 0;   // running stage / chunk counters
            for (int64_t unit = walker; unit < nunits; unit += nwalkers) {
                const int64_t sp = unit / ntiles;
                const int kg0 = (int)((int64_t)p.KG * sp / p.nsplit), kg1 = (int)((int64_t)p.KG * (sp + 1) / p.nsplit);
                const int nchunks = (kg1 - kg0 + CHUNK_GROUPS - 1) / CHUNK_GROUPS;
                int kg = kg0;
                for (int c = 0; c < nchunks; c++, ch++) {
                    const int buf = ch & 1;
                    mbar_wait(&tmem_empty[buf], ((ch >> 1) & 1) ^ 1);   // epilogue has drained this buffer
                    tc_fence_after();
                    const uint32_t d_re = tmem_base + buf * BUF_COLS, d_im = d_re + BN;
                    const int kend = min(kg1, kg0 + (c + 1) * CHUNK_GROUPS);
                    for (bool first = true; kg < kend; kg++, it++, first = false) {
                        const int s = it % TSTAGES;
                        mbar_wait(&full[s], (it / TSTAGES) & 1);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(tiles + s * SM::STAGE);
                        const uint64_t da = make_desc(sa), db = make_desc(sa + SM::A_BYTES);
                        // chunk c of a row is +32*c bytes: descriptor start-address field += 2*c
                        const uint64_t a_rh = da, a_rl = da + 2, a_ih = da + 4, a_il = da + 6;
                        const uint64_t b_rh = db, b_rl = db + 2, b_ih = db + 4, b_il = db + 6;
                        const uint32_t acc = first ? 0u : 1u;
                        constexpr bool T = false, X = true;   // MMA kind: tf32 / bf16
                        if constexpr (REAL) {   // line = hi0 | x0 | hi1 | x1: two 8-k groups
                            if (mixed) {
                                umma<CTA2, T>(d_re, a_rh, b_rh, IDESC, acc);
                                umma<CTA2, X>(d_re, a_rl, b_rl, IDESC_X, 1u);
                                umma<CTA2, T>(d_re, a_ih, b_ih, IDESC, 1u);
                                umma<CTA2, X>(d_re, a_il, b_il, IDESC_X, 1u);
                            } else {
                                umma<CTA2, T>(d_re, a_rh, b_rh, IDESC, acc);
                                umma<CTA2, T>(d_re, a_rh, b_rl, IDESC, 1u);
                                umma<CTA2, T>(d_re, a_rl, b_rh, IDESC, 1u);
                                umma<CTA2, T>(d_re, a_ih, b_ih, IDESC, 1u);
                                umma<CTA2, T>(d_re, a_ih, b_il, IDESC, 1u);
                                umma<CTA2, T>(d_re, a_il, b_ih, IDESC, 1u);
                            }
                        } else if (mixed) {     // chunks 1 / 3 are the bf16 cross-term pairs
                            umma<CTA2, T>(d_re, a_rh, b_rh, IDESC, acc);
                            umma<CTA2, T>(d_im, a_rh, b_ih, IDESC, acc);
                            umma<CTA2, X>(d_re, a_rl, b_rl, IDESC_X, 1u);
                            umma<CTA2, X>(d_im, a_rl, b_il, IDESC_X, 1u);
                            umma<CTA2, T>(d_re, a_ih, b_ih, IDESC_NEG, 1u);
                            umma<CTA2, T>(d_im, a_ih, b_rh, IDESC, 1u);
                            umma<CTA2, X>(d_re, a_il, b_il, IDESC_XNEG, 1u);
                            umma<CTA2, X>(d_im, a_il, b_rl, IDESC_X, 1u);
                        } else {
                            umma<CTA2, T>(d_re, a_rh, b_rh, IDESC, acc);
                            umma<CTA2, T>(d_im, a_rh, b_ih, IDESC, acc);
                            umma<CTA2, T>(d_re, a_rh, b_rl, IDESC, 1u);
                            umma<CTA2, T>(d_im, a_rh, b_il, IDESC, 1u);
                            umma<CTA2, T>(d_re, a_rl, b_rh, IDESC, 1u);
                            umma<CTA2, T>(d_im, a_rl, b_ih, IDESC, 1u);
                            umma<CTA2, T>(d_re, a_ih, b_ih, IDESC_NEG, 1u);
                            umma<CTA2, T>(d_im, a_ih, b_rh, IDESC, 1u);
                            umma<CTA2, T>(d_re, a_ih, b_il, IDESC_NEG, 1u);
                            umma<CTA2, T>(d_im, a_ih, b_rl, IDESC, 1u);
                            umma<CTA2, T>(d_re, a_il, b_ih, IDESC_NEG, 1u);
                            umma<CTA2, T>(d_im, a_il, b_rh, IDESC, 1u);
                        }
                        umma_commit<CTA2>(&empty[s]);          // frees the smem stage (in both CTAs) when these MMAs have read it
                    }
                    umma_commit<CTA2>(&tmem_full[buf]);        // this chunk's accumulators are complete
                }
            }
        }
    } else if (warp == 10) {   // ---- dist mode only: flag signaller (one thread). Unit k of this CTA is complete once the epilogue has
                               //      posted k + 1; then one cumulative system fence and the release store into the owner's flag array.
        if (lane == 0) {
            uint32_t k = 0;
            for (int64_t unit = walker; unit < nunits; unit += nwalkers, k++) {
                const int64_t u = (unit % ntiles) * NCTA + rank;
                const long long t0 = clock64();
                uint32_t seen;
                do {
                    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(sig_posted)) : "memory");
                    if (seen <= k) {
                        __nanosleep(128);
                        if (clock64() - t0 > 4000000000LL) __trap();
                    }
                } while (seen <= k);
                __threadfence_system();
                st_release_sys(p.dist_flags[u % p.dist_nranks] + u * p.dist_nranks + p.dist_rank, p.dist_epoch);
            }
        }
    } else {               // ---- epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp-2)/4
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;   // 0..255 among the epilogue threads
        uint32_t ch = 0, tcount = 0;
        for (int64_t unit = walker; unit < nunits; unit += nwalkers, tcount++) {
            const int64_t tile = unit % ntiles, sp = unit / ntiles;
            const int kg0 = (int)((int64_t)p.KG * sp / p.nsplit), kg1 = (int)((int64_t)p.KG * (sp + 1) / p.nsplit);
            const int nchunks = (kg1 - kg0 + CHUNK_GROUPS - 1) / CHUNK_GROUPS;
            const TileCoord tc = tile_coord<BN, PM>(p, tile);
            int64_t *cols = sColC + (tcount & 1) * BN;
            for (int i = et; i < BN; i += TTHREADS - 64) cols[i] = (tc.n0 + i < p.N) ? p.colC[tc.n0 + i] : 0;
            const int64_t m = (int64_t)tc.m0 + (int64_t)rank * TBM + row;
            const bool row_ok = m < p.M;
            const int64_t crow = (row_ok ? p.rowC[m] : 0) + p.batC[tc.l];
            float accr[HALF], acci[REAL ? 1 : HALF];
#pragma unroll
            for (int j = 0; j < HALF; j++) accr[j] = 0.f;
#pragma unroll
            for (int j = 0; j < (REAL ? 1 : HALF); j++) acci[j] = 0.f;
            for (int c = 0; c < nchunks; c++, ch++) {
                const int buf = ch & 1;
                mbar_wait(&tmem_full[buf], (ch >> 1) & 1);
                tc_fence_after();
                const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BUF_COLS + half * HALF;
#pragma unroll
                for (int sub = 0; sub < HALF / 32; sub++) {
                    uint32_t v[32];
                    tmem_ld32(tq + sub * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j++) accr[sub * 32 + j] += __uint_as_float(v[j]);
                    if constexpr (!REAL) {
                        tmem_ld32(tq + BN + sub * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j++) acci[sub * 32 + j] += __uint_as_float(v[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CTA2) mbar_arrive_cluster(&tmem_empty[buf], 0);   // the leader's MMA warp waits for both CTAs
                    else mbar_arrive(&tmem_empty[buf]);
                }
            }
            if (p.dist_nranks) {   // cross-GPU split-K: partial sub-tile -> own workspace (unit-linear, row fastest), then flag the owner
                const int64_t u = tile * NCTA + rank;
                const size_t base = (size_t)u * (TBM * BN) + row;
#pragma unroll
                for (int j = 0; j < HALF; j++) {
                    const size_t o = base + (size_t)(half * HALF + j) * TBM;
                    if constexpr (REAL) reinterpret_cast<float *>(p.ws)[o] = accr[j];
                    else reinterpret_cast<float2 *>(p.ws)[o] = make_float2(accr[j], acci[j]);
                }
                if (p.dist_fence_all) __threadfence_system();
                asm volatile("bar.sync 1, 256;" ::: "memory");   // the stores of all 256 epilogue threads are done (CTA scope)
                // Hand the unit to the signal warp: the system-scope fence and the remote flag store wait for NVLink round trips
                // (tens of microseconds while the reducers saturate the links) and must not sit in the epilogue's critical path -
                // measured: with the signalling thread inside the epilogue the overlapped GEMM ran 43 % slower per tile at 375 W.
                if (et == 0)
                    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(sig_posted)), "r"(tcount + 1) : "memory");
                continue;
            }
            if (p.nsplit > 1) {   // partial tile -> workspace, row fastest (coalesced), every row / column of the tile (1-CTA only)
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const size_t base = (size_t)unit * (TBM * BN) + row;
#pragma unroll
                for (int j = 0; j < HALF; j++) {
                    const size_t o = base + (size_t)(half * HALF + j) * TBM;
                    if constexpr (REAL) reinterpret_cast<float *>(p.ws)[o] = accr[j];
                    else reinterpret_cast<float2 *>(p.ws)[o] = make_float2(accr[j], acci[j]);
                }
                continue;
            }
            // the column table of this tile was written by all epilogue threads: named barrier over the 8 epilogue warps
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < HALF; j++) {
                    const int cidx = half * HALF + j;
                    if (tc.n0 + cidx < p.N) {
                        if constexpr (REAL) *scatter_ptr(p.sc, reinterpret_cast<float *>(p.C), crow + cols[cidx]) = accr[j];
                        else *scatter_ptr(p.sc, reinterpret_cast<float2 *>(p.C), crow + cols[cidx]) = make_float2(accr[j], acci[j]);
                    }
                }
            }
        }
        if (p.sc.nranks) __threadfence_system();   // peer stores must be visible before the cross-rank barrier
        if (p.timeline && et == 0) atomicMax(&p.timeline[1], global_ns());
    }
    tc_fence_before();
    if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // pair: no CTA frees TMEM / exits while the other still uses it
    if (warp == 1) tmem_dealloc<CTA2>(tmem_base, TMEM_COLS);
}

// C[rowC[m] + colC[n] + batC[l]] = sum over slices of ws[(s * ntiles + tile) * 128 * BN + col * 128 + row]
template <int BN, bool REAL>
__global__ void __launch_bounds__(256) tf32_splitk_reduce_kernel(const __grid_constant__ Tf32Params p) {
    using E = typename std::conditional<REAL, float, float2>::type;
    const E *ws = reinterpret_cast<const E *>(p.ws);
    const int64_t per = (int64_t)TBM * BN, total = p.ntiles * per, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int64_t tile = idx / per;
        const int e = (int)(idx - tile * per), r = e % TBM, c = e / TBM;
        const TileCoord tc = tile_coord<BN>(p, tile);
        if (tc.m0 + r >= p.M || tc.n0 + c >= p.N) continue;
        E acc = ws[idx];
        for (int s = 1; s < p.nsplit; s++) {
            const E v = ws[(int64_t)s * total + idx];
            if constexpr (REAL) acc += v;
            else { acc.x += v.x; acc.y += v.y; }
        }
        reinterpret_cast<E *>(p.C)[p.rowC[tc.m0 + r] + p.colC[tc.n0 + c] + p.batC[tc.l]] = acc;
    }
}

// ---- cross-GPU split-K: the owner's reducer (kernels.cuh DistDesc) -----------------------------------------------------------
struct DistParams {
    unsigned long long *timeline;
    const void *ws[MB200_MAX_PEERS];
    void *c[MB200_MAX_PEERS];
    int *flags[MB200_MAX_PEERS];
    const void *mc_ws;
    void *mc_c;
    int nranks, rank, epoch;
    int64_t nunits;
};

__device__ __forceinline__ float4 ld_relaxed_sys_f4(const void *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const void *p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f4(void *p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st_f2(void *p, float2 v) {
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void multimem_st_f1(void *p, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// bounded spin: a rank that never signals must surface as a trapped kernel, not as a hung GPU (~10 s at 2 GHz)
__device__ __forceinline__ void spin_until(const int *flag, int epoch) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        __nanosleep(64);
        if (clock64() - t0 > 20000000000LL) __trap();
    }
}

// Measured on the box (tools/diag_allreduce.py, globaltimer stamps): a reducer CTA and a GEMM CTA never share an SM, whatever the
// register / shared-memory arithmetic says - a small reducer CTA resident on an SM keeps the GEMM's CTA (pair) off it, and with the
// GEMM resident first the reducer only starts when GEMM CTAs retire. So the SMs are PARTITIONED: the GEMM's persistent grid leaves
// `reserve` SMs free (whole TPCs for the CTA-pair kernel: the reducer is then launched as 2-CTA clusters too) and the reducer runs
// fat CTAs (512 threads, 8 x 16 bytes in flight per thread) on exactly those SMs.
constexpr int RTHREADS = 512;
// Work item = half of a unit's vectors, one round of 8 vectors per thread: 20 reserved SMs hold 1.3 MB in flight, which the ~3.4 us
// NVLS round trip turns into ~390 GB/s per direction - what NCCL's own NVLS all-reduce gets on this box for 134 MB (0.383 ms). The
// fabric, not the reducer, is the limit: 16 vectors in flight per thread was measured SLOWER (0.523 vs 0.486 ms at N = 8: congestion
// stretches both the reduction and the GEMM's own traffic).
template <bool MC> struct RCfg { static constexpr int U = 8, PARTS = 2; };

// Work item = (owned unit, quarter); items are strided over the CTAs (one CTA per SM, next to the GEMM's CTA). A 16-byte vector is
// 2 consecutive rows (complex) or 4 (real) of one column of the sub-tile. Memory-level parallelism is what makes this kernel: a
// thread keeps 8 multimem.ld_reduce (MC) or 4 peer loads per step of the rank loop in flight - what the register budget of a
// co-resident CTA allows - and the grid supplies the rest (148 x 128 x 8 x 16 B = 2.4 MB in flight per GPU).
template <int BN, bool REAL, bool CTA2, bool MC>
__global__ void __launch_bounds__(RTHREADS, 1) tf32_allreduce_kernel(const __grid_constant__ Tf32Params p, const __grid_constant__ DistParams d) {
    constexpr int NCTA = CTA2 ? 2 : 1;
    constexpr int PM = TBM * NCTA;
    constexpr int VE = REAL ? 4 : 2;                 // elements per 16-byte vector
    constexpr int ESZ = REAL ? 4 : 8;
    constexpr int RPARTS = RCfg<MC>::PARTS, U = RCfg<MC>::U;
    constexpr int NVEC = TBM * BN / VE, PVEC = NVEC / RPARTS;
    static_assert(PVEC % RTHREADS == 0, "part size");
    __shared__ int64_t sRow[TBM];
    __shared__ int64_t sCol[BN];
    const int tid = threadIdx.x;
    const int64_t owned = (d.nunits - d.rank + d.nranks - 1) / d.nranks;   // units rank, rank + nranks, ...
    const int64_t items = owned > 0 ? owned * RPARTS : 0;
    if (d.timeline && tid == 0) atomicMin(&d.timeline[2], global_ns());
    for (int64_t j = blockIdx.x; j < items; j += gridDim.x) {
        const int64_t u = (j / RPARTS) * d.nranks + d.rank;
        const int part = (int)(j % RPARTS);
        const int64_t tile = u / NCTA;
        const int sub = (int)(u % NCTA);
        const TileCoord tc = tile_coord<BN, PM>(p, tile);
        const int64_t m0 = (int64_t)tc.m0 + sub * TBM;
        __syncthreads();                               // previous item's tables are no longer read
        if (tid < TBM) sRow[tid] = (m0 + tid < p.M) ? p.rowC[m0 + tid] + p.batC[tc.l] : -1;
        for (int i = tid; i < BN; i += RTHREADS) sCol[i] = (tc.n0 + i < p.N) ? p.colC[tc.n0 + i] : -1;
        if (tid < d.nranks) spin_until(d.flags[d.rank] + u * d.nranks + tid, d.epoch);
        __syncthreads();                               // all nranks partial sub-tiles of this unit are visible
        if (d.timeline && tid == 0 && j == blockIdx.x) atomicMin(&d.timeline[3], global_ns());
        const size_t ubase = (size_t)u * (TBM * BN) * ESZ;
        for (int v0 = part * PVEC + tid; v0 < (part + 1) * PVEC; v0 += RTHREADS * U) {
            float4 acc[U];
            if constexpr (MC) {
#pragma unroll
                for (int k = 0; k < U; k++) {
                    const int v = v0 + k * RTHREADS;
                    if (v < (part + 1) * PVEC) acc[k] = multimem_ld_reduce_f4(reinterpret_cast<const char *>(d.mc_ws) + ubase + (size_t)v * 16);
                }
            } else {
#pragma unroll
                for (int k = 0; k < U; k++) {
                    const int v = v0 + k * RTHREADS;
                    if (v < (part + 1) * PVEC) acc[k] = ld_relaxed_sys_f4(reinterpret_cast<const char *>(d.ws[0]) + ubase + (size_t)v * 16);
                }
                for (int s = 1; s < d.nranks; s++) {   // rank order: deterministic, identical on every rank
                    float4 x[U];
                    const char *w = reinterpret_cast<const char *>(d.ws[s]) + ubase;
#pragma unroll
                    for (int k = 0; k < U; k++) {
                        const int v = v0 + k * RTHREADS;
                        if (v < (part + 1) * PVEC) x[k] = ld_relaxed_sys_f4(w + (size_t)v * 16);
                    }
#pragma unroll
                    for (int k = 0; k < U; k++) { acc[k].x += x[k].x; acc[k].y += x[k].y; acc[k].z += x[k].z; acc[k].w += x[k].w; }
                }
            }
#pragma unroll
            for (int k = 0; k < U; k++) {
                const int v = v0 + k * RTHREADS;
                if (v >= (part + 1) * PVEC) continue;
                const int e = v * VE, r = e % TBM, c = e / TBM;
                const int64_t col = sCol[c];
                if (col < 0) continue;
                const int64_t r0 = sRow[r];
                bool contig = r0 >= 0 && ((r0 + col) % VE) == 0;
#pragma unroll
                for (int i = 1; i < VE; i++) contig = contig && sRow[r + i] == r0 + i;
                if (contig) {
                    const size_t co = (size_t)(r0 + col) * ESZ;
                    if constexpr (MC) multimem_st_f4(reinterpret_cast<char *>(d.mc_c) + co, acc[k]);
                    else
                        for (int s = 0; s < d.nranks; s++) *reinterpret_cast<float4 *>(reinterpret_cast<char *>(d.c[s]) + co) = acc[k];
                } else {
                    const float a4[4] = {acc[k].x, acc[k].y, acc[k].z, acc[k].w};
#pragma unroll
                    for (int i = 0; i < VE; i++) {
                        const int64_t ri = sRow[r + i];
                        if (ri < 0) continue;
                        const size_t co = (size_t)(ri + col) * ESZ;
                        if constexpr (REAL) {
                            if constexpr (MC) multimem_st_f1(reinterpret_cast<char *>(d.mc_c) + co, a4[i]);
                            else
                                for (int s = 0; s < d.nranks; s++) *reinterpret_cast<float *>(reinterpret_cast<char *>(d.c[s]) + co) = a4[i];
                        } else {
                            const float2 z = make_float2(a4[2 * i], a4[2 * i + 1]);
                            if constexpr (MC) multimem_st_f2(reinterpret_cast<char *>(d.mc_c) + co, z);
                            else
                                for (int s = 0; s < d.nranks; s++) *reinterpret_cast<float2 *>(reinterpret_cast<char *>(d.c[s]) + co) = z;
                        }
                    }
                }
            }
        }
    }
    // every store of this CTA is visible system-wide, then the last CTA to finish raises done[rank] on every rank
    __threadfence_system();
    __syncthreads();
    if (d.timeline && tid == 0) atomicMax(&d.timeline[4], global_ns());
    if (tid == 0) {
        int *counter = d.flags[d.rank] + d.nunits * d.nranks + d.nranks;
        const int prev = atomicAdd(counter, 1);
        if (prev == (int)gridDim.x - 1) {
            *counter = 0;                              // next call (stream-ordered after this kernel) starts from zero
            __threadfence_system();
            for (int s = 0; s < d.nranks; s++) st_release_sys(d.flags[s] + d.nunits * d.nranks + d.rank, d.epoch);
        }
    }
}

// a rank's C is complete when every owner has raised its done flag here
__global__ void dist_wait_done_kernel(const int *done, int nranks, int epoch, unsigned long long *timeline) {
    if ((int)threadIdx.x < nranks) spin_until(done + threadIdx.x, epoch);
    __syncthreads();
    if (timeline && threadIdx.x == 0) timeline[5] = global_ns();
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// packed operand: [L][rows][W*K] floats, K-major (W = 4 complex, 2 real), rows of KG*128 bytes
bool make_map(CUtensorMap *map, const void *base, int64_t K, int W, int64_t rows, int64_t L, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)(W * K), (cuuint64_t)rows, (cuuint64_t)L};
    cuuint64_t strides[2] = {(cuuint64_t)(4 * W * K), (cuuint64_t)(4 * W * K) * (cuuint64_t)rows};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// CTA-pair policy: MB200_CTA_PAIR=0 disables, 2 forces the pair kernel whenever the shape allows it (tests, A/B)
int pair_mode() {
    static const int m = [] { const char *e = getenv("MB200_CTA_PAIR"); return e ? atoi(e) : 1; }();
    return m;
}
bool pair_mode_allows() { return pair_mode() != 0; }

// dist mode decides the kernel variant from the shape alone (every rank, the GEMM and the reducer must agree)
template <int BN, bool REAL>
constexpr bool pair_ok() { return (REAL && BN == 256) || (!REAL && BN == 128); }
template <int BN, bool REAL>
bool dist_uses_pair(int64_t M) { return pair_ok<BN, REAL>() && pair_mode_allows() && M >= 2 * TBM; }

template <int BN, bool REAL>
cudaError_t launch_bn(const void *packA, const void *packB, const GettParams &g, bool mixed, cudaStream_t s, bool *pair,
                      const DistDesc *dist) {
    CUtensorMap mapA, mapB;
    constexpr int W = REAL ? 2 : 4;
    Tf32Params p{};
    if (dist) {
        p.dist_nranks = dist->nranks; p.dist_rank = dist->rank; p.dist_epoch = dist->epoch;
        for (int r = 0; r < dist->nranks; r++) p.dist_flags[r] = dist->flags[r];
        p.ws = dist->ws[dist->rank];
        p.timeline = dist->timeline;
        // default: one cumulative fence + release store by the signalling thread after the CTA barrier (what CUTLASS's stream-K
        // semaphore and a cooperative-groups grid sync rely on); MB200_DIST_FENCE=all adds a system fence per epilogue thread
        static const bool fence_all = [] { const char *e = getenv("MB200_DIST_FENCE"); return e && std::string(e) == "all"; }();
        p.dist_fence_all = fence_all ? 1 : 0;
    }
    p.sc = g.sc;
    p.C = g.C;
    p.rowC = g.rowC; p.colC = g.colC; p.batC = g.batC;
    p.M = g.M; p.N = g.N; p.L = g.L;
    p.mixed = mixed ? 1 : 0;
    p.KG = REAL ? (int)((g.K + 15) / 16) : (int)(g.K / 8);   // a half-filled last line reads zeros (TMA out-of-bounds fill)
    const int64_t ntiles = ((g.M + TBM - 1) / TBM) * ((g.N + BN - 1) / BN) * g.L;
    if (ntiles <= 0) return cudaSuccess;
    p.nsplit = 1;
    // CTA pairs (256 x BN tiles, cta_group::2) for the wide tile shapes once there is work for every SM
    constexpr bool PAIR_OK = (REAL && BN == 256) || (!REAL && BN == 128);
    if constexpr (PAIR_OK) {
        const int64_t ptiles = ((g.M + 2 * TBM - 1) / (2 * TBM)) * ((g.N + BN - 1) / BN) * g.L;
        const bool want = dist ? dist_uses_pair<BN, REAL>(g.M)
                               : (pair_mode() == 2 ? g.M > TBM : (pair_mode() == 1 && g.M >= 2 * TBM && ptiles >= 148));
        if (want) {
            if (!make_map(&mapA, packA, g.K, W, g.M, g.L, TBM) || !make_map(&mapB, packB, g.K, W, g.N, g.L, BN / 2))
                return cudaErrorInvalidValue;
            p.ntiles = ptiles;
            if (pair) *pair = true;
            cudaLaunchConfig_t cfg{};
            const int64_t tpcs = 74 - (dist ? (dist->reserve_sms + 1) / 2 : 0);   // dist mode: whole TPCs left to the reducer
            const int64_t pairs = ptiles < tpcs ? ptiles : tpcs;   // persistent: one CTA pair per TPC
            cfg.gridDim = dim3((unsigned)(2 * pairs));
            cfg.blockDim = dim3(dist ? TTHREADS + 32 : TTHREADS);
            cfg.dynamicSmemBytes = Tf32Smem<BN, REAL, true>::TOTAL;
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            const cudaError_t le = cudaLaunchKernelEx(&cfg, tf32_gemm_kernel<BN, REAL, true>, mapA, mapB, p);
            if (le == cudaSuccess || dist) return le;   // dist: the reducer decodes pair units, no silent change of variant
            // a device that cannot place the cluster (partitioned GPU, shared-memory carve-out): clear the launch-configuration
            // error and run the single-CTA kernel below
            if (le != cudaErrorInvalidConfiguration && le != cudaErrorLaunchOutOfResources && le != cudaErrorInvalidValue &&
                le != cudaErrorNotSupported)
                return le;
            (void)cudaGetLastError();
            if (pair) *pair = false;
        }
    }
    if (!make_map(&mapA, packA, g.K, W, g.M, g.L, TBM) || !make_map(&mapB, packB, g.K, W, g.N, g.L, BN))
        return cudaErrorInvalidValue;
    p.ntiles = ntiles;
    // split-K when the tiles cannot fill the SMs: at most one slice per 128-k TMEM chunk
    const int64_t chunks = (g.K + CHUNK_K - 1) / CHUNK_K;
    static const int sk_mode = [] { const char *e = getenv("MB200_SPLITK"); return e ? atoi(e) : 1; }();
    if (sk_mode && g.sc.nranks == 0 && !dist && ntiles * 4 <= 148 * 3 && chunks >= 2)
        p.nsplit = (int)std::min<int64_t>(chunks, std::max<int64_t>(1, 148 / ntiles));
    void *ws = nullptr;
    if (p.nsplit > 1) {   // (never in dist mode: p.ws is the rank's cross-GPU workspace there)
        cudaError_t e = cudaMallocAsync(&ws, (size_t)p.nsplit * ntiles * TBM * BN * (REAL ? 4 : 8), s);
        if (e != cudaSuccess) return e;
        p.ws = ws;
    }
    const int64_t nunits = ntiles * p.nsplit;
    const int64_t sms = 148 - (dist ? dist->reserve_sms : 0);
    const int64_t grid = nunits < sms ? nunits : sms;   // persistent: one CTA per SM (dist mode: minus the reducer's SMs)
    tf32_gemm_kernel<BN, REAL, false><<<(unsigned)grid, dist ? TTHREADS + 32 : TTHREADS, Tf32Smem<BN, REAL, false>::TOTAL, s>>>(mapA, mapB, p);
    if (p.nsplit > 1) {
        const int64_t threads = ntiles * TBM * BN;
        tf32_splitk_reduce_kernel<BN, REAL><<<(unsigned)std::min<int64_t>((threads + 255) / 256, 148 * 16), 256, 0, s>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (ws) cudaFreeAsync(ws, s);
    return e;
}

}  // namespace

bool tf32_available() { return encode_fn() != nullptr; }

// opt-in shared memory sizes; called once per device from mb200_create
cudaError_t tf32_configure() {
    cudaError_t e = cudaSuccess;
#define MB200_TCFG(BN, REAL, CTA2)                                                                                          \
    if (e == cudaSuccess)                                                                                                   \
        e = cudaFuncSetAttribute(tf32_gemm_kernel<BN, REAL, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Tf32Smem<BN, REAL, CTA2>::TOTAL)
    MB200_TCFG(128, false, false); MB200_TCFG(64, false, false); MB200_TCFG(256, true, false); MB200_TCFG(128, true, false);
    MB200_TCFG(128, false, true); MB200_TCFG(256, true, true);
#undef MB200_TCFG
    return e;
}

// C (scattered through rowC/colC/batC) = packA [L][M][W*K] x packB [L][N][W*K]^T, K % 8 == 0;
// dtype ComplexF32 (W = 4) or Float32 (W = 2)
cudaError_t launch_tf32_gemm(int dtype, const void *packA, const void *packB, const GettParams &g, bool mixed, cudaStream_t s, bool *pair,
                             const DistDesc *dist) {
    if (pair) *pair = false;
    if (g.K % 8 != 0 || g.K < 8) return cudaErrorInvalidValue;
    if (dtype == MB200_F32) {
        if (g.N > 128) return launch_bn<256, true>(packA, packB, g, mixed, s, pair, dist);
        return launch_bn<128, true>(packA, packB, g, mixed, s, pair, dist);
    }
    if (g.N > 64) return launch_bn<128, false>(packA, packB, g, mixed, s, pair, dist);
    return launch_bn<64, false>(packA, packB, g, mixed, s, pair, dist);
}

namespace {
template <int BN, bool REAL>
DistGeometry geometry_bn(int64_t M, int64_t N, int64_t L) {
    const bool pair = dist_uses_pair<BN, REAL>(M);
    const int64_t pm = pair ? 2 * TBM : TBM;
    const int64_t tiles = ((M + pm - 1) / pm) * ((N + BN - 1) / BN) * L;
    return DistGeometry{BN, pair ? 1 : 0, tiles * (pair ? 2 : 1), (int64_t)TBM * BN};
}

template <int BN, bool REAL>
cudaError_t launch_allreduce_bn(const GettParams &g, const DistDesc &dist, cudaStream_t s) {
    const DistGeometry geo = geometry_bn<BN, REAL>(g.M, g.N, g.L);
    Tf32Params p{};
    p.rowC = g.rowC; p.colC = g.colC; p.batC = g.batC;
    p.M = g.M; p.N = g.N; p.L = g.L;
    p.ntiles = geo.nunits / (geo.pair ? 2 : 1);
    DistParams d{};
    d.nranks = dist.nranks; d.rank = dist.rank; d.epoch = dist.epoch; d.nunits = geo.nunits;
    for (int r = 0; r < dist.nranks; r++) { d.ws[r] = dist.ws[r]; d.c[r] = dist.c[r]; d.flags[r] = dist.flags[r]; }
    d.mc_ws = dist.mc_ws; d.mc_c = dist.mc_c;
    d.timeline = dist.timeline;
    const int64_t owned = (geo.nunits - dist.rank + dist.nranks - 1) / dist.nranks;
    // a rank that owns nothing still launches one CTA: its done flag must go up
    // grid: the reserved SMs when the reducer runs next to the GEMM (overlap), every SM when it runs after it
    const int64_t want = dist.reserve_sms > 0 ? dist.reserve_sms : 148;
    const bool mc = d.mc_ws != nullptr;
    unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(owned * (mc ? RCfg<true>::PARTS : RCfg<false>::PARTS), want));
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(RTHREADS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    if constexpr (pair_ok<BN, REAL>()) {
        if (geo.pair) {
            if (dist.reserve_sms > 0 && grid >= 2) {   // whole TPCs, like the GEMM's CTA pairs
                grid &= ~1u;
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
            }
            cfg.gridDim = dim3(grid);
            return mc ? cudaLaunchKernelEx(&cfg, tf32_allreduce_kernel<BN, REAL, true, true>, p, d)
                      : cudaLaunchKernelEx(&cfg, tf32_allreduce_kernel<BN, REAL, true, false>, p, d);
        }
    }
    cfg.gridDim = dim3(grid);
    return mc ? cudaLaunchKernelEx(&cfg, tf32_allreduce_kernel<BN, REAL, false, true>, p, d)
              : cudaLaunchKernelEx(&cfg, tf32_allreduce_kernel<BN, REAL, false, false>, p, d);
}
}  // namespace

DistGeometry tf32_dist_geometry(int dtype, int64_t M, int64_t N, int64_t L) {
    if (dtype == MB200_F32) return N > 128 ? geometry_bn<256, true>(M, N, L) : geometry_bn<128, true>(M, N, L);
    return N > 64 ? geometry_bn<128, false>(M, N, L) : geometry_bn<64, false>(M, N, L);
}

cudaError_t launch_tf32_allreduce(int dtype, const GettParams &g, const DistDesc &dist, cudaStream_t s) {
    if (dtype == MB200_F32) return g.N > 128 ? launch_allreduce_bn<256, true>(g, dist, s) : launch_allreduce_bn<128, true>(g, dist, s);
    return g.N > 64 ? launch_allreduce_bn<128, false>(g, dist, s) : launch_allreduce_bn<64, false>(g, dist, s);
}

cudaError_t launch_dist_wait_done(const DistDesc &dist, int64_t nunits, cudaStream_t s) {
    dist_wait_done_kernel<<<1, 32, 0, s>>>(dist.flags[dist.rank] + nunits * dist.nranks, dist.nranks, dist.epoch, dist.timeline);
    return cudaGetLastError();
}

}  // namespace mb200
