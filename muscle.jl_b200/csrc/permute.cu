// K1 — N-d permute / matricise (Julia `permutedims`, src/Tensor.jl:302-308 → Base.permutedims), with
// an optional complex interleaved → planar split in the same pass. HBM-bound: every element is
// read once and written once, both sides coalesced through a shared-memory tile.
//
// Canonical problem (built by the caller): modes in SOURCE order (mode 0 = source unit stride,
// extent-1 modes dropped, modes adjacent in both layouts merged) + each mode's DESTINATION stride.
// Let j be the mode with destination stride 1.
//   * A tile is a product of three index ranges, all with power-of-two tile extents
//       v : a chunk of mode 0 when it is the unit-stride mode of BOTH layouts (j == 0), else absent
//       x : a chunk of the "row space"    = leading source modes (contiguous in the source)
//       y : a chunk of the "column space" = leading destination modes (contiguous in the destination)
//     Every other mode is an outer mode decoded from blockIdx.
//   * Reads walk (v,x) fastest → source-contiguous runs; writes walk (v,y) fastest → destination-
//     contiguous runs. Offsets are separable (off = v + tabX[x] + tabY[y]), so each tile builds four
//     small shared tables once; the element loops are shift/mask + two table reads + one
//     LDG/STS or LDS/STG — no divisions.
//   * smem pitch per y is padded so the transposed read is bank-conflict free for 4/8/16 B elements.
//   * When the tile needs no transposition (no x or no y range) elements go straight from global to
//     global in one loop, no shared memory.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#include "kernels.cuh"

namespace mb200 {

namespace {

constexpr int PT_THREADS = 256;
constexpr int PT_MAX_TAB = 128;   // entries per table of the persistent kernel (TX, TY <= 128)
constexpr int PT_BIG_TAB = 256;   // entries per table of the one-tile-per-CTA kernels (TX, TY <= 256)

struct PermK {
    int n_out;
    unsigned o_ext[MB200_MAX_MODES];
    int64_t o_ss[MB200_MAX_MODES], o_ds[MB200_MAX_MODES];
    int64_t v_ext;
    int nx, ny;
    unsigned x_ext[MB200_MAX_MODES], y_ext[MB200_MAX_MODES];
    int64_t x_ss[MB200_MAX_MODES], x_ds[MB200_MAX_MODES];
    int64_t y_ss[MB200_MAX_MODES], y_ds[MB200_MAX_MODES];
    int64_t SX, SY;
    int lv, lx, ly;          // log2 of the tile extents VT, TX, TY
    unsigned v_chunks, x_chunks, y_chunks;
    int pitch;               // smem elements per y
    int direct;              // 1: no transposition needed, global → global
    int64_t plane_stride;    // planar: im plane offset in scalars; split writers: operand format / side (see put)
};

// digits of idx over `ext` (32-bit divisions), accumulated into 64-bit offsets
__device__ __forceinline__ void digits_off(unsigned idx, int n, const unsigned *ext, const int64_t *ss,
                                           const int64_t *ds, int64_t &so, int64_t &dof) {
    so = 0; dof = 0;
    for (int i = 0; i < n; i++) {
        unsigned e = ext[i], q = idx / e, d = idx - q * e;
        idx = q;
        so += (int64_t)d * ss[i];
        dof += (int64_t)d * ds[i];
    }
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// round-to-nearest tf32 (x - tf32_rn(x) is exact in fp32 and half the size of the truncation remainder); a finite x within
// 2^-12 of FLT_MAX would round up to infinity: truncate those
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    if ((r & 0x7F800000u) == 0x7F800000u) r = __float_as_uint(x) & 0xFFFFE000u;   // inf / nan stay what they were
    return __uint_as_float(r);
}
// one 32-bit word holding two bf16 values: `lo` in bits [0,16) (the even k position of a K-major bf16 operand), `hi` in [16,32).
// Round to nearest, except that a finite value may not round up to infinity (inf * 0 in the cross terms would poison the sum).
__device__ __forceinline__ float bf16_pair(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    if ((r & 0x7F800000u) == 0x7F800000u || (r & 0x00007F80u) == 0x00007F80u)
        asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));      // rz keeps inf / nan inputs as they are
    return __uint_as_float(r);
}
// The "x" chunk of the mixed TF32 + BF16 operand format (tf32.cu): for the row operand (side 1) the pair (bf16(x), bf16(x_lo)),
// for the column operand (side 2) the pair (bf16(x_lo), bf16(x)), so that one K = 16 bf16 MMA over a chunk pair
// yields sum_k x*y_lo + x_lo*y — both cross terms of the split product.
__device__ __forceinline__ float cross_word(float x, float lo, int side) {
    return side == 1 ? bf16_pair(x, lo) : bf16_pair(lo, x);
}

// PLANAR: 0 = interleaved (plain), 1 = two planes, 2 = ComplexF32 hi/lo split into four 8-word chunks,
//         3 = Float32 hi/lo split into two 8-word chunks.
// aux: PLANAR 1 -> offset of the imaginary plane (scalars); PLANAR 2/3 -> 0: 3xTF32 format (hi = truncated tf32, second chunk = the
//      fp32 remainder), 1 / 2: mixed TF32 + BF16 format of the row / column operand (hi = rounded tf32, second chunk = cross_word)
template <typename E, typename S, int PLANAR>
__device__ __forceinline__ void put(void *dstv, int64_t off, const E &val, int64_t aux) {
    if constexpr (PLANAR == 1) {
        S *d = reinterpret_cast<S *>(dstv);
        d[off] = val.x;
        d[off + aux] = val.y;
    } else if constexpr (PLANAR == 2) {
        S *d = reinterpret_cast<S *>(dstv);
        if (aux == 0) {
            const float rh = tf32_hi(val.x), ih = tf32_hi(val.y);
            d[off] = rh;
            d[off + 8] = val.x - rh;
            d[off + 16] = ih;
            d[off + 24] = val.y - ih;
        } else {
            const float rh = tf32_rn(val.x), ih = tf32_rn(val.y);
            d[off] = rh;
            d[off + 8] = cross_word(val.x, val.x - rh, (int)aux);
            d[off + 16] = ih;
            d[off + 24] = cross_word(val.y, val.y - ih, (int)aux);
        }
    } else if constexpr (PLANAR == 3) {
        S *d = reinterpret_cast<S *>(dstv);
        if (aux == 0) {
            const float h = tf32_hi(val);
            d[off] = h;
            d[off + 8] = val - h;
        } else {
            const float h = tf32_rn(val);
            d[off] = h;
            d[off + 8] = cross_word(val, val - h, (int)aux);
        }
    } else {
        reinterpret_cast<E *>(dstv)[off] = val;
    }
}

// per-tile scalars (uniform across the CTA)
struct TileCtx {
    int64_t base_s, base_d;
    int vt, tx, ty;
    unsigned x0, y0;
};

__device__ __forceinline__ TileCtx decode_tile(const PermK &p, unsigned b) {
    TileCtx t;
    const unsigned vc = b % p.v_chunks; b /= p.v_chunks;
    const unsigned xc = b % p.x_chunks; b /= p.x_chunks;
    const unsigned yc = b % p.y_chunks; b /= p.y_chunks;
    digits_off(b, p.n_out, p.o_ext, p.o_ss, p.o_ds, t.base_s, t.base_d);
    const int64_t v0 = (int64_t)vc << p.lv, x0 = (int64_t)xc << p.lx, y0 = (int64_t)yc << p.ly;
    t.vt = (int)min((int64_t)1 << p.lv, p.v_ext - v0);
    t.tx = (int)min((int64_t)1 << p.lx, p.SX - x0);
    t.ty = (int)min((int64_t)1 << p.ly, p.SY - y0);
    t.base_s += v0; t.base_d += v0;  // v has stride 1 on both sides
    t.x0 = (unsigned)x0; t.y0 = (unsigned)y0;
    return t;
}

template <int N> struct TileTabsT { int64_t srcX[N], dstX[N], srcY[N], dstY[N]; };
using TileTabs = TileTabsT<PT_MAX_TAB>;
using TileTabsBig = TileTabsT<PT_BIG_TAB>;

template <class Tabs>
__device__ __forceinline__ void build_tabs(const PermK &p, const TileCtx &t, Tabs &tab, int tid) {
    for (int i = tid; i < t.tx; i += PT_THREADS) {
        int64_t so, dof;
        digits_off(t.x0 + i, p.nx, p.x_ext, p.x_ss, p.x_ds, so, dof);
        tab.srcX[i] = so; tab.dstX[i] = dof;
    }
    for (int i = tid; i < t.ty; i += PT_THREADS) {
        int64_t so, dof;
        digits_off(t.y0 + i, p.ny, p.y_ext, p.y_ss, p.y_ds, so, dof);
        tab.srcY[i] = so; tab.dstY[i] = dof;
    }
}

// No transposition needed (no x range or no y range): read order == write order, global -> global.
template <typename E, typename S, int PLANAR>
__global__ void __launch_bounds__(PT_THREADS) permute_direct_kernel(const __grid_constant__ PermK p,
                                                                    const E *__restrict__ src, void *__restrict__ dstv) {
    __shared__ TileTabsBig tab;
    const int tid = threadIdx.x;
    const TileCtx t = decode_tile(p, blockIdx.x);
    build_tabs(p, t, tab, tid);
    __syncthreads();
    const int total = 1 << (p.lv + p.lx + p.ly);
    const int mv = (1 << p.lv) - 1, mx = (1 << p.lx) - 1;
#pragma unroll 4
    for (int idx = tid; idx < total; idx += PT_THREADS) {
        const int v = idx & mv, x = (idx >> p.lv) & mx, y = idx >> (p.lv + p.lx);
        if (v < t.vt && x < t.tx && y < t.ty) {
            E val = src[t.base_s + v + tab.srcX[x] + tab.srcY[y]];
            put<E, S, PLANAR>(dstv, t.base_d + v + tab.dstX[x] + tab.dstY[y], val, p.plane_stride);
        }
    }
}

// Transposing tile, one tile per CTA (partial tiles, 4-byte and Float64 elements).
template <typename E, typename S, int PLANAR>
__global__ void __launch_bounds__(PT_THREADS) permute_simple_kernel(const __grid_constant__ PermK p,
                                                                    const E *__restrict__ src, void *__restrict__ dstv) {
    __shared__ TileTabsBig tab;
    extern __shared__ __align__(16) unsigned char tile_raw[];
    E *tile = reinterpret_cast<E *>(tile_raw);
    const int tid = threadIdx.x;
    const TileCtx t = decode_tile(p, blockIdx.x);
    build_tabs(p, t, tab, tid);
    __syncthreads();
    const int total = 1 << (p.lv + p.lx + p.ly);
    const int lvx = p.lv + p.lx, mv = (1 << p.lv) - 1, mvx = (1 << lvx) - 1, my = (1 << p.ly) - 1;
#pragma unroll 4
    for (int idx = tid; idx < total; idx += PT_THREADS) {   // gather: (v,x) fastest
        const int vx = idx & mvx, v = idx & mv, x = vx >> p.lv, y = idx >> lvx;
        if (v < t.vt && x < t.tx && y < t.ty) tile[y * p.pitch + vx] = src[t.base_s + v + tab.srcX[x] + tab.srcY[y]];
    }
    __syncthreads();
#pragma unroll 4
    for (int idx = tid; idx < total; idx += PT_THREADS) {   // scatter: (v,y) fastest
        const int v = idx & mv, y = (idx >> p.lv) & my, x = idx >> (p.lv + p.ly);
        if (v < t.vt && x < t.tx && y < t.ty) {
            E val = tile[y * p.pitch + (x << p.lv) + v];
            put<E, S, PLANAR>(dstv, t.base_d + v + tab.dstX[x] + tab.dstY[y], val, p.plane_stride);
        }
    }
}

// Transposing tiles. Persistent CTAs walk tiles blockIdx.x, +gridDim.x, ...; the next tile's elements are
// already in flight into registers while the current tile is scattered from shared memory (two tile buffers,
// three table sets, one __syncthreads per tile).
template <typename E, typename S, int PLANAR, int NE>
__global__ void __launch_bounds__(PT_THREADS) permute_tiled_kernel(const __grid_constant__ PermK p, unsigned ntiles,
                                                                   const E *__restrict__ src, void *__restrict__ dstv) {
    __shared__ TileTabs tabs[3];
    extern __shared__ __align__(16) unsigned char tile_raw[];
    E *tile0 = reinterpret_cast<E *>(tile_raw);
    const int tile_stride = p.pitch << p.ly;   // elements per tile buffer
    const int tid = threadIdx.x;
    const int lvx = p.lv + p.lx, mv = (1 << p.lv) - 1, mvx = (1 << lvx) - 1, my = (1 << p.ly) - 1;
    const int total = 1 << (lvx + p.ly);

    // With power-of-two tile extents and 256 threads, a thread's (v,x) slot on the gather side and its (v,y) slot
    // on the scatter side do not change from element to element (idx = tid + 256*i): only y (resp. x) advances by a
    // fixed step. That removes most of the per-element index arithmetic and half of the table reads.
    const int lvy = p.lv + p.ly;
    const bool hoist = lvx <= 8 && lvy <= 8;
    const int g_vx = tid & mvx, g_v = tid & mv, g_x = g_vx >> p.lv, g_y0 = tid >> lvx, g_ystep = PT_THREADS >> lvx;
    const int s_v = tid & mv, s_y = (tid >> p.lv) & my, s_x0 = tid >> lvy, s_xstep = PT_THREADS >> lvy;

    E regs[NE];
    auto gather = [&](const TileCtx &t, const TileTabs &tab) {
        if (hoist) {
            const bool ok = g_v < t.vt && g_x < t.tx;
            const E *base = src + (t.base_s + g_v + tab.srcX[g_x]);
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const int y = g_y0 + i * g_ystep;
                if (ok && y < t.ty) regs[i] = base[tab.srcY[y]];
            }
        } else {
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const int idx = tid + i * PT_THREADS;
                const int vx = idx & mvx, v = idx & mv, x = vx >> p.lv, y = idx >> lvx;
                if (v < t.vt && x < t.tx && y < t.ty) regs[i] = src[t.base_s + v + tab.srcX[x] + tab.srcY[y]];
            }
        }
    };

    unsigned cur_id = blockIdx.x;
    if (cur_id >= ntiles) return;
    TileCtx cur = decode_tile(p, cur_id);
    build_tabs(p, cur, tabs[0], tid);
    __syncthreads();
    gather(cur, tabs[0]);
    for (int it = 0;; it++) {
        E *tile = tile0 + (size_t)(it & 1) * tile_stride;
        // registers -> smem slot y*pitch + (x*VT + v)   (conflict-free: consecutive lanes, consecutive slots)
        if (hoist) {
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const int y = g_y0 + i * g_ystep;
                if (y <= my) tile[y * p.pitch + g_vx] = regs[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const int idx = tid + i * PT_THREADS;
                if (idx < total) tile[(idx >> lvx) * p.pitch + (idx & mvx)] = regs[i];   // small tiles use fewer slots
            }
        }
        const unsigned next_id = cur_id + gridDim.x;
        const bool has_next = next_id < ntiles;
        TileCtx nxt = cur;
        if (has_next) {
            nxt = decode_tile(p, next_id);
            build_tabs(p, nxt, tabs[(it + 1) % 3], tid);
        }
        __syncthreads();
        if (has_next) gather(nxt, tabs[(it + 1) % 3]);          // in flight during the scatter below
        const TileTabs &tab = tabs[it % 3];
        if (hoist) {                                            // scatter: (v,y) fastest
            const bool ok = s_v < cur.vt && s_y < cur.ty;
            const int64_t dbase = cur.base_d + s_v + tab.dstY[s_y];
            const E *trow = tile + s_y * p.pitch + s_v;
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const int x = s_x0 + i * s_xstep;
                if (ok && x < cur.tx) put<E, S, PLANAR>(dstv, dbase + tab.dstX[x], trow[x << p.lv], p.plane_stride);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NE; i++) {
                const int idx = tid + i * PT_THREADS;
                const int v = idx & mv, y = (idx >> p.lv) & my, x = idx >> (p.lv + p.ly);
                if (v < cur.vt && x < cur.tx && y < cur.ty) {
                    E val = tile[y * p.pitch + (x << p.lv) + v];
                    put<E, S, PLANAR>(dstv, cur.base_d + v + tab.dstX[x] + tab.dstY[y], val, p.plane_stride);
                }
            }
        }
        if (!has_next) break;
        cur = nxt;
        cur_id = next_id;
    }
}


// Register-tile transposition (no shared memory for data): used when the source and destination unit-stride modes
// differ (pure transposition); x = source-leading modes, y = destination-leading modes (each side may skip modes the
// other took), addressed through per-tile tables, VEC-blocks contiguous along x in the source and along y in the destination.
// A thread owns VEC x VEC blocks (VEC = elements per 16 bytes): it reads VEC source rows (y) with one 16-byte load
// each, transposes in registers, and writes VEC destination rows (x) with one 16-byte store each. Lanes are laid
// 2^lxl (x) by 2^(5-lxl) (y) blocks, so neighbouring lanes touch neighbouring 16-byte pieces and every request covers
// whole 32-byte sectors on both sides (measured best: 4 x 8 lanes = 64 B source pieces, 128 B destination pieces,
// 64 KB tiles: 8192^2 Float32 transpose 1.79 -> 5.5 TB/s, ComplexF32 3.9 -> 5.5, d=2-fastest ComplexF64 3.1 -> 5.2). UNR blocks per thread are loaded before any is stored (128 B in flight per
// thread). 2 memory instructions + 2 table reads per 16 bytes moved, against 4 + 2 per ELEMENT through the smem tile.
// ESZ = element bytes (4, 8, 16); data is moved as raw bits.
constexpr int PT_REG_TAB = 1024;

template <int ESZ>
__global__ void __launch_bounds__(PT_THREADS) permute_regT_kernel(const __grid_constant__ PermK p, int lxl_target,
                                                                  const unsigned char *__restrict__ src,
                                                                  unsigned char *__restrict__ dst) {
    constexpr int VEC = 16 / ESZ;                 // block edge
    constexpr int LB = VEC == 4 ? 2 : (VEC == 2 ? 1 : 0);
    constexpr int UNR = ESZ / 2;                  // blocks in flight per thread: 2 / 4 / 8 -> 8 x 16 B
    __shared__ int64_t srcY[PT_REG_TAB], dstX[PT_REG_TAB];     // per element
    __shared__ int64_t srcX[PT_REG_TAB], dstY[PT_REG_TAB];     // per VEC-block
    const int tid = threadIdx.x;
    const TileCtx t = decode_tile(p, blockIdx.x);
    for (int i = tid; i < t.ty; i += PT_THREADS) {
        int64_t so, dof;
        digits_off(t.y0 + i, p.ny, p.y_ext, p.y_ss, p.y_ds, so, dof);
        srcY[i] = so * ESZ;
        if ((i & (VEC - 1)) == 0) dstY[i >> LB] = dof * ESZ;
    }
    for (int i = tid; i < t.tx; i += PT_THREADS) {
        int64_t so, dof;
        digits_off(t.x0 + i, p.nx, p.x_ext, p.x_ss, p.x_ds, so, dof);
        dstX[i] = dof * ESZ;
        if ((i & (VEC - 1)) == 0) srcX[i >> LB] = so * ESZ;
    }
    __syncthreads();
    const unsigned char *sbase = src + t.base_s * ESZ;
    unsigned char *dbase = dst + t.base_d * ESZ;
    const int lbx = p.lx - LB, lby = p.ly - LB;
    const int lyl0 = min(5 - min(lxl_target, lbx), lby);
    const int lxl = min(5 - lyl0, lbx), lyl = lyl0, lxh = lbx - lxl;
    const int nblk = 1 << (lbx + lby);
    const int mxl = (1 << lxl) - 1, myl = (1 << lyl) - 1, mxh = (1 << lxh) - 1;
    for (int b0 = tid; b0 < nblk; b0 += PT_THREADS * UNR) {
        uint4 r[UNR][VEC];
        int xs[UNR], ys[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int b = b0 + u * PT_THREADS;
            const int xb = (b & mxl) | (((b >> (lxl + lyl)) & mxh) << lxl);
            const int yb = ((b >> lxl) & myl) | ((b >> (lxl + lyl + lxh)) << lyl);
            const int x = xb << LB, y = yb << LB;
            const bool ok = b < nblk && x < t.tx && y < t.ty;
            xs[u] = ok ? x : -1;
            ys[u] = y;
            if (ok) {
#pragma unroll
                for (int j = 0; j < VEC; j++)
                    r[u][j] = __ldg(reinterpret_cast<const uint4 *>(sbase + srcX[x >> LB] + srcY[y + j]));
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int x = xs[u], y = ys[u];
            if (x < 0) continue;
            uint4 w[VEC];
            if constexpr (ESZ == 4) {
                w[0] = make_uint4(r[u][0].x, r[u][1].x, r[u][2].x, r[u][3].x);
                w[1] = make_uint4(r[u][0].y, r[u][1].y, r[u][2].y, r[u][3].y);
                w[2] = make_uint4(r[u][0].z, r[u][1].z, r[u][2].z, r[u][3].z);
                w[3] = make_uint4(r[u][0].w, r[u][1].w, r[u][2].w, r[u][3].w);
            } else if constexpr (ESZ == 8) {
                w[0] = make_uint4(r[u][0].x, r[u][0].y, r[u][1].x, r[u][1].y);
                w[1] = make_uint4(r[u][0].z, r[u][0].w, r[u][1].z, r[u][1].w);
            } else {
                w[0] = r[u][0];
            }
#pragma unroll
            for (int i = 0; i < VEC; i++)
                *reinterpret_cast<uint4 *>(dbase + dstY[y >> LB] + dstX[x + i]) = w[i];
        }
    }
}

// Table-driven pack into the tcgen05 operand format (tf32.cu) for operands the strided-permutation pack cannot express: summed
// extents that do not tile groups of 8 k (K = 100, bond dimensions 3, 5, 6, 10 ...; K is zero-padded to Kp = 8 * ceil(K / 8)) and
// strided (non-dense) operands. dst[l][row][group(k)] <- src[row_tab[row] + k_tab[k] + bat_tab[l]] through a 32 x 32 tile:
// the gather runs with lanes along k (K-major source) or along rows, the writes always with lanes along k (32-byte runs).
struct PackGather {
    const int64_t *row_tab, *k_tab, *bat_tab;
    int64_t rows, K, Kp, L;
    int kmajor, aux;   // aux: split writer format / side (see put)
    // Row visiting order of the line-writer pack (kernels.cuh PackRowOrder): the rows of a tile are 64 consecutive values of n', whose
    // digits run over the row modes in SOURCE-stride order; n = sum digit_i * weight_i is the GEMM's row index. nd == 0: n' = n.
    int nd;
    int64_t d_ext[MB200_PACK_DIGITS], d_w[MB200_PACK_DIGITS];
};
template <bool REAL>
__global__ void __launch_bounds__(256) pack_gather_kernel(const __grid_constant__ PackGather q, const void *__restrict__ srcv,
                                                          float *__restrict__ dst) {
    using E = typename std::conditional<REAL, float, float2>::type;
    constexpr int W = REAL ? 2 : 4;
    __shared__ E tile[32][33];
    const E *src = reinterpret_cast<const E *>(srcv);
    const int64_t tiles_k = (q.Kp + 31) / 32, tiles_r = (q.rows + 31) / 32;
    int64_t b = blockIdx.x;
    const int64_t k0 = (b % tiles_k) * 32; b /= tiles_k;
    const int64_t r0 = (b % tiles_r) * 32;
    const int64_t l = b / tiles_r;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t bat = q.bat_tab[l];
    if (q.kmajor) {
        const int64_t k = k0 + tx;
        const bool kok = k < q.K;
        const int64_t koff = kok ? q.k_tab[k] : 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = ty + 8 * i;
            E v{};
            if (kok && r0 + r < q.rows) v = src[q.row_tab[r0 + r] + koff + bat];
            tile[r][tx] = v;
        }
    } else {
        const bool rok = r0 + tx < q.rows;
        const int64_t roff = rok ? q.row_tab[r0 + tx] : 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int kk = ty + 8 * i;
            E v{};
            if (rok && k0 + kk < q.K) v = src[roff + q.k_tab[k0 + kk] + bat];
            tile[tx][kk] = v;
        }
    }
    __syncthreads();
    const int64_t k = k0 + tx;
    if (k >= q.Kp) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = ty + 8 * i;
        const int64_t row = r0 + r;
        if (row < q.rows)
            put<E, float, REAL ? 3 : 2>(dst, (l * q.rows + row) * (W * q.Kp) + (k >> 3) * (8 * W) + (k & 7), tile[r][tx], q.aux);
    }
}

// Line-writer pack into the tcgen05 operand format (tf32.cu): any layout (table-driven like pack_gather_kernel), but the destination
// is written as whole 128-byte lines with 16-byte stores - a warp store instruction covers 512 contiguous bytes - instead of four
// scattered 4-byte stores per element (ncu on the 4-byte writers: 35-56 % of HBM peak on 24 bytes moved per ComplexF32 element).
//   tile = 64 rows x 32 k. Gather: lanes along k (K-major source) or along rows, 8 elements in flight per thread, into a
//   shared-memory tile [row][k]. Write: 16-byte piece p of the line of (row, 8-k group): chunk type p >> 1 (re_hi, re_x, im_hi,
//   im_x - or hi, x for Float32), 4 consecutive k each; the four lanes that need the same 4 elements read them as a broadcast.
// Algorithmic bytes per element: sizeof(T) read + 2 sizeof(T) written (the hi and x planes) - declared in DESIGN 3.5.
// split of 4 values into their hi words (tf32) and x words (fp32 remainder, or the bf16 cross-term pair of the mixed scheme).
// AUX 0: 3xTF32 (hi = truncated tf32, x = remainder), 1 / 2: TF32 + BF16, row / column operand. The fast path is 3 instructions per
// value (cvt.rna.tf32, FADD, cvt.rn.bf16x2): rounding can reach infinity only from an exponent of 254 or 255, so ONE test of the
// four input magnitudes replaces the per-conversion overflow guards; the rare slow path is the guarded scalar code of put().
template <int AUX>
__device__ __forceinline__ void split4(const float (&in)[4], float4 &hi, float4 &xw) {
    float h[4], x[4];
    if constexpr (AUX == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) { h[i] = tf32_hi(in[i]); x[i] = in[i] - h[i]; }
    } else {
        const uint32_t m01 = max(__float_as_uint(in[0]) & 0x7FFFFFFFu, __float_as_uint(in[1]) & 0x7FFFFFFFu);
        const uint32_t m23 = max(__float_as_uint(in[2]) & 0x7FFFFFFFu, __float_as_uint(in[3]) & 0x7FFFFFFFu);
        if (max(m01, m23) < 0x7F000000u) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t r, w;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(in[i]));
                h[i] = __uint_as_float(r);
                const float lo = in[i] - h[i];
                if constexpr (AUX == 1) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(lo), "f"(in[i]));      // (x, x_lo): x in the low half
                else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(in[i]), "f"(lo));                          // (x_lo, x)
                x[i] = __uint_as_float(w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) { h[i] = tf32_rn(in[i]); x[i] = cross_word(in[i], in[i] - h[i], AUX); }
        }
    }
    hi = make_float4(h[0], h[1], h[2], h[3]);
    xw = make_float4(x[0], x[1], x[2], x[3]);
}

template <bool REAL, int AUX>
__global__ void __launch_bounds__(256) pack_lines_kernel(const __grid_constant__ PackGather q, const void *__restrict__ srcv,
                                                         float *__restrict__ dst) {
    using E = typename std::conditional<REAL, float, float2>::type;
    constexpr int W = REAL ? 2 : 4;
    constexpr int TR = 64, TK = 32;
    constexpr int PITCH = TK + (REAL ? 4 : 2);            // keeps every row 16-byte aligned
    __shared__ __align__(16) E tile[TR][PITCH];
    __shared__ int64_t sRow[TR], sK[TK], sDst[TR];
    const E *src = reinterpret_cast<const E *>(srcv);
    const int tid = threadIdx.x;
    const int64_t tiles_k = (q.Kp + TK - 1) / TK, tiles_r = (q.rows + TR - 1) / TR;
    int64_t b = blockIdx.x;
    const int64_t k0 = (b % tiles_k) * TK; b /= tiles_k;
    const int64_t r0 = (b % tiles_r) * TR;
    const int64_t l = b / tiles_r;
    if (tid < TR) {
        int64_t n = r0 + tid;
        const bool ok = n < q.rows;
        if (ok && q.nd) {   // n' -> n: digits of n' in source-stride order, recombined with their weights in the GEMM's row order
            int64_t rem = n;
            n = 0;
            for (int i = 0; i < q.nd; i++) {
                const int64_t e = q.d_ext[i], dgt = rem % e;
                rem /= e;
                n += dgt * q.d_w[i];
            }
        }
        sRow[tid] = ok ? q.row_tab[n] : -1;
        sDst[tid] = ok ? n : -1;
    } else if (tid < TR + TK) sK[tid - TR] = (k0 + tid - TR < q.K) ? q.k_tab[k0 + tid - TR] : -1;
    const E *srcb = src + q.bat_tab[l];
    __syncthreads();
    E v[8];
    if (q.kmajor) {
        const int kk = tid & 31, rb = tid >> 5;
        const int64_t ko = sK[kk];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int64_t ro = sRow[rb + 8 * i];
            v[i] = E{};
            if (ko >= 0 && ro >= 0) v[i] = srcb[ro + ko];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) tile[rb + 8 * i][kk] = v[i];
    } else {
        const int rr = tid & 63, kb = tid >> 6;
        const int64_t ro = sRow[rr];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int64_t ko = sK[kb + 4 * i];
            v[i] = E{};
            if (ko >= 0 && ro >= 0) v[i] = srcb[ro + ko];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) tile[rr][kb + 4 * i] = v[i];
    }
    __syncthreads();
    // Write phase. A work item = 4 consecutive k of one component (re / im; Float32: the value) of one row: it produces the matching
    // 16-byte pieces of the hi chunk and of the x chunk. Complex: 4 items per (row, 8-k group) -> lanes 4j..4j+3 cover one 128-byte
    // line; the 32 lanes of a warp cover two rows' worth of 4 groups... item index = tid + 256 * i.
    constexpr int ITEMS_PER_GROUP = REAL ? 2 : 4;                       // (half) x (re, im)
    constexpr int NITEMS = TR * (TK / 8) * ITEMS_PER_GROUP;             // 512 (real) / 1024 (complex)
    float *dl = dst + l * q.rows * (W * q.Kp);
#pragma unroll
    for (int i = 0; i < NITEMS / 256; i++) {
        const int it = tid + 256 * i;
        const int sub = it % ITEMS_PER_GROUP, g = (it / ITEMS_PER_GROUP) & 3, r = it / (ITEMS_PER_GROUP * 4);
        const int64_t row = sDst[r], kg = k0 + g * 8;
        if (row < 0 || kg >= q.Kp) continue;
        const int half = sub & 1, comp = sub >> 1;                      // comp: 0 = re (or the real value), 1 = im
        const E *e = &tile[r][g * 8 + half * 4];
        float in[4];
        if constexpr (REAL) {
            const float4 t = *reinterpret_cast<const float4 *>(e);
            in[0] = t.x; in[1] = t.y; in[2] = t.z; in[3] = t.w;
        } else {
            const float4 t0 = *reinterpret_cast<const float4 *>(e), t1 = *reinterpret_cast<const float4 *>(e + 2);
            if (comp == 0) { in[0] = t0.x; in[1] = t0.z; in[2] = t1.x; in[3] = t1.z; }
            else { in[0] = t0.y; in[1] = t0.w; in[2] = t1.y; in[3] = t1.w; }
        }
        float4 hi, xw;
        split4<AUX>(in, hi, xw);
        float *d = dl + row * (W * q.Kp) + (kg >> 3) * (8 * W) + comp * 16 + half * 4;   // chunk order: hi | x | (im) hi | x
        *reinterpret_cast<float4 *>(d) = hi;
        *reinterpret_cast<float4 *>(d + 8) = xw;
    }
}

// ---- TMA-staged transposition (north star: "tiled through shared memory or staged via TMA") ---------------------------------
// Pure transpositions (source unit-stride mode x != destination unit-stride mode y) of 4 / 8 / 16-byte elements with at most
// 5 canonical modes: both tensors are described by 5-D tiled tensor maps over 32-bit words (an element = ESZ / 4 words folded into
// the unit-stride dimension of EACH map, so 16-byte elements need no extra dimension). A persistent CTA walks 16 KB tiles through
// a 3-stage ring: one thread issues cp.async.bulk.tensor loads (box [TX][TY], source order: x fastest) that complete on per-stage
// mbarriers, all 8 warps transpose the tile in shared memory (diagonal lane mapping: element (x = lane, y = (lane + i) % 32) of a
// 32 x 32 sub-tile per step - conflict-free on both sides without padding, which a dense TMA box would not allow), and one thread
// issues the cp.async.bulk.tensor store (box [TY][TX], destination order: y fastest) as a bulk group. Ragged edges cost nothing:
// the load zero-fills out-of-bounds elements, the store clips them. No global load / store instruction, no offset tables.
__device__ __forceinline__ uint32_t pt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pt_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pt_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pt_mbar_wait(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(pt_smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) __trap();   // a descriptor bug must not hang the GPU
    }
}
__device__ __forceinline__ void pt_tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(pt_smem_u32(dst)), "l"(map), "r"(pt_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void pt_tma_store_5d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(pt_smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

struct TmaPermK {
    unsigned x_chunks, y_chunks;
    unsigned o_ext[3];     // outer modes (map dimensions 2..4), extent 1 when absent
    unsigned ntiles;
};

template <int ESZ, int TX, int TY>
struct TmaPermCfg {
    static constexpr int STAGES = 3;
    static constexpr int TILE_BYTES = TX * TY * ESZ;
    static constexpr int SMEM = 2 * STAGES * TILE_BYTES + 128 /*align*/ + 64 /*barriers*/;
};

template <int ESZ, int TX, int TY>
__global__ void __launch_bounds__(PT_THREADS) permute_tma_kernel(const __grid_constant__ CUtensorMap mapS,
                                                                 const __grid_constant__ CUtensorMap mapD,
                                                                 const __grid_constant__ TmaPermK k) {
    using Cfg = TmaPermCfg<ESZ, TX, TY>;
    using E = typename std::conditional<ESZ == 16, uint4, typename std::conditional<ESZ == 8, uint2, uint32_t>::type>::type;
    constexpr int STAGES = Cfg::STAGES, W = ESZ / 4;
    extern __shared__ unsigned char tma_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tma_raw) + 127) & ~(uintptr_t)127);
    E *tin = reinterpret_cast<E *>(base);                                     // [STAGES][TY][TX]
    E *tout = reinterpret_cast<E *>(base + STAGES * Cfg::TILE_BYTES);         // [STAGES][TX][TY]
    uint64_t *full = reinterpret_cast<uint64_t *>(base + 2 * STAGES * Cfg::TILE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapS) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapD) : "memory");
        for (int s = 0; s < STAGES; s++) pt_mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (blockIdx.x >= k.ntiles) return;
    const unsigned n_my = (k.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto coords = [&](unsigned it, int (&c)[5]) {   // tile -> (x chunk, y chunk, outer digits)
        unsigned t = blockIdx.x + it * gridDim.x;
        c[0] = (int)(t % k.x_chunks) * TX; t /= k.x_chunks;
        c[1] = (int)(t % k.y_chunks) * TY; t /= k.y_chunks;
        c[2] = (int)(t % k.o_ext[0]); t /= k.o_ext[0];
        c[3] = (int)(t % k.o_ext[1]); t /= k.o_ext[1];
        c[4] = (int)t;
    };
    auto issue_load = [&](unsigned it) {
        int c[5];
        coords(it, c);
        const int s = it % STAGES;
        pt_mbar_expect_tx(&full[s], Cfg::TILE_BYTES);
        pt_tma_load_5d(tin + (size_t)s * TX * TY, &mapS, &full[s], c[0] * W, c[1], c[2], c[3], c[4]);
    };
    if (tid == 0)
        for (unsigned it = 0; it < n_my && it < (unsigned)STAGES; it++) issue_load(it);
    for (unsigned it = 0; it < n_my; it++) {
        const int s = it % STAGES;
        pt_mbar_wait(&full[s], (it / STAGES) & 1);
        // the bulk store that read out[s] three tiles ago has finished reading shared memory
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
        __syncthreads();
        const E *in = tin + (size_t)s * TX * TY;
        E *out = tout + (size_t)s * TX * TY;
        constexpr int NSUB = (TX / 32) * (TY / 32);
#pragma unroll 4
        for (int item = warp; item < NSUB * 32; item += PT_THREADS / 32) {
            const int sub = item >> 5, i = item & 31;
            const int x = (sub % (TX / 32)) * 32 + lane, y = (sub / (TX / 32)) * 32 + ((lane + i) & 31);
            out[x * TY + y] = in[y * TX + x];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA store
        __syncthreads();
        if (tid == 0) {
            int c[5];
            coords(it, c);
            pt_tma_store_5d(&mapD, out, c[1] * W, c[0], c[2], c[3], c[4]);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (it + STAGES < n_my) issue_load(it + STAGES);            // in[s] has been read by every thread (barrier above)
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays valid until the stores have read it
}

inline int floor_log2(int64_t v) { int l = 0; while ((int64_t)2 << l <= v) l++; return l; }
inline int ceil_log2(int64_t v) { int l = 0; while (((int64_t)1 << l) < v) l++; return l; }

}  // namespace

namespace {

// Tile selection: assigns every canonical mode a role (outer / v / x / y) and picks power-of-two tile extents.
// `cap` = log2 of the largest x / y tile extent (table size of the kernel that will run).
// Returns false when the problem does not fit the 32-bit digit math.
bool build_tile(const PermuteParams &q, size_t esz, int cap, int ltile_bump, PermK &k, std::vector<int> &dord) {
    const int n = q.n;
    const int G = (int)(128 / esz);                                   // elements per 128 B smem phase
    const int ltile = (esz == 16 ? 11 : (esz == 8 ? 12 : 13)) + ltile_bump;   // 32 KB tiles
    dord.resize(n);
    for (int i = 0; i < n; i++) dord[i] = i;
    std::stable_sort(dord.begin(), dord.end(), [&](int a, int b) { return q.dst_stride[a] < q.dst_stride[b]; });
    std::vector<int64_t> sstride(n);
    {
        int64_t st = 1;
        for (int i = 0; i < n; i++) { sstride[i] = st; st *= q.ext[i]; }
    }
    k = PermK{};
    k.plane_stride = q.split ? q.split - 1 : q.plane_stride;
    std::vector<int> role(n, 0);  // 0 outer, 1 v, 2 x, 3 y
    k.v_ext = 1;
    int xs = 0;
    size_t ys = 0;
    int lv = 0;
    if (n > 0 && dord[0] == 0) {  // shared unit-stride mode
        role[0] = 1;
        k.v_ext = q.ext[0];
        // short runs: one chunk covers the whole mode; long runs: power-of-two chunks up to the tile
        lv = q.ext[0] <= 16 ? ceil_log2(q.ext[0]) : std::min(floor_log2(q.ext[0]), ltile);
        xs = 1; ys = 1;
    }
    const int lrem = ltile - lv;
    // if v-runs are already >= 128 B, no transposition is needed: spend the rest of the tile on y only
    const bool long_runs = ((int64_t)1 << lv) * (int64_t)esz >= 128;
    int lx_t = long_runs ? 0 : std::min(lrem / 2, cap);
    int ly_t = std::min(lrem - lx_t, cap);
    int64_t SX = 1, SY = 1;
    // the destination unit-stride mode (after v) always belongs to y, so x cannot steal it
    if (ys < (size_t)n && role[dord[ys]] == 0 && ly_t > 0) { role[dord[ys]] = 3; SY *= q.ext[dord[ys]]; ys++; }
    // The smem-tile kernels need x / y to be LEADING runs of the source / destination order; the register-tile kernel
    // (cap 10) addresses both sides through tables, so a side may skip modes the other side already took: for an
    // interleaving permutation (dest order m4 m0 m5 m1 ...) the tile becomes {m0 m1} x {m4 m5} and writes 4 KB runs
    // (m4 m0 m5) instead of 512 B ones.
    const bool skip_taken = cap > 8;
    auto grow_x = [&](int lg) {
        while (SX < ((int64_t)1 << lg) && xs < n) {
            if (role[xs] == 0) { role[xs] = 2; SX *= q.ext[xs]; }
            else if (!skip_taken) break;
            xs++;
        }
    };
    auto grow_y = [&](int lg) {
        while (SY < ((int64_t)1 << lg) && ys < (size_t)n) {
            if (role[dord[ys]] == 0) { role[dord[ys]] = 3; SY *= q.ext[dord[ys]]; }
            else if (!skip_taken) break;
            ys++;
        }
    };
    // x to its target, then y with whatever x could not use, then x again with whatever y could not use
    grow_x(lx_t);
    grow_y(std::min(lrem - std::min(lx_t, ceil_log2(SX)), cap));
    grow_x(std::min(lrem - std::min(cap, ceil_log2(SY)), cap));
    for (int i = 0; i < n; i++) {
        if (role[i] == 2) { k.x_ext[k.nx] = (unsigned)q.ext[i]; k.x_ss[k.nx] = sstride[i]; k.x_ds[k.nx] = q.dst_stride[i]; k.nx++; }
        if (role[i] == 0) { k.o_ext[k.n_out] = (unsigned)q.ext[i]; k.o_ss[k.n_out] = sstride[i]; k.o_ds[k.n_out] = q.dst_stride[i]; k.n_out++; }
    }
    for (int i : dord)
        if (role[i] == 3) { k.y_ext[k.ny] = (unsigned)q.ext[i]; k.y_ss[k.ny] = sstride[i]; k.y_ds[k.ny] = q.dst_stride[i]; k.ny++; }
    k.SX = SX; k.SY = SY;
    // tile extents (powers of two): shrink to the space, give the slack to the other side
    int lx = std::min(lx_t, ceil_log2(SX));
    int ly = std::min(std::min(lrem - lx, cap), ceil_log2(SY));
    lx = std::min(std::min(lrem - ly, cap), ceil_log2(SX));
    k.lv = lv; k.lx = lx; k.ly = ly;
    const int VT = 1 << lv, TX = 1 << lx, TY = 1 << ly;
    k.direct = (lx == 0 || ly == 0) ? 1 : 0;
    const int row = VT * TX;
    k.pitch = ((row + G - 1) / G) * G + std::min(VT, G);
    k.v_chunks = (unsigned)((k.v_ext + VT - 1) / VT);
    k.x_chunks = (unsigned)((SX + TX - 1) / TX);
    k.y_chunks = (unsigned)((SY + TY - 1) / TY);
    int64_t outer = 1;
    for (int i = 0; i < k.n_out; i++) outer *= k.o_ext[i];
    if (SX >= ((int64_t)1 << 31) || SY >= ((int64_t)1 << 31) || outer >= ((int64_t)1 << 31)) return false;
    return true;
}

// Register-tile kernel eligibility for a tile built with the wide tables.
bool regT_ok(const PermK &k, size_t esz, const void *src, const void *dst) {
    if (k.v_ext != 1 || k.direct || k.nx == 0 || k.ny == 0) return false;
    if ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) return false;
    const int vec = (int)(16 / esz), lb = vec == 4 ? 2 : (vec == 2 ? 1 : 0);
    // a VEC-block along x must be contiguous in the source, one along y contiguous in the destination
    if (k.x_ss[0] != 1 || k.x_ext[0] % vec || k.y_ds[0] != 1 || k.y_ext[0] % vec) return false;
    if (k.lx < lb || k.ly < lb) return false;
    bool ok = true;   // every other stride keeps 16-byte alignment
    for (int i = 0; i < k.n_out; i++) ok = ok && k.o_ss[i] % vec == 0 && k.o_ds[i] % vec == 0;
    for (int i = 0; i < k.ny; i++) ok = ok && k.y_ss[i] % vec == 0 && (i == 0 || k.y_ds[i] % vec == 0);
    for (int i = 0; i < k.nx; i++) ok = ok && k.x_ds[i] % vec == 0 && (i == 0 || k.x_ss[i] % vec == 0);
    return ok;
}

}  // namespace

namespace {

typedef CUresult (*PtEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PtEncodeTiledFn pt_encode_fn() {
    static PtEncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PtEncodeTiledFn>(p);
    });
    return fn;
}

// 0: not eligible (caller falls back to the register-tile / smem-tile kernels), 1: launched, -1: launch error in *err
template <int ESZ, int TX, int TY>
int launch_tma_cfg(const PermuteParams &q, int j, const void *src, void *dst, cudaStream_t s, cudaError_t *err) {
    PtEncodeTiledFn fn = pt_encode_fn();
    if (!fn) return 0;
    constexpr int W = ESZ / 4;
    int64_t sstride[MB200_MAX_MODES];
    {
        int64_t st = 1;
        for (int i = 0; i < q.n; i++) { sstride[i] = st; st *= q.ext[i]; }
    }
    // map dimension order: (x = mode 0, y = mode j, others) for the source, (y, x, others) for the destination
    int others[3], no = 0;
    for (int i = 1; i < q.n; i++)
        if (i != j) others[no++] = i;
    cuuint64_t dimS[5], dimD[5], strS[4], strD[4];
    cuuint32_t boxS[5] = {TX * W, TY, 1, 1, 1}, boxD[5] = {TY * W, TX, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    dimS[0] = (cuuint64_t)q.ext[0] * W; dimS[1] = (cuuint64_t)q.ext[j];
    dimD[0] = (cuuint64_t)q.ext[j] * W; dimD[1] = (cuuint64_t)q.ext[0];
    strS[0] = (cuuint64_t)sstride[j] * ESZ;          // byte stride of map dimension 1
    strD[0] = (cuuint64_t)q.dst_stride[0] * ESZ;
    TmaPermK k{};
    for (int d = 0; d < 3; d++) {
        const bool have = d < no;
        const int m = have ? others[d] : 0;
        dimS[2 + d] = dimD[2 + d] = have ? (cuuint64_t)q.ext[m] : 1;
        strS[1 + d] = have ? (cuuint64_t)sstride[m] * ESZ : (cuuint64_t)q.total * ESZ;
        strD[1 + d] = have ? (cuuint64_t)q.dst_stride[m] * ESZ : (cuuint64_t)q.total * ESZ;
        k.o_ext[d] = have ? (unsigned)q.ext[m] : 1u;
    }
    for (int d = 0; d < 4; d++)
        if (strS[d] % 16 || strD[d] % 16 || strS[d] >= ((cuuint64_t)1 << 40) || strD[d] >= ((cuuint64_t)1 << 40)) return 0;
    for (int d = 0; d < 5; d++)
        if (dimS[d] >= ((cuuint64_t)1 << 32) || dimD[d] >= ((cuuint64_t)1 << 32)) return 0;
    CUtensorMap mapS, mapD;
    if (fn(&mapS, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, const_cast<void *>(src), dimS, strS, boxS, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    if (fn(&mapD, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, dst, dimD, strD, boxD, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    k.x_chunks = (unsigned)((q.ext[0] + TX - 1) / TX);
    k.y_chunks = (unsigned)((q.ext[j] + TY - 1) / TY);
    const int64_t ntiles = (int64_t)k.x_chunks * k.y_chunks * k.o_ext[0] * k.o_ext[1] * k.o_ext[2];
    if (ntiles <= 0 || ntiles > 0x7fffffffLL) return 0;
    k.ntiles = (unsigned)ntiles;
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, 148 * 2);
    permute_tma_kernel<ESZ, TX, TY><<<grid, PT_THREADS, TmaPermCfg<ESZ, TX, TY>::SMEM, s>>>(mapS, mapD, k);
    *err = cudaGetLastError();
    return *err == cudaSuccess ? 1 : -1;
}

int try_permute_tma(const PermuteParams &q, size_t esz, const void *src, void *dst, cudaStream_t s, cudaError_t *err) {
    if (q.n < 2 || q.n > 5 || q.dst_stride[0] == 1) return 0;
    if ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) return 0;
    int j = -1;
    for (int i = 1; i < q.n; i++)
        if (q.dst_stride[i] == 1) j = i;
    if (j < 0 || q.ext[0] < 32 || q.ext[j] < 32) return 0;      // short runs: the register-tile kernel wastes less
    if (esz == 16) return launch_tma_cfg<16, 32, 32>(q, j, src, dst, s, err);
    if (esz == 8) return launch_tma_cfg<8, 64, 32>(q, j, src, dst, s, err);
    if (esz == 4) return launch_tma_cfg<4, 64, 64>(q, j, src, dst, s, err);
    return 0;
}

}  // namespace

cudaError_t launch_permute(int dtype, const PermuteParams &q_in, const void *src, void *dst, cudaStream_t s) {
    if (q_in.total <= 0) return cudaSuccess;
    PermuteParams q = q_in;
    size_t esz = dtype_size(dtype);
    const bool planar = q.plane_stride != 0;
    const bool plain = !planar && !q.split;
    for (int i = 0; i < q.n; i++)
        if (q.ext[i] >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;  // 32-bit digit math

    // Plain permutations move raw bits, so a shared unit-stride mode can be folded into wider elements:
    // (4 x float) -> one 16-byte element etc. Everything then runs on the 16-byte kernels.
    if (plain && q.n > 0) {
        while (esz < 16 && q.dst_stride[0] == 1 && q.ext[0] % 2 == 0 &&
               ((((uintptr_t)src) | ((uintptr_t)dst)) & (2 * esz - 1)) == 0) {
            bool even = true;
            for (int i = 1; i < q.n; i++) even = even && q.dst_stride[i] % 2 == 0;
            if (!even) break;
            esz *= 2;
            q.ext[0] /= 2;
            q.total /= 2;
            for (int i = 1; i < q.n; i++) q.dst_stride[i] /= 2;
        }
        if (q.ext[0] == 1 && q.n > 1) {   // the whole mode became one element
            for (int i = 1; i < q.n; i++) { q.ext[i - 1] = q.ext[i]; q.dst_stride[i - 1] = q.dst_stride[i]; }
            q.n--;
        }
    }
    const int ltile = esz == 16 ? 11 : (esz == 8 ? 12 : 13);

    // MB200_PERMUTE_TMA: 0 = never, 1 = whenever eligible (pure transposition, <= 5 modes, runs >= 32 elements on both sides)
    static const int tma_mode = [] { const char *e = getenv("MB200_PERMUTE_TMA"); return e ? atoi(e) : 0; }();
    if (plain && q.tma >= 0 && (tma_mode || q.tma > 0)) {
        cudaError_t terr = cudaSuccess;
        const int r = try_permute_tma(q, esz, src, dst, s, &terr);
        if (r == 1) return cudaSuccess;
        if (r < 0) return terr;
    }

    PermK k;
    std::vector<int> dord;
    bool reg = false;
    static const int regt_mask = [] { const char *e = getenv("MB200_PERMUTE_REGT"); return e ? atoi(e) : 4 | 8 | 16; }();
    static const int regt_bump = [] { const char *e = getenv("MB200_REGT_BUMP"); return e ? atoi(e) : 1; }();
    static const int regt_lxl = [] { const char *e = getenv("MB200_REGT_LXL"); return e ? atoi(e) : 2; }();
    if (plain && q.n >= 2 && q.dst_stride[0] != 1 && (regt_mask & (int)esz)) {   // pure transposition: try the register-tile kernel
        if (!build_tile(q, esz, 10, regt_bump, k, dord)) return cudaErrorInvalidValue;
        reg = regT_ok(k, esz, src, dst);
    }
    if (!reg && !build_tile(q, esz, 8, 0, k, dord)) return cudaErrorInvalidValue;
    const int lv = k.lv, lx = k.lx, ly = k.ly;
    int64_t outer = 1;
    for (int i = 0; i < k.n_out; i++) outer *= k.o_ext[i];
    int64_t grid = (int64_t)k.v_chunks * k.x_chunks * k.y_chunks * outer;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    size_t smem = k.direct ? 0 : ((size_t)k.pitch << ly) * esz;   // one tile buffer
    const unsigned ntiles = (unsigned)grid;

    if (reg) {
        const unsigned char *sp = (const unsigned char *)src;
        unsigned char *dp = (unsigned char *)dst;
        if (esz == 4) permute_regT_kernel<4><<<ntiles, PT_THREADS, 0, s>>>(k, regt_lxl, sp, dp);
        else if (esz == 8) permute_regT_kernel<8><<<ntiles, PT_THREADS, 0, s>>>(k, regt_lxl, sp, dp);
        else permute_regT_kernel<16><<<ntiles, PT_THREADS, 0, s>>>(k, regt_lxl, sp, dp);
        return cudaGetLastError();
    }

    // persistent grid for the transposing kernel: 2 CTAs per SM (two 33 KB tile buffers + 3 table sets each)
    const unsigned pgrid = std::min<unsigned>(ntiles, 148u * 2u);
    // the persistent kernel takes full-size tiles of 8- and 16-byte elements; everything else one tile per CTA
    const bool full_tile = (lv + lx + ly) == ltile && lx <= 7 && ly <= 7;
#define MB200_PERM(E, S, P)                                                                                   \
    do {                                                                                                      \
        constexpr int NE = (int)((32 * 1024 / sizeof(E)) / PT_THREADS);                                       \
        constexpr bool PERSIST_OK = sizeof(E) == 16 || (sizeof(E) == 8 && (P) != 0) || std::is_same<E, float2>::value || (P) == 3; \
        if (k.direct) {                                                                                       \
            permute_direct_kernel<E, S, P><<<ntiles, PT_THREADS, 0, s>>>(k, (const E *)src, dst);             \
        } else if (PERSIST_OK && full_tile) {                                                                 \
            permute_tiled_kernel<E, S, P, NE><<<pgrid, PT_THREADS, 2 * smem, s>>>(k, ntiles, (const E *)src, dst); \
        } else {                                                                                              \
            permute_simple_kernel<E, S, P><<<ntiles, PT_THREADS, smem, s>>>(k, (const E *)src, dst);          \
        }                                                                                                     \
    } while (0)
    if (plain) {   // raw bits: the container type only fixes the element size
        if (esz == 4) MB200_PERM(float, float, 0);
        else if (esz == 8) MB200_PERM(float2, float, 0);
        else MB200_PERM(double2, double, 0);
    } else if (dtype == MB200_C64) {
        if (q.split) MB200_PERM(float2, float, 2);
        else MB200_PERM(float2, float, 1);
    } else if (dtype == MB200_C128) {
        MB200_PERM(double2, double, 1);
    } else if (dtype == MB200_F32 && q.split) {
        MB200_PERM(float, float, 3);
    } else {
        return cudaErrorInvalidValue;   // planar / split need a complex dtype
    }
#undef MB200_PERM
    return cudaGetLastError();
}

cudaError_t launch_pack_gather(int dtype, const void *src, const int64_t *row_tab, const int64_t *k_tab, const int64_t *bat_tab,
                               int64_t rows, int64_t K, int64_t Kp, int64_t L, int kmajor, int split, float *dst, cudaStream_t s,
                               const PackRowOrder *order) {
    if (rows <= 0 || L <= 0 || Kp <= 0) return cudaSuccess;
    if (split < 1 || (dtype != MB200_C64 && dtype != MB200_F32)) return cudaErrorInvalidValue;
    PackGather q{};
    q.row_tab = row_tab; q.k_tab = k_tab; q.bat_tab = bat_tab;
    q.rows = rows; q.K = K; q.Kp = Kp; q.L = L; q.kmajor = kmajor; q.aux = split - 1;
    if (order && order->nd > 0 && !kmajor) {
        q.nd = order->nd;
        for (int i = 0; i < order->nd; i++) { q.d_ext[i] = order->ext[i]; q.d_w[i] = order->weight[i]; }
    }
    const int64_t grid = ((Kp + 31) / 32) * ((rows + 31) / 32) * L;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    static const int lines = [] { const char *e = getenv("MB200_PACK_LINES"); return e ? atoi(e) : 1; }();
    if (lines && (((uintptr_t)dst) & 15) == 0) {   // whole 128-byte lines, 16-byte stores (MB200_PACK_LINES=0: the 4-byte writer, A/B)
        const int64_t g2 = ((Kp + 31) / 32) * ((rows + 63) / 64) * L;
        const bool real = dtype == MB200_F32;
        switch (q.aux) {
            case 0: if (real) pack_lines_kernel<true, 0><<<(unsigned)g2, 256, 0, s>>>(q, src, dst);
                    else pack_lines_kernel<false, 0><<<(unsigned)g2, 256, 0, s>>>(q, src, dst); break;
            case 1: if (real) pack_lines_kernel<true, 1><<<(unsigned)g2, 256, 0, s>>>(q, src, dst);
                    else pack_lines_kernel<false, 1><<<(unsigned)g2, 256, 0, s>>>(q, src, dst); break;
            default: if (real) pack_lines_kernel<true, 2><<<(unsigned)g2, 256, 0, s>>>(q, src, dst);
                     else pack_lines_kernel<false, 2><<<(unsigned)g2, 256, 0, s>>>(q, src, dst); break;
        }
        return cudaGetLastError();
    }
    if (dtype == MB200_F32) pack_gather_kernel<true><<<(unsigned)grid, 256, 0, s>>>(q, src, dst);
    else pack_gather_kernel<false><<<(unsigned)grid, 256, 0, s>>>(q, src, dst);
    return cudaGetLastError();
}

// opt-in shared memory sizes of every permute instantiation; called once per device from mb200_create
cudaError_t permute_configure() {
    cudaError_t e = cudaSuccess;
#define MB200_PCFG(E, S, P)                                                                                              \
    do {                                                                                                                 \
        constexpr int NE = (int)((32 * 1024 / sizeof(E)) / PT_THREADS);                                                  \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(permute_tiled_kernel<E, S, P, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(permute_simple_kernel<E, S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);    \
    } while (0)
    MB200_PCFG(float, float, 0); MB200_PCFG(float, float, 3);
    MB200_PCFG(float2, float, 0); MB200_PCFG(float2, float, 1); MB200_PCFG(float2, float, 2);
    MB200_PCFG(double2, double, 0); MB200_PCFG(double2, double, 1);
#undef MB200_PCFG
#define MB200_TCFG(ESZ, TX, TY)                                                                                                  \
    if (e == cudaSuccess)                                                                                                        \
        e = cudaFuncSetAttribute(permute_tma_kernel<ESZ, TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, TmaPermCfg<ESZ, TX, TY>::SMEM)
    MB200_TCFG(16, 32, 32); MB200_TCFG(8, 64, 32); MB200_TCFG(4, 64, 64);
#undef MB200_TCFG
    return e;
}

}  // namespace mb200
