// K1 — N-d permute / matricise (Julia `permutedims`, src/Tensor.jl:302-308 → Base.permutedims), with
// an optional complex interleaved → planar split in the same pass. HBM-bound: every element is
// read once and written once, both sides coalesced through a shared-memory tile.
//
// Canonical problem (built by the caller): modes in SOURCE order (mode 0 = source unit stride,
// extent-1 modes dropped, modes adjacent in both layouts merged) + each mode's DESTINATION stride.
// Let j be the mode with destination stride 1.
//   * A tile is a product of three index ranges
//       v : a chunk of mode 0 when it is the unit-stride mode of BOTH layouts (j == 0), else absent
//       x : a chunk of the "row space"    = leading source modes (contiguous in the source)
//       y : a chunk of the "column space" = leading destination modes (contiguous in the destination)
//     Every other mode is an outer mode decoded from blockIdx.
//   * Reads walk (v,x) fastest → source-contiguous runs; writes walk (v,y) fastest → destination-
//     contiguous runs. Offsets are separable, so each tile builds four small shared tables once
//     (source/dest offset and smem slot per (v,x) and per (v,y)); the element loops are then
//     shift/mask + two table reads + one LDG/STS or LDS/STG — no divisions.
//   * smem pitch is odd, so the transposed tile read is bank-conflict free for 4/8/16 B elements.
#include <algorithm>
#include <vector>

#include "kernels.cuh"

namespace mb200 {

namespace {

constexpr int PT_THREADS = 256;
constexpr int PT_MAX_TAB = 512;   // entries per table

struct PermK {
    // outer modes
    int n_out;
    int64_t o_ext[MB200_MAX_MODES], o_ss[MB200_MAX_MODES], o_ds[MB200_MAX_MODES];
    // v (mode 0 chunk; v_ext == 1 means absent)
    int64_t v_ext; int VT;
    // x / y spaces
    int nx, ny;
    int64_t x_ext[MB200_MAX_MODES], x_ss[MB200_MAX_MODES], x_ds[MB200_MAX_MODES];
    int64_t y_ext[MB200_MAX_MODES], y_ss[MB200_MAX_MODES], y_ds[MB200_MAX_MODES];
    int64_t SX, SY; int TX, TY;
    int64_t v_chunks, x_chunks, y_chunks;
    int pitch;              // smem elements per y (>= VT*TX, odd)
    int64_t plane_stride;   // planar: im plane offset in scalars
};

__device__ __forceinline__ void digits_off(int64_t idx, int n, const int64_t *ext, const int64_t *ss,
                                           const int64_t *ds, int64_t &so, int64_t &dof) {
    so = 0; dof = 0;
    for (int i = 0; i < n; i++) {
        int64_t e = ext[i], d = idx % e;
        idx /= e;
        so += d * ss[i];
        dof += d * ds[i];
    }
}

__device__ __forceinline__ int ceil_log2(int v) { return v <= 1 ? 0 : 32 - __clz(v - 1); }

template <typename E, typename S, bool PLANAR>
__global__ void __launch_bounds__(PT_THREADS) permute_kernel(const __grid_constant__ PermK p,
                                                             const E *__restrict__ src, void *__restrict__ dstv) {
    // read-side tables indexed by rx = v + vt*x ; write-side tables indexed by cy = v + vt*y
    __shared__ int64_t sSrcRX[PT_MAX_TAB], sDstCY[PT_MAX_TAB];
    __shared__ int64_t sSrcY[PT_MAX_TAB], sDstX[PT_MAX_TAB];
    __shared__ int sPosRX[PT_MAX_TAB], sPosCY[PT_MAX_TAB];
    extern __shared__ __align__(16) unsigned char tile_raw[];
    E *tile = reinterpret_cast<E *>(tile_raw);

    int64_t b = blockIdx.x;
    const int64_t vc = b % p.v_chunks; b /= p.v_chunks;
    const int64_t xc = b % p.x_chunks; b /= p.x_chunks;
    const int64_t yc = b % p.y_chunks; b /= p.y_chunks;
    int64_t base_s, base_d;
    digits_off(b, p.n_out, p.o_ext, p.o_ss, p.o_ds, base_s, base_d);
    const int64_t v0 = vc * p.VT, x0 = xc * p.TX, y0 = yc * p.TY;
    const int vt = (int)min((int64_t)p.VT, p.v_ext - v0);
    const int tx = (int)min((int64_t)p.TX, p.SX - x0);
    const int ty = (int)min((int64_t)p.TY, p.SY - y0);
    base_s += v0; base_d += v0;  // v has stride 1 on both sides
    const int tid = threadIdx.x;

    for (int i = tid; i < tx; i += PT_THREADS) {
        int64_t so, dof;
        digits_off(x0 + i, p.nx, p.x_ext, p.x_ss, p.x_ds, so, dof);
        sDstX[i] = dof;
        for (int v = 0; v < vt; v++) { sSrcRX[v + vt * i] = so + v; sPosRX[v + vt * i] = i * p.VT + v; }
    }
    for (int i = tid; i < ty; i += PT_THREADS) {
        int64_t so, dof;
        digits_off(y0 + i, p.ny, p.y_ext, p.y_ss, p.y_ds, so, dof);
        sSrcY[i] = so;
        for (int v = 0; v < vt; v++) { sDstCY[v + vt * i] = dof + v; sPosCY[v + vt * i] = i * p.pitch + v; }
    }
    __syncthreads();

    {   // gather: (v,x) fastest
        const int inner = vt * tx, sh = ceil_log2(inner), mask = (1 << sh) - 1;
        const int total = ty << sh;
#pragma unroll 4
        for (int idx = tid; idx < total; idx += PT_THREADS) {
            int rx = idx & mask, y = idx >> sh;
            if (rx < inner) tile[y * p.pitch + sPosRX[rx]] = src[base_s + sSrcRX[rx] + sSrcY[y]];
        }
    }
    __syncthreads();
    {   // scatter: (v,y) fastest
        const int inner = vt * ty, sh = ceil_log2(inner), mask = (1 << sh) - 1;
        const int total = tx << sh;
        S *dst = reinterpret_cast<S *>(dstv);
#pragma unroll 4
        for (int idx = tid; idx < total; idx += PT_THREADS) {
            int cy = idx & mask, x = idx >> sh;
            if (cy < inner) {
                E val = tile[sPosCY[cy] + x * p.VT];
                int64_t off = base_d + sDstCY[cy] + sDstX[x];
                if constexpr (PLANAR) {
                    dst[off] = val.x;
                    dst[off + p.plane_stride] = val.y;
                } else {
                    reinterpret_cast<E *>(dst)[off] = val;
                }
            }
        }
    }
}

}  // namespace

cudaError_t launch_permute(int dtype, const PermuteParams &q, const void *src, void *dst, cudaStream_t s) {
    if (q.total <= 0) return cudaSuccess;
    const int n = q.n;
    const size_t esz = dtype_size(dtype);
    const int tile_elems = esz == 16 ? 1024 : (esz == 8 ? 2048 : 4096);

    // destination order of the source modes
    std::vector<int> dord(n);
    for (int i = 0; i < n; i++) dord[i] = i;
    std::stable_sort(dord.begin(), dord.end(), [&](int a, int b) { return q.dst_stride[a] < q.dst_stride[b]; });
    std::vector<int64_t> sstride(n);
    {
        int64_t st = 1;
        for (int i = 0; i < n; i++) { sstride[i] = st; st *= q.ext[i]; }
    }

    PermK k{};
    k.plane_stride = q.plane_stride;
    std::vector<int> role(n, 0);  // 0 outer, 1 v, 2 x, 3 y
    k.v_ext = 1; k.VT = 1;
    int xs = 0;       // next source-order candidate for x
    size_t ys = 0;    // next destination-order candidate for y
    if (n > 0 && dord[0] == 0) {  // shared unit-stride mode
        role[0] = 1;
        k.v_ext = q.ext[0];
        k.VT = (int)std::min<int64_t>(q.ext[0], 128);
        xs = 1; ys = 1;
    }
    // target tile shape
    int rem = std::max(1, tile_elems / k.VT);
    int tx_target, ty_target;
    if (k.VT >= 64) { tx_target = 1; ty_target = rem; }
    else {
        int r = 1;
        while (r * r * 2 <= rem) r *= 2;       // ~sqrt, power of two
        tx_target = r; ty_target = rem / r;
        if (esz == 16 && k.VT == 1) { tx_target = 32; ty_target = 32; }
    }
    int64_t SX = 1, SY = 1;
    // the destination unit-stride mode always goes to y first (if it is not v)
    if (ys < (size_t)n && role[dord[ys]] == 0 && ty_target > 1) { role[dord[ys]] = 3; SY *= q.ext[dord[ys]]; ys++; }
    bool grow_x = tx_target > 1, grow_y = ty_target > 1;
    while (grow_x || grow_y) {
        if (grow_x) {
            if (SX >= tx_target || xs >= n || role[xs] != 0) grow_x = false;
            else { role[xs] = 2; SX *= q.ext[xs]; xs++; }
        }
        if (grow_y) {
            if (SY >= ty_target || ys >= (size_t)n || role[dord[ys]] != 0) grow_y = false;
            else { role[dord[ys]] = 3; SY *= q.ext[dord[ys]]; ys++; }
        }
    }
    for (int i = 0; i < n; i++) {
        if (role[i] == 2) { k.x_ext[k.nx] = q.ext[i]; k.x_ss[k.nx] = sstride[i]; k.x_ds[k.nx] = q.dst_stride[i]; k.nx++; }
        if (role[i] == 0) { k.o_ext[k.n_out] = q.ext[i]; k.o_ss[k.n_out] = sstride[i]; k.o_ds[k.n_out] = q.dst_stride[i]; k.n_out++; }
    }
    for (int i : dord)
        if (role[i] == 3) { k.y_ext[k.ny] = q.ext[i]; k.y_ss[k.ny] = sstride[i]; k.y_ds[k.ny] = q.dst_stride[i]; k.ny++; }
    k.SX = SX; k.SY = SY;
    // tile extents: use what the other side leaves unused, bounded by the table size
    k.TX = (int)std::min<int64_t>(SX, std::max(1, tx_target));
    k.TY = (int)std::min<int64_t>(SY, std::max(1, rem / k.TX));
    if ((int64_t)k.TX * k.TY * k.VT < rem * k.VT && SX > k.TX) k.TX = (int)std::min<int64_t>(SX, std::max(1, rem / k.TY));
    while (k.VT * k.TX > PT_MAX_TAB) k.TX = std::max(1, k.TX / 2);
    while (k.VT * k.TY > PT_MAX_TAB) k.TY = std::max(1, k.TY / 2);
    k.pitch = (k.VT * k.TX) | 1;
    k.v_chunks = (k.v_ext + k.VT - 1) / k.VT;
    k.x_chunks = (SX + k.TX - 1) / k.TX;
    k.y_chunks = (SY + k.TY - 1) / k.TY;
    int64_t outer = 1;
    for (int i = 0; i < k.n_out; i++) outer *= k.o_ext[i];
    int64_t grid = k.v_chunks * k.x_chunks * k.y_chunks * outer;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    size_t smem = (size_t)k.pitch * k.TY * esz;

    const bool planar = q.plane_stride != 0;
#define MB200_PERM(E, S, P) permute_kernel<E, S, P><<<(unsigned)grid, PT_THREADS, smem, s>>>(k, (const E *)src, dst)
    switch (dtype) {
        case MB200_F32: MB200_PERM(float, float, false); break;
        case MB200_F64: MB200_PERM(double, double, false); break;
        case MB200_C64: if (planar) MB200_PERM(float2, float, true); else MB200_PERM(float2, float, false); break;
        default: if (planar) MB200_PERM(double2, double, true); else MB200_PERM(double2, double, false); break;
    }
#undef MB200_PERM
    return cudaGetLastError();
}

}  // namespace mb200
