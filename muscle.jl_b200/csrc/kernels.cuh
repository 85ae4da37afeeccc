// Launch interfaces of the sm_100a kernels behind libmuscle_b200 (one translation unit each).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "common.hpp"

namespace mb200 {

// ---- offset tables -----------------------------------------------------------------------------
// out[x] = sum_i digit_i(x) * stride[i], digits of x over `ext` (mode 0 fastest). n == 0 -> out[0]=0.
struct TableSpec {
    int n;
    int64_t ext[MB200_MAX_MODES];
    int64_t stride[MB200_MAX_MODES];
};
cudaError_t launch_build_table(int64_t *out, int64_t size, const TableSpec &spec, cudaStream_t s);

// ---- K5: direct single-kernel contraction ----------------------------------------------------------
struct DirectParams {
    int nc;  // C-walk modes: left, right, batch (extent-1 removed)
    int64_t c_ext[MB200_MAX_MODES], c_sc[MB200_MAX_MODES], c_sa[MB200_MAX_MODES], c_sb[MB200_MAX_MODES];
    int nk;  // summed modes
    int64_t k_ext[MB200_MAX_MODES], k_sa[MB200_MAX_MODES], k_sb[MB200_MAX_MODES];
    int64_t total_c, total_k;
};
cudaError_t launch_direct(int dtype, const DirectParams &p, const void *A, const void *B, void *C,
                          cudaStream_t s);

// ---- apply: a tiny operator (<= 8 x 8) contracted into a huge tensor (gate application, K <= 8, one free side <= 8) ---
// C[bigC(r) + jc[j]] = sum_k X[bigX(r) + kx[k]] * S[js[j] + ks[k]]   for every row r of the big operand X
struct ApplyParams {
    int nbig;
    int64_t big_ext[MB200_MAX_MODES], big_sx[MB200_MAX_MODES], big_sc[MB200_MAX_MODES];
    int64_t total_big;
    int J, K;
    int64_t kx[8], ks[8], js[8], jc[8];
};
cudaError_t launch_apply(int dtype, const ApplyParams &p, const void *X, const void *S, void *C, cudaStream_t s);

// ---- K2 / FP32 SIMT: gather-GEMM ---------------------------------------------------------------------
// C[rowC[m] + colC[n] + batC[l]] = sum_k A[rowA[m] + kA[k] + batA[l]] * B[colB[n] + kB[k] + batB[l]]
// All offsets are in ELEMENTS of the compute dtype.
// Epilogue redirection for the fused reduce-scatter: element at flat offset e of C goes to
// peer[e >> shift] + (rank << shift) + (e & mask). nranks == 0: plain store to C.
struct ScatterDesc {
    void *peer[MB200_MAX_PEERS];
    int nranks, rank, shift;
};
template <typename E>
__device__ __forceinline__ E *scatter_ptr(const ScatterDesc &sc, E *C, int64_t off) {
    if (sc.nranks == 0) return C + off;
    const int owner = (int)(off >> sc.shift);
    return reinterpret_cast<E *>(sc.peer[owner]) + (((int64_t)sc.rank << sc.shift) + (off & (((int64_t)1 << sc.shift) - 1)));
}

struct GettParams {
    const void *A, *B;
    void *C;
    ScatterDesc sc;
    const int64_t *rowA, *kA, *batA;
    const int64_t *colB, *kB, *batB;
    const int64_t *rowC, *colC, *batC;
    int64_t M, N, K, L;
    int a_kmajor, b_kmajor;
    int k_pairs;   // Float64 only: both operands are K-major with a unit-stride summed mode of even extent and every other stride
                   // even, so consecutive k pairs are 16-byte elements in both tensors (gett.cu CoreD2)
};
cudaError_t launch_gett_f64(int dtype, const GettParams &p, cudaStream_t s);   // F64 / C128, DMMA
cudaError_t launch_simt_f32(int dtype, const GettParams &p, cudaStream_t s);   // F32 / C64, FFMA
cudaError_t gett_configure();  // opt-in shared memory attributes; call once per device
cudaError_t permute_configure();
cudaError_t tf32_configure();

cudaError_t launch_reduce_slots(int dtype, void *out, const void *staging, int64_t slab_elems, int nslots, cudaStream_t s);
// flag barrier of the fused reduce-scatter: flags.peer[r] = rank r's flag array (nranks ints), flags.rank = this rank
cudaError_t launch_signal_peers(const ScatterDesc &flags, int epoch, cudaStream_t s);
cudaError_t launch_reduce_slots_wait(int dtype, void *out, const void *staging, int64_t slab_elems, int nslots, const int *flags,
                                     int epoch, cudaStream_t s);

// ---- unary_einsum / hadamard (elementwise.cu) ------------------------------------------------------------
// y[sum_i d_i * c_sy[i]] = sum over k of x[sum_i d_i * c_sx[i] + sum_j k_j * k_sx[j]]
struct UnaryParams {
    int nc;  // output-walk modes, y memory order (extent-1 removed, contiguous neighbours merged)
    int64_t c_ext[MB200_MAX_MODES], c_sx[MB200_MAX_MODES], c_sy[MB200_MAX_MODES];
    int nk;  // summed modes, ascending x stride
    int64_t k_ext[MB200_MAX_MODES], k_sx[MB200_MAX_MODES];
    int64_t total_c, total_k;
};
int unary_nsplit(const UnaryParams &p);   // > 1: the split form is used and needs nsplit * total_c elements of scratch
cudaError_t launch_unary(int dtype, const UnaryParams &p, const void *X, void *Y, void *part, int nsplit, cudaStream_t s);

// c[i] = a[i] * b[sum_m digit_m(i) * sb[m]] over a's (merged) modes; a and c dense, same layout
struct HadamardParams {
    int n;
    int64_t ext[MB200_MAX_MODES], sb[MB200_MAX_MODES];
    int64_t total;
    int b_vec_aligned;   // b may be read with 16-byte loads along mode 0
};
cudaError_t launch_hadamard(int dtype, const HadamardParams &p, const void *A, const void *B, void *C, cudaStream_t s);

// ---- thin SVD (svd.cu): one-sided Jacobi, cooperative launch ---------------------------------------------
// A rows x cols (dense column-major) -> U rows x k, S k (real), Vt cols x k = conj(right vectors), k = min(rows, cols).
// work_G: max(rows, cols) * k elements, work_V: k * k elements, counters: 2 ints.
cudaError_t launch_svd(int dtype, const void *A, int rows, int cols, void *U, void *S, void *Vt, void *work_G, void *work_V,
                       int *counters, double tol, int max_sweeps, cudaStream_t s);

// thin QR (Householder): A m x n -> Q m x k, R k x n (k = min(m, n)); work_W m*n elements, work_rd k elements,
// work_tau k reals
cudaError_t launch_qr(int dtype, const void *A, int m, int n, void *Q, void *Rm, void *work_W, void *work_rd, void *work_tau,
                      cudaStream_t s);

// ---- dtype promotion (mixed-eltype operands) -------------------------------------------------------
cudaError_t launch_convert(int dtype_dst, void *dst, int dtype_src, const void *src, int64_t n,
                           cudaStream_t s);

// ---- K1: permute / matricise ---------------------------------------------------------------------------
struct PermuteParams {
    int n;                               // canonical modes (extent-1 dropped, adjacent merged)
    int64_t ext[MB200_MAX_MODES];        // in SOURCE order (mode 0 = source unit stride)
    int64_t dst_stride[MB200_MAX_MODES]; // stride of each source mode in the destination (elements)
    int64_t total;
    int64_t plane_stride;                // != 0: complex dst written planar, im plane at +plane_stride
    int tma;                             // 1: prefer the TMA-staged transposition kernel when eligible, -1: never, 0: library policy
    int split;                           // != 0: write an operand of the tcgen05 kernel (tf32.cu) — per 8 k, ComplexF32 -> four chunks of 8
                                         //    words (re_hi, re_x, im_hi, im_x at +0,+8,+16,+24), Float32 -> two (hi, x at +0,+8).
                                         //    1: 3xTF32 format (x = fp32 remainder); 2 / 3: mixed TF32 + BF16 format of the row / column
                                         //    operand (x = bf16 pair carrying both cross terms)
};
cudaError_t launch_permute(int dtype, const PermuteParams &p, const void *src, void *dst, cudaStream_t s);
// Table-driven pack of one operand into the tcgen05 operand format, any strides and any K (zero-padded to Kp, a multiple of 8):
// dst[l][row][W * Kp] <- src[row_tab[row] + k_tab[k] + bat_tab[l]]; split = PermuteParams::split (1, 2 or 3)
// Optional row visiting order: the row modes listed by ascending SOURCE stride with their extents and their weights in the GEMM's row
// index (product of the extents of the modes before them in the C-order walk). A tile of 64 consecutive n' then covers runs that are
// contiguous in the source even when C's order scatters them (rank-8 dim-8 operands: 8-byte gathers become 64-byte runs).
#define MB200_PACK_DIGITS 6
struct PackRowOrder { int nd; int64_t ext[MB200_PACK_DIGITS], weight[MB200_PACK_DIGITS]; };
cudaError_t launch_pack_gather(int dtype, const void *src, const int64_t *row_tab, const int64_t *k_tab, const int64_t *bat_tab,
                               int64_t rows, int64_t K, int64_t Kp, int64_t L, int kmajor, int split, float *dst, cudaStream_t s,
                               const PackRowOrder *order = nullptr);

// ---- K3: tcgen05 / TMEM 3xTF32 ComplexF32 GEMM on packed operands (tf32.cu) ---------------------------------
bool tf32_available();
// mixed: operands are in the TF32 + BF16 format (split 2 / 3), else 3xTF32 (split 1); *pair (optional) reports whether the
// CTA-pair kernel (cta_group::2, 256-row tiles) was launched
cudaError_t launch_tf32_gemm(int dtype, const void *packA, const void *packB, const GettParams &g, bool mixed, cudaStream_t s,
                             bool *pair = nullptr, const struct DistDesc *dist = nullptr);

// ---- cross-GPU split-K: summed-index slice with the all-reduce fused into the contraction (tf32.cu) -----------------------
// Every rank contracts its K-slice into PARTIAL tiles, stored tile-linear (unit = one CTA's 128 x BN sub-tile, row fastest)
// in its own workspace ws[rank]; the epilogue then raises flag[unit][rank] in the flag array of the unit's OWNER
// (owner = unit % nranks) with a system-scope release store. The owner's reducer kernel - running concurrently with the
// GEMM on a second stream, on SMs the GEMM's persistent grid leaves free - waits for the nranks flags of each unit it owns, adds the nranks partial sub-tiles in rank
// order (peer loads over NVLink, or ONE multimem.ld_reduce through an NVLS multicast mapping: the switch adds), and stores
// the finished sub-tile through the rowC / colC / batC tables into the C of EVERY rank (peer stores, or one multimem.st).
// When its last unit is done it raises done[rank] on every rank; a rank's C is complete when all nranks done flags have
// arrived (wait_done kernel). All ranks end with bit-identical C. Flag array layout (int32): [nunits * nranks] unit flags,
// [nranks] done flags, [1] CTA counter; flags carry the call's epoch (strictly increasing), so nothing is ever reset.
struct DistDesc {
    int nranks, rank, epoch;
    void *ws[MB200_MAX_PEERS];     // partial-tile workspace of every rank (peer mappings; [rank] = own)
    void *c[MB200_MAX_PEERS];      // output C of every rank (peer mappings; [rank] = own)
    int *flags[MB200_MAX_PEERS];   // flag array of every rank
    void *mc_ws, *mc_c;            // NVLS multicast mappings of ws / c, or NULL
    int reserve_sms;               // > 0: the reducer runs NEXT TO the GEMM on this many SMs, which the GEMM's persistent grid leaves free;
                                   // 0: the reducer runs after the GEMM on every SM (no overlap)
    unsigned long long *timeline;  // optional 8 x u64 (device): globaltimer stamps [0] GEMM first CTA start, [1] GEMM last epilogue end,
                                   // [2] reducer first CTA start, [3] reducer first unit ready, [4] reducer end, [5] wait_done end
};
struct DistGeometry { int BN; int pair; int64_t nunits; int64_t unit_elems; };   // unit_elems = 128 * BN
// geometry of the dist-mode GEMM for a (dtype, M, N, L) problem: which kernel variant runs, how many units it produces
DistGeometry tf32_dist_geometry(int dtype, int64_t M, int64_t N, int64_t L);
cudaError_t launch_tf32_allreduce(int dtype, const GettParams &g, const DistDesc &dist, cudaStream_t s);
cudaError_t launch_dist_wait_done(const DistDesc &dist, int64_t nunits, cudaStream_t s);

}  // namespace mb200
