/* libmuscle_b200.so — C ABI of the B200-native `binary_einsum` backend for Muscle.jl.
 *
 * This is the drop-in boundary: what a `struct BackendB200 <: Muscle.Backend` binds with `ccall`
 * (julia/MuscleB200.jl). It replaces, for the pairwise-contraction hot path only,
 *   - Muscle.binary_einsum(::BackendBase, inds_c, a, b)      src/Operations/binary_einsum.jl:76-96
 *   - Muscle.binary_einsum!(::BackendBase, c, a, b)          src/Operations/binary_einsum.jl:98-121
 *   - the cuTENSOR route  cuTENSOR.contract!(1, A, modes_a, …, 0, C, modes_c, …)
 *                                                            ext/MuscleCUDAExt.jl:22-41
 * Labels never cross the boundary: the caller flattens `Index` objects to int32 mode ids exactly
 * as the reference does (ext/MuscleCUDAExt.jl:24-27).
 *
 * Conventions
 *   - arrays are Julia `Array`s: dense or strided, COLUMN-MAJOR, complex interleaved (re,im);
 *     strides are in ELEMENTS; `strides == NULL` means dense column-major in the given mode order.
 *   - every entry returns an `mb200_status_t`; the message is in `mb200_last_error_string()`
 *     (thread-local). INVALID_ARGUMENT / NOT_SUPPORTED map to Julia `ArgumentError`
 *     (binary_einsum.jl:53-55,82-83), DIMENSION_MISMATCH to `DimensionMismatch` (src/Tensor.jl:23).
 *   - calls are stream-ordered on the handle's stream and do not synchronise, except the `_host`
 *     entry and `mb200_stream_sync`.
 *   - there is NO CPU fallback: without a usable CUDA device compute entries return CUDA_ERROR.
 */
#ifndef MUSCLE_B200_H
#define MUSCLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_MAX_MODES 32

typedef enum {
    MB200_F32 = 0,  /* Float32    */
    MB200_F64 = 1,  /* Float64    */
    MB200_C64 = 2,  /* ComplexF32 */
    MB200_C128 = 3  /* ComplexF64 */
} mb200_dtype_t;

typedef enum {
    MB200_OK = 0,
    MB200_INVALID_ARGUMENT = 1,   /* -> ArgumentError       */
    MB200_NOT_SUPPORTED = 2,      /* -> ArgumentError       */
    MB200_DIMENSION_MISMATCH = 3, /* -> DimensionMismatch   */
    MB200_CUDA_ERROR = 4,
    MB200_OUT_OF_MEMORY = 5,
    MB200_INTERNAL_ERROR = 6
} mb200_status_t;

/* which kernel family a contraction is routed to (reported by mb200_plan_describe) */
typedef enum {
    MB200_PATH_AUTO = 0,
    MB200_PATH_DIRECT = 1,   /* K5: single-kernel direct contraction (launch-bound / skinny sizes) */
    MB200_PATH_GETT_F64 = 2, /* K2: FP64 DMMA (mma.sync m8n8k4 f64) gather-GEMM, C128 4M / F64    */
    MB200_PATH_SIMT_F32 = 3, /* FP32 FFMA gather-GEMM (C64 / F32; shapes the tcgen05 path rejects) */
    MB200_PATH_TCGEN05_TF32 = 4 /* K3: tcgen05/TMEM split-operand (TF32 + BF16) GEMM on TMA-fed operands */
} mb200_path_t;

/* Arithmetic of Float32 / ComplexF32 contractions (per handle, mb200_set_compute_type). Float64 / ComplexF64 always run
 * native FP64 FMAs (DMMA) and are not affected.
 *   DEFAULT  the north-star scheme: shapes with M >= 64, N >= 32, K >= 64 and >= 2^27 MACs run on tcgen05 tensor cores with
 *            split-operand error compensation, x = tf32_rn(x) + lo: hi*hi in TF32 plus both cross terms in ONE bf16 MMA
 *            (lo*lo is dropped, the cross terms carry bf16's 8-bit significand), chunked FP32 accumulation. Accuracy contract:
 *            relative Frobenius error <= 1e-5 against an FP64 reference (measured 1.3e-6 ComplexF32 / 0.9e-6 Float32, flat in
 *            K up to 16384) - about 10x the error of FP32 FMAs. Integer-valued operands are reproduced exactly only while
 *            every operand fits TF32's 11-bit significand (|x| <= 2048); smaller / skinnier shapes use FP32 FMAs, so the
 *            accuracy changes at the eligibility threshold.
 *   FP32     strict: every product and sum is an FP32 FMA on every shape (what cuTENSOR's default COMPUTE_32F and Muscle's
 *            BackendBase give); the tensor-core path is never taken.
 *   3XTF32   the classic three-pass split (hi*hi + hi*lo + lo*hi, all TF32) on the same tensor-core path. */
typedef enum {
    MB200_COMPUTE_DEFAULT = 0,
    MB200_COMPUTE_FP32 = 1,
    MB200_COMPUTE_3XTF32 = 2
} mb200_compute_type_t;

typedef struct mb200_handle_s *mb200_handle_t;

/* ---- library / handle ------------------------------------------------------------------ */
int mb200_version(void);
const char *mb200_last_error_string(void);
int mb200_device_count(int *count);
/* one handle per (device, host thread); owns plan cache, offset tables, workspace, stream */
int mb200_create(mb200_handle_t *handle, int device);
int mb200_destroy(mb200_handle_t handle);
/* `cuda_stream` is a cudaStream_t (NULL = legacy default stream). Not owned. */
int mb200_set_stream(mb200_handle_t handle, void *cuda_stream);
int mb200_stream_sync(mb200_handle_t handle);
/* force a kernel family (testing / benchmarking); MB200_PATH_AUTO restores the planner's choice */
int mb200_set_path(mb200_handle_t handle, int path);
/* mb200_compute_type_t for Float32 / ComplexF32 contractions on this handle (default MB200_COMPUTE_DEFAULT) */
int mb200_set_compute_type(mb200_handle_t handle, int compute_type);
int mb200_get_compute_type(mb200_handle_t handle, int *compute_type);

/* ---- device / pinned-host memory helpers (Julia has no CUDA.jl on this path) -------------- */
int mb200_malloc(mb200_handle_t handle, void **dptr, size_t bytes);
int mb200_free(mb200_handle_t handle, void *dptr);
int mb200_host_alloc(void **hptr, size_t bytes); /* pinned */
int mb200_host_free(void *hptr);
int mb200_memcpy_h2d(mb200_handle_t handle, void *dst, const void *src, size_t bytes); /* async on stream */
int mb200_memcpy_d2h(mb200_handle_t handle, void *dst, const void *src, size_t bytes); /* async on stream */
int mb200_memset(mb200_handle_t handle, void *dptr, int value, size_t bytes);

/* ---- the hot path ------------------------------------------------------------------------
 * C[modesC] = sum over modes absent from C of A[modesA] * B[modesB]        (alpha = 1, beta = 0)
 * Mode classes: batch/hyper = in A, B and C; summed = in A and B, not C; free = in one operand and C.
 * A mode carried by ONE operand and absent from C ("dangling") is summed over, as cuTENSOR.contract! and OMEinsum do
 * (ext/MuscleCUDAExt.jl:30-38, ext/MuscleOMEinsumExt.jl:40-59; BackendBase rejects it, binary_einsum.jl:83): launch-bound
 * sizes fold the sum into the contraction kernel (the other operand is broadcast along it), larger operands are
 * pre-reduced by one unary_einsum pass.
 * Rejected (INVALID_ARGUMENT): a mode repeated inside one tensor, a C mode found in neither
 * operand, nmodes > MB200_MAX_MODES.
 * DIMENSION_MISMATCH: a shared mode with different extents in A and B.
 * dtypes may differ between A and B (Float64 x ComplexF64 ...): the result dtype must be the
 * promotion of the two (Base.promote_eltype, ext/MuscleCUDAExt.jl:16).
 * Pointers are DEVICE pointers on the handle's device. Extents of C are implied by its modes.
 */
int mb200_binary_einsum(mb200_handle_t handle,
                        void *C, int dtypeC, int nmodeC, const int32_t *modesC, const int64_t *stridesC,
                        const void *A, int dtypeA, int nmodeA, const int32_t *modesA,
                        const int64_t *extentsA, const int64_t *stridesA,
                        const void *B, int dtypeB, int nmodeB, const int32_t *modesB,
                        const int64_t *extentsB, const int64_t *stridesB);

/* Same contract with HOST pointers (dense column-major only, strides must be NULL): copies A and B
 * to the device, contracts, copies C back, synchronises. This is what replays the reference's
 * host-array test batteries and what `e2e` in bench.py times. */
int mb200_binary_einsum_host(mb200_handle_t handle,
                             void *C, int dtypeC, int nmodeC, const int32_t *modesC,
                             const void *A, int dtypeA, int nmodeA, const int32_t *modesA,
                             const int64_t *extentsA,
                             const void *B, int dtypeB, int nmodeB, const int32_t *modesB,
                             const int64_t *extentsB);

/* ---- planner introspection (host only; no device needed) ----------------------------------- */
typedef struct {
    int64_t M, N, K, L;          /* GEMM-equivalent sizes after classification (L = batch)      */
    int32_t swapped;             /* 1: B supplies the row (M) modes so C's fastest mode is a row */
    int32_t path;                /* mb200_path_t the planner picked                              */
    int32_t compute_dtype;       /* promoted dtype                                               */
    int32_t n_left, n_right, n_sum, n_batch;
    int32_t left[MB200_MAX_MODES];  /* row modes, fastest first, in the order the kernel walks   */
    int32_t right[MB200_MAX_MODES]; /* column modes                                               */
    int32_t sum[MB200_MAX_MODES];   /* summed modes                                               */
    int32_t batch[MB200_MAX_MODES]; /* batch (hyper) modes                                        */
    int32_t a_kmajor, b_kmajor;  /* 1: the operand's unit-stride mode is a summed mode           */
    double flops;                /* 8*M*N*K*L complex, 2*M*N*K*L real                            */
    double bytes;                /* sizeof(T)*(|A|+|B|+|C|)                                      */
    int32_t tc_eligible;         /* 1: ComplexF32 / Float32 shape the tcgen05 path accepts (M >= 64, N >= 32, K >= 64) */
    int32_t tc_permute_pack;     /* 1: its operands are packed by a K1 permutation; 0 with tc_eligible: by the
                                    table-driven gather pack (summed extents not tiling groups of 8 k, strided operands) */
} mb200_plan_info_t;

int mb200_plan_describe(int dtypeC, int nmodeC, const int32_t *modesC, const int64_t *stridesC,
                        int dtypeA, int nmodeA, const int32_t *modesA, const int64_t *extentsA,
                        const int64_t *stridesA,
                        int dtypeB, int nmodeB, const int32_t *modesB, const int64_t *extentsB,
                        const int64_t *stridesB,
                        mb200_plan_info_t *info);

/* ---- K1: permute / matricise (Julia `permutedims(src, perm)`; perm is 0-based here) -------
 * dst mode d = src mode perm[d]; both dense column-major. `extents` are the SOURCE extents.
 * flags: MB200_PERMUTE_PLANAR writes a complex dst as two planes (all re, then all im). */
#define MB200_PERMUTE_PLANAR 1u
/* kernel choice for pure transpositions (A/B measurements, tests): TMA = the TMA-staged kernel (cp.async.bulk.tensor load -> shared-
 * memory transposition -> cp.async.bulk.tensor store) whenever the layout is eligible (<= 5 canonical modes, runs of >= 32 elements
 * on both sides, 16-byte aligned strides), NO_TMA = never; neither flag: the library's policy (MB200_PERMUTE_TMA env, default from
 * the measurements in DESIGN 3.5). Ineligible layouts silently take the register-tile / shared-memory-tile kernels. */
#define MB200_PERMUTE_TMA 2u
#define MB200_PERMUTE_NO_TMA 4u
int mb200_permute(mb200_handle_t handle, void *dst, const void *src, int dtype, int nmode,
                  const int64_t *extents, const int32_t *perm, uint32_t flags);

/* ---- the einsum family next to the hot path -----------------------------------
 * unary_einsum: Y[modesY] = sum over the modes of X absent from Y; a mode REPEATED inside X is read on its
 * diagonal (trace / diagonal extraction). Replaces `unary_einsum(!)(::BackendOMEinsum, y, x)`
 * (ext/MuscleOMEinsumExt.jl:25-38 -> OMEinsum.einsum!; front-end src/Operations/unary_einsum.jl:26-46).
 * INVALID_ARGUMENT: a mode of Y not in X ("Output indices must be a subset of input indices",
 * MuscleOMEinsumExt.jl:32), a mode repeated inside Y, dtypeY != dtypeX. DIMENSION_MISMATCH: a repeated mode
 * of X with different extents. Device pointers; strides NULL = dense column-major. */
int mb200_unary_einsum(mb200_handle_t handle,
                       void *Y, int dtypeY, int nmodeY, const int32_t *modesY, const int64_t *stridesY,
                       const void *X, int dtypeX, int nmodeX, const int32_t *modesX,
                       const int64_t *extentsX, const int64_t *stridesX);

/* hadamard: C = A .* broadcast(B), modes(B) a subset of modes(A) in any order; C has A's extents and layout
 * (dense column-major) and dtype promote(A, B); C may alias A when dtypeA == dtypeC. Replaces
 * `hadamard!(::BackendBase, c, a, b)` (src/Operations/hadamard.jl:50-77: permutedims of b + reshape +
 * broadcast multiply). INVALID_ARGUMENT: a mode of B not in A (hadamard.jl:10,29), repeated modes;
 * DIMENSION_MISMATCH: extents of a shared mode differ. */
int mb200_hadamard(mb200_handle_t handle, void *C, int dtypeC,
                   const void *A, int dtypeA, int nmodeA, const int32_t *modesA, const int64_t *extentsA,
                   const void *B, int dtypeB, int nmodeB, const int32_t *modesB, const int64_t *extentsB);

/* ---- thin SVD of a matricised tensor ---------------------------------------------
 * A (rows x cols, dense column-major, device) = U * diag(S) * V^H with k = min(rows, cols):
 * U rows x k, S k singular values (Float32 for Float32/ComplexF32, Float64 otherwise) in descending order,
 * Vt cols x k = conj(V) - exactly the three arrays `tensor_svd_thin(::BackendBase, A)` tensorifies
 * (src/Operations/tensor_svd.jl:100-124; `Vt = reshape(conj(V), ...)` :121), so A[u,v] = sum_s U[u,s] S[s] Vt[v,s].
 * Hand-written one-sided Jacobi (svd.cu). tol <= 0 selects sqrt(max(rows, cols)) * eps; max_sweeps <= 0 selects 30.
 * Stream-ordered, no host synchronisation. Limits: rows, cols < 2^31 and rows * cols elements of workspace.
 * The working copy is A scaled by an exact power of two so that max |a_ij| is in [1, 2): squared norms and inner products cannot
 * under- / overflow whatever A's magnitude (S is scaled back). Rank-deficient input: columns whose norm vanishes are replaced
 * by vectors orthonormal to all the others (Gram-Schmidt), so U and Vt are ALWAYS isometric, as LAPACK's are. */
int mb200_svd_thin(mb200_handle_t handle, void *U, void *S, void *Vt, const void *A, int dtype,
                   int64_t rows, int64_t cols, double tol, int max_sweeps);
/* Status of the last mb200_svd_thin on this handle (synchronises the stream): Jacobi sweeps used, whether the last sweep rotated
 * nothing (converged = 0: max_sweeps exhausted, the factors are approximate), how many null columns were completed. */
int mb200_svd_last_info(mb200_handle_t handle, int *sweeps, int *converged, int *completed_columns);

/* Thin QR: A (rows x cols, dense column-major, device) = Q * R with k = min(rows, cols): Q rows x k (orthonormal
 * columns), R k x cols (upper triangular) - the two arrays `tensor_qr_thin(::BackendBase, A)` tensorifies
 * (src/Operations/tensor_qr.jl:57-79). Hand-written Householder (svd.cu); R's diagonal carries the phase
 * -e^{i arg x_0} (QR is unique up to a diagonal phase). Stream-ordered. */
int mb200_qr_thin(mb200_handle_t handle, void *Q, void *R, const void *A, int dtype, int64_t rows, int64_t cols);

/* ---- multi-GPU partition planner (host only) ------------------------------------------------
 * One process per GPU (torch.distributed / NCCL does the plumbing). Mirrors Dagger's block
 * sharding, ext/MuscleDaggerExt/binary_einsum.jl:64-119: splitting a free or batch mode gives
 * independent output slabs (no collective); splitting a summed mode gives full-size partial
 * outputs that are add-reduced (treereduce(AddComputeOp) there, ncclAllReduce(sum) here). */
typedef enum { MB200_SHARD_NONE = 0, MB200_SHARD_FREE = 1, MB200_SHARD_BATCH = 2, MB200_SHARD_SUM = 3 } mb200_shard_kind_t;
typedef struct {
    int32_t kind;       /* mb200_shard_kind_t                                      */
    int32_t mode;       /* the mode id that is split                               */
    int64_t begin, end; /* this rank's half-open range of that mode                */
    int32_t needs_allreduce;
} mb200_shard_info_t;
/* prefer_sum != 0 asks for a summed-mode slice (config 5); otherwise the slowest free/batch mode
 * of C with extent >= nranks is split. */
int mb200_shard_plan(int nmodeC, const int32_t *modesC,
                     int nmodeA, const int32_t *modesA, const int64_t *extentsA,
                     int nmodeB, const int32_t *modesB, const int64_t *extentsB,
                     int nranks, int rank, int prefer_sum, mb200_shard_info_t *info);

/* ---- fused contraction + reduce-scatter over peer memory (summed-index slice) ---------------
 * One process per GPU. Every rank contracts its K-slice; instead of writing a full-size partial C and
 * all-reducing it (Dagger's treereduce(AddComputeOp), ext/MuscleDaggerExt/binary_einsum.jl:107-115), the GEMM
 * epilogue stores each output element straight into the staging buffer of the rank that OWNS it
 * (owner = flat offset in the dense column-major C >> slab_shift), slot `rank`, over NVLink peer mappings.
 * After a cross-rank barrier every rank sums its nranks slots (mb200_reduce_slots) and holds its slab of C.
 * staging[r] must point to nranks * (1 << slab_shift) elements on rank r (own buffer for r == rank,
 * mb200_ipc_import mappings for the others). Supported for the tensor-core paths (ComplexF64 DMMA,
 * ComplexF32 tcgen05); NOT_SUPPORTED otherwise (use an all-reduce). */
#define MB200_IPC_HANDLE_BYTES 64
#define MB200_MAX_PEERS 8
int mb200_ipc_export(mb200_handle_t handle, void *dptr, unsigned char *handle_out /* 64 bytes */);
int mb200_ipc_import(mb200_handle_t handle, const unsigned char *handle_in /* 64 bytes */, void **peer_ptr);
int mb200_ipc_release(mb200_handle_t handle, void *peer_ptr);
int mb200_binary_einsum_scatter(mb200_handle_t handle,
                                int dtypeC, int nmodeC, const int32_t *modesC,
                                const void *A, int dtypeA, int nmodeA, const int32_t *modesA,
                                const int64_t *extentsA, const int64_t *stridesA,
                                const void *B, int dtypeB, int nmodeB, const int32_t *modesB,
                                const int64_t *extentsB, const int64_t *stridesB,
                                void *const *staging, int nranks, int rank, int slab_shift);
/* out[e] = sum over s < nslots of staging_local[s * slab_elems + e]   (e < slab_elems) */
int mb200_reduce_slots(mb200_handle_t handle, void *out, const void *staging_local, int dtype,
                       int64_t slab_elems, int nslots);

/* ---- summed-index slice with the all-reduce FUSED into the contraction ("cross-GPU split-K") -------------
 * One process per GPU. Dagger contracts every pair of summed-index blocks and add-reduces the partial outputs
 * (treereduce(AddComputeOp), ext/MuscleDaggerExt/binary_einsum.jl:107-115); the NCCL baseline here is contraction then
 * ncclAllReduce. This entry overlaps the two: every rank contracts its K-slice into partial 128 x BN sub-tiles ("units") kept
 * in its own workspace and raises a flag in the unit's owner (unit % nranks); the owner's reducer kernel runs CONCURRENTLY on
 * a side stream, adds the nranks partials of each unit as soon as all of them exist (peer loads over NVLink, or one
 * multimem.ld_reduce when an NVLS multicast mapping is given) and stores the finished unit into the C of every rank (peer
 * stores or multimem.st) through C's offset tables, i.e. in any requested index order. Every rank ends with the same bits.
 * Buffers (comm): ws / c / flags of EVERY rank as pointers valid on this device (own + IPC or VMM peer mappings); flags must
 * be zero-initialised once; `epoch` starts at 1 and increases with every call on the same buffers. A call may reuse the
 * buffers of the previous call as soon as that call's DIST_WAIT phase has been enqueued (stream order) on every rank.
 * phases: CONTRACT | REDUCE | WAIT (7) is the production call. The separate bits exist so that a single GPU can emulate
 * several ranks in tests (all contractions first, then all reducers, then the waits).
 * Supported on the ComplexF32 / Float32 tcgen05 path; NOT_SUPPORTED otherwise (use an NCCL all-reduce). */
#define MB200_DIST_CONTRACT 1
#define MB200_DIST_REDUCE 2
#define MB200_DIST_WAIT 4
typedef struct {
    int32_t nranks, rank, epoch;
    void *ws[MB200_MAX_PEERS];    /* partial-unit workspace of every rank, ws_bytes each                     */
    void *c[MB200_MAX_PEERS];     /* output C of every rank (dense column-major in modesC order)              */
    void *flags[MB200_MAX_PEERS]; /* flag array of every rank, flag_bytes each, zeroed once                   */
    void *mc_ws, *mc_c;           /* NVLS multicast mappings of ws and c over all ranks, or both NULL          */
    size_t ws_bytes, flag_bytes;  /* sizes of the buffers above (checked against mb200_allreduce_workspace)    */
} mb200_comm_t;
/* sizes of the per-rank workspace and flag array for this contraction (host only; dense operands) */
int mb200_allreduce_workspace(mb200_handle_t handle, int dtypeC, int nmodeC, const int32_t *modesC,
                              int dtypeA, int nmodeA, const int32_t *modesA, const int64_t *extentsA,
                              int dtypeB, int nmodeB, const int32_t *modesB, const int64_t *extentsB,
                              int nranks, size_t *ws_bytes, size_t *flag_bytes);
int mb200_binary_einsum_allreduce(mb200_handle_t handle, int dtypeC, int nmodeC, const int32_t *modesC,
                                  const void *A, int dtypeA, int nmodeA, const int32_t *modesA,
                                  const int64_t *extentsA, const int64_t *stridesA,
                                  const void *B, int dtypeB, int nmodeB, const int32_t *modesB,
                                  const int64_t *extentsB, const int64_t *stridesB,
                                  const mb200_comm_t *comm, int phases);

/* Diagnostics: with MB200_DIST_TIMELINE=1 every fused all-reduce stamps %globaltimer (ns) into 8 slots - [0] first GEMM CTA starts,
 * [1] last GEMM epilogue ends, [2] first reducer CTA starts, [3] first reducer unit ready, [4] last reducer CTA done, [5] all done
 * flags seen - readable after the call with this entry (synchronises the device). */
int mb200_dist_timeline(mb200_handle_t handle, unsigned long long *out8);

/* ---- CUDA-graph replay of a fixed sequence of calls (n-ary contraction chains) ------------
 * Everything enqueued on the handle's stream between graph_begin and graph_end (binary_einsum, unary_einsum,
 * hadamard, permute ... on fixed device pointers) is captured instead of executed and can then be replayed with one
 * launch: a launch-bound chain of small contractions costs one graph launch instead of one kernel launch (plus
 * planner work) per step. The handle's stream must not be the legacy default stream. Every plan the sequence needs
 * must already be cached (run the sequence once before capturing): a plan miss during capture is NOT_SUPPORTED.
 * A graph co-owns every cached plan its kernel nodes point into (offset tables): evicting such a plan from the handle's
 * LRU cache, or destroying the handle, leaves the graph replayable; the tables go when mb200_graph_destroy runs. */
typedef struct mb200_graph_s *mb200_graph_t;
int mb200_graph_begin(mb200_handle_t handle);
int mb200_graph_end(mb200_handle_t handle, mb200_graph_t *graph);
int mb200_graph_launch(mb200_handle_t handle, mb200_graph_t graph);
int mb200_graph_destroy(mb200_graph_t graph);

/* Flag barrier for the fused reduce-scatter (replaces a collective between the scatter contraction and the slot sum).
 * Every rank owns an array of nranks ints (zero-initialised once, peer-mapped like the staging buffers).
 * mb200_signal_peers: enqueued after mb200_binary_einsum_scatter, stores `epoch` into entry `rank` of every rank's
 * array. mb200_reduce_slots_wait: like mb200_reduce_slots, but the kernel first waits until all nslots entries of the
 * LOCAL array have reached `epoch` (epochs must increase from call to call; a rank that never signals traps the
 * waiting kernel after ~10 s instead of hanging the GPU). */
int mb200_signal_peers(mb200_handle_t handle, void *const *flag_arrays, int nranks, int rank, int epoch);
int mb200_reduce_slots_wait(mb200_handle_t handle, void *out, const void *staging_local, int dtype,
                            int64_t slab_elems, int nslots, const void *flags_local, int epoch);

/* ---- counters (bench.py's gpu_launches claim) ----------------------------------------------- */
typedef struct {
    uint64_t launches_total;
    uint64_t launches_direct, launches_gett_f64, launches_simt_f32, launches_tcgen05;
    uint64_t launches_permute, launches_table, launches_convert, launches_reduce;
    uint64_t plans_built, plans_hit;
    uint64_t launches_unary, launches_hadamard, graph_launches, launches_svd;
    uint64_t launches_tcgen05_pair;   /* of launches_tcgen05: run by the CTA-pair (cta_group::2) kernel */
} mb200_stats_t;
int mb200_get_stats(mb200_handle_t handle, mb200_stats_t *stats);
int mb200_reset_stats(mb200_handle_t handle);

#ifdef __cplusplus
}
#endif
#endif /* MUSCLE_B200_H */
