// libmuscle_b200 — C ABI (include/muscle_b200.h). Host-side glue: handles, plan cache, offset tables,
// dtype promotion, launches. No CPU compute path exists here: every compute entry ends in a kernel
// launch or an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"
#include "plan.hpp"

using namespace mb200;

namespace {

// A cached plan is shared between the handle's LRU cache and every captured graph whose kernel nodes carry raw pointers
// into its offset tables (GettParams rowA/rowC/... are passed by value at capture time): the tables are freed when the
// LAST owner lets go, so evicting a plan from the cache can never pull the tables from under a graph replay.
struct CachedPlan {
    Plan plan;
    int device = 0;
    int64_t *tables = nullptr;  // one device allocation: rowA rowC colB colC kA kB batA batB batC
    GettParams gp{};
    DirectParams dp{};
    ApplyParams ap{};
    bool use_apply = false, apply_big_is_row = true;
    std::list<std::string>::iterator lru;
    CachedPlan() = default;
    CachedPlan(const CachedPlan &) = delete;
    CachedPlan &operator=(const CachedPlan &) = delete;
    ~CachedPlan() {
        if (!tables) return;
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(device);
        // launches that read these tables may be in flight on ANY stream the handle has used (mb200_set_stream can
        // change it between calls): wait for the whole device, not only the current stream
        cudaDeviceSynchronize();
        cudaFree(tables);
        cudaSetDevice(cur);
    }
};

constexpr size_t PLAN_CACHE_CAP = 128;

}  // namespace

struct mb200_handle_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    int forced_path = MB200_PATH_AUTO;
    int compute_type = MB200_COMPUTE_DEFAULT;
    std::unordered_map<std::string, std::shared_ptr<CachedPlan>> cache;
    std::list<std::string> lru;
    mb200_stats_t stats{};
    std::mutex mu;
    bool capturing = false;
    std::vector<std::shared_ptr<CachedPlan>> captured;   // plans touched since mb200_graph_begin
    // cross-GPU split-K: the reducer runs on a high-priority side stream, concurrently with the GEMM on `stream`
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    unsigned long long *timeline = nullptr;   // MB200_DIST_TIMELINE=1: globaltimer stamps of the last fused all-reduce (diagnostics)
    int *svd_info = nullptr;   // 8 ints written by the last mb200_svd_thin on this handle (sweeps, converged, ...)
};

struct mb200_graph_s {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int device = 0;
    std::vector<std::shared_ptr<CachedPlan>> plans;      // keeps the offset tables its kernel nodes point into alive
};

namespace {

int cuda_fail(cudaError_t e, const char *what) {
    int st = (e == cudaErrorMemoryAllocation) ? MB200_OUT_OF_MEMORY : MB200_CUDA_ERROR;
    return fail(st, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

#define MB200_CUDA(call)                                   \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

#define MB200_CHECK_HANDLE(h) \
    if (!(h)) return fail(MB200_INVALID_ARGUMENT, "handle is NULL")

TableSpec spec_of(const std::vector<GroupMode> &g, int which) {
    TableSpec s{};
    s.n = (int)g.size();
    for (int i = 0; i < s.n; i++) {
        s.ext[i] = g[i].extent;
        s.stride[i] = which == 0 ? g[i].sa : (which == 1 ? g[i].sb : g[i].sc);
    }
    return s;
}

int build_cached(mb200_handle_t h, CachedPlan &cp) {
    const Plan &p = cp.plan;
    if (p.path == MB200_PATH_DIRECT) {
        DirectParams &d = cp.dp;
        std::memset(&d, 0, sizeof d);
        auto add_c = [&](const GroupMode &g, bool in_a, bool in_b) {
            d.c_ext[d.nc] = g.extent;
            d.c_sc[d.nc] = g.sc;
            d.c_sa[d.nc] = in_a ? g.sa : 0;
            d.c_sb[d.nc] = in_b ? g.sb : 0;
            d.nc++;
        };
        for (auto &g : p.mleft) add_c(g, true, false);
        for (auto &g : p.mright) add_c(g, false, true);
        for (auto &g : p.mbatch) add_c(g, true, true);
        for (auto &g : p.msum) {
            d.k_ext[d.nk] = g.extent;
            d.k_sa[d.nk] = g.sa;
            d.k_sb[d.nk] = g.sb;
            d.nk++;
        }
        d.total_c = p.M * p.N * p.L;
        d.total_k = p.K;
        if (p.apply_like && !p.empty_output && p.M * p.N > 0) {
            // big operand = the one with the large free group; the other one is the (<= 8 x 8) operator
            const bool big_row = p.M >= p.N;
            const std::vector<GroupMode> &big = big_row ? p.mleft : p.mright;
            const std::vector<GroupMode> &small = big_row ? p.mright : p.mleft;
            ApplyParams &a = cp.ap;
            std::memset(&a, 0, sizeof a);
            a.nbig = (int)big.size();
            a.total_big = big_row ? p.M : p.N;
            for (int i = 0; i < a.nbig; i++) {
                a.big_ext[i] = big[i].extent;
                a.big_sx[i] = big_row ? big[i].sa : big[i].sb;
                a.big_sc[i] = big[i].sc;
            }
            auto enumerate = [](const std::vector<GroupMode> &g, int which, int64_t *out, int64_t n) {
                for (int64_t x = 0; x < n; x++) {
                    int64_t r = x, off = 0;
                    for (const GroupMode &m : g) {
                        off += (r % m.extent) * (which == 0 ? m.sa : (which == 1 ? m.sb : m.sc));
                        r /= m.extent;
                    }
                    out[x] = off;
                }
            };
            a.J = (int)(big_row ? p.N : p.M);
            a.K = (int)p.K;
            enumerate(small, big_row ? 1 : 0, a.js, a.J);
            enumerate(small, 2, a.jc, a.J);
            enumerate(p.msum, big_row ? 0 : 1, a.kx, a.K);
            enumerate(p.msum, big_row ? 1 : 0, a.ks, a.K);
            cp.use_apply = true;
            cp.apply_big_is_row = big_row;
        }
        return MB200_OK;
    }
    // gather-GEMM: offset tables, built on the device, cached with the plan
    const int64_t M = p.M, N = p.N, K = p.K, L = p.L;
    const int64_t total = 2 * M + 2 * N + 2 * K + 3 * L;
    MB200_CUDA(cudaMalloc(&cp.tables, (size_t)total * sizeof(int64_t)));
    int64_t *t = cp.tables;
    int64_t *rowA = t; t += M;
    int64_t *rowC = t; t += M;
    int64_t *colB = t; t += N;
    int64_t *colC = t; t += N;
    int64_t *kA = t; t += K;
    int64_t *kB = t; t += K;
    int64_t *batA = t; t += L;
    int64_t *batB = t; t += L;
    int64_t *batC = t; t += L;
    cudaStream_t s = h->stream;
    MB200_CUDA(launch_build_table(rowA, M, spec_of(p.mleft, 0), s));
    MB200_CUDA(launch_build_table(rowC, M, spec_of(p.mleft, 2), s));
    MB200_CUDA(launch_build_table(colB, N, spec_of(p.mright, 1), s));
    MB200_CUDA(launch_build_table(colC, N, spec_of(p.mright, 2), s));
    MB200_CUDA(launch_build_table(kA, K, spec_of(p.msum, 0), s));
    MB200_CUDA(launch_build_table(kB, K, spec_of(p.msum, 1), s));
    MB200_CUDA(launch_build_table(batA, L, spec_of(p.mbatch, 0), s));
    MB200_CUDA(launch_build_table(batB, L, spec_of(p.mbatch, 1), s));
    MB200_CUDA(launch_build_table(batC, L, spec_of(p.mbatch, 2), s));
    h->stats.launches_table += 9;
    h->stats.launches_total += 9;
    GettParams &g = cp.gp;
    g.rowA = rowA; g.rowC = rowC; g.colB = colB; g.colC = colC; g.kA = kA; g.kB = kB;
    g.batA = batA; g.batB = batB; g.batC = batC;
    g.M = M; g.N = N; g.K = K; g.L = L;
    g.a_kmajor = p.a_kmajor; g.b_kmajor = p.b_kmajor;
    g.k_pairs = 0;
    if (p.dtype == MB200_F64 && !p.msum.empty() && p.msum[0].sa == 1 && p.msum[0].sb == 1 && p.msum[0].extent % 2 == 0) {
        bool even = true;   // every other stride of A and B keeps the pairs 16-byte aligned
        for (size_t i = 1; i < p.msum.size(); i++) even = even && p.msum[i].sa % 2 == 0 && p.msum[i].sb % 2 == 0;
        for (const GroupMode &m : p.mleft) even = even && m.sa % 2 == 0;
        for (const GroupMode &m : p.mright) even = even && m.sb % 2 == 0;
        for (const GroupMode &m : p.mbatch) even = even && m.sa % 2 == 0 && m.sb % 2 == 0;
        g.k_pairs = even ? 1 : 0;
    }
    return MB200_OK;
}

void evict_one(mb200_handle_t h) {
    if (h->lru.empty()) return;
    std::string key = h->lru.back();
    h->lru.pop_back();
    auto it = h->cache.find(key);
    if (it != h->cache.end()) h->cache.erase(it);   // ~CachedPlan frees the tables once no captured graph holds the plan
}

// Split scheme of the tcgen05 path: mixed TF32 + BF16 (8 MMAs per 8 k of a complex product) unless MB200_SPLIT_SCHEME=3xtf32
// asks for the all-TF32 scheme (12 MMAs; A/B measurements, tools/ab_c64.py).
bool tf32_mixed(const mb200_handle_t h) {
    static const bool env_mixed = [] { const char *e = getenv("MB200_SPLIT_SCHEME"); return !(e && std::string(e) == "3xtf32"); }();
    return env_mixed && h->compute_type != MB200_COMPUTE_3XTF32;
}

// K1 pack of one operand into the tcgen05 kernel's operand format: [batch][rows][4*K] floats, K-major,
// every 8-k group stored as re_hi | re_lo | im_hi | im_lo chunks of 8 floats (tf32.cu). Expressed as an
// ordinary strided permutation: the leading summed modes tile the group of 8 (the mode that completes it is split
// into (need, e/need) with destination strides (kstride, 32)); later summed modes get 4*kstride, row modes
// rows*4K, batch modes beyond.
bool build_pack_params(const Plan &p, int which, bool mixed, PermuteParams &q, int64_t &rows) {
    struct Md { int64_t ext, ss, ds; };
    std::vector<Md> v;
    const int64_t K = p.K;
    const int64_t W = p.dtype == MB200_F32 ? 2 : 4;   // floats per k: (hi, lo) real, (re_hi, re_lo, im_hi, im_lo) complex
    int64_t kstride = 1;
    bool grouped = false;   // the first group of 8 k has been tiled by the modes seen so far
    for (const GroupMode &g : p.sum) {
        const int64_t ss = which ? g.sb : g.sa;
        if (grouped) {
            v.push_back({g.extent, ss, W * kstride});        // kstride is a multiple of 8: (kstride / 8) groups of 8 W floats
        } else {
            const int64_t need = 8 / kstride;
            if (g.extent % need == 0) {                       // completes the group: split into (need, extent / need)
                v.push_back({need, ss, kstride});
                if (g.extent / need > 1) v.push_back({g.extent / need, ss * need, 8 * W});
                grouped = true;
            } else if (need % g.extent == 0) {                // still inside the group of 8
                v.push_back({g.extent, ss, kstride});
            } else {
                return false;
            }
        }
        kstride *= g.extent;
    }
    if (!grouped) return false;
    int64_t rstride = 1;
    for (const GroupMode &g : (which ? p.right : p.left)) {
        v.push_back({g.extent, which ? g.sb : g.sa, rstride * W * K});
        rstride *= g.extent;
    }
    rows = rstride;
    int64_t bstride = 1;
    for (const GroupMode &g : p.batch) {
        v.push_back({g.extent, which ? g.sb : g.sa, bstride * rows * W * K});
        bstride *= g.extent;
    }
    std::sort(v.begin(), v.end(), [](const Md &a, const Md &b) { return a.ss < b.ss; });
    if (v.size() > (size_t)MB200_MAX_MODES) return false;
    q = PermuteParams{};
    q.n = (int)v.size();
    q.total = 1;
    int64_t expect = 1;
    for (int i = 0; i < q.n; i++) {
        if (v[i].ss != expect) return false;   // not dense: the planner should have caught it
        expect *= v[i].ext;
        q.ext[i] = v[i].ext;
        q.dst_stride[i] = v[i].ds;
        q.total *= v[i].ext;
    }
    q.split = mixed ? 2 + which : 1;
    return true;
}

// Row visiting order of the line-writer pack for one operand (which = 0: row operand / mleft, 1: column operand / mright): the
// merged row modes sorted by the operand's own stride. nd = 0 when that is already the C-order walk or there are too many modes.
PackRowOrder pack_row_order(const Plan &p, int which) {
    PackRowOrder o{};
    const std::vector<GroupMode> &g = which ? p.mright : p.mleft;
    if (g.size() < 2 || g.size() > (size_t)MB200_PACK_DIGITS) return o;
    std::vector<int> idx(g.size());
    std::vector<int64_t> w(g.size());
    int64_t acc = 1;
    for (size_t i = 0; i < g.size(); i++) { idx[i] = (int)i; w[i] = acc; acc *= g[i].extent; }
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return (which ? g[a].sb : g[a].sa) < (which ? g[b].sb : g[b].sa); });
    bool identity = true;
    for (size_t i = 0; i < idx.size(); i++) identity = identity && idx[i] == (int)i;
    if (identity) return o;
    o.nd = (int)g.size();
    for (size_t i = 0; i < idx.size(); i++) { o.ext[i] = g[idx[i]].extent; o.weight[i] = w[idx[i]]; }
    return o;
}

// span (in elements) touched by a possibly strided tensor
int64_t span_of(const TensorDesc &t) {
    int64_t s = 1;
    for (int i = 0; i < t.n; i++) {
        if (t.ext[i] == 0) return 0;
        s += (t.ext[i] - 1) * t.stride[i];
    }
    return s;
}

int ensure_side_stream(mb200_handle_t h) {
    if (h->side) return MB200_OK;
    int lo = 0, hi = 0;
    MB200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    (void)lo;
    MB200_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));   // the reducer's few CTAs are placed first
    MB200_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    MB200_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    return MB200_OK;
}

constexpr int DIST_CONTRACT = 1, DIST_REDUCE = 2, DIST_WAIT = 4;

int contract_device(mb200_handle_t h, void *C, TensorDesc &dC, const int64_t *stridesC, const void *A,
                    const TensorDesc &dA, const void *B, const TensorDesc &dB, const ScatterDesc *sc = nullptr,
                    const DistDesc *dist = nullptr, int phases = 0, size_t dist_ws_bytes = 0, size_t dist_flag_bytes = 0) {
    Plan plan;
    int st = make_plan(dA, dB, dC, stridesC, h->forced_path, plan);
    if (st != MB200_OK) return st;
    if (h->compute_type == MB200_COMPUTE_FP32 && plan.path == MB200_PATH_TCGEN05_TF32) {
        // strict FP32: every product and sum is an FP32 FMA (cuTENSOR COMPUTE_32F / BackendBase accuracy)
        if ((st = make_plan(dA, dB, dC, stridesC, MB200_PATH_SIMT_F32, plan)) != MB200_OK) return st;
    }
    if (plan.empty_output) return MB200_OK;
    if ((!C && !sc && !dist) || (!A && dA.numel() > 0) || (!B && dB.numel() > 0))
        return fail(MB200_INVALID_ARGUMENT, "NULL data pointer");
    if (dist) {
        if (!(plan.path == MB200_PATH_TCGEN05_TF32 && tf32_available()))
            return fail(MB200_NOT_SUPPORTED, "the fused all-reduce runs on the ComplexF32 / Float32 tcgen05 path (this contraction is "
                                             "planned on path %d); use an NCCL all-reduce of the partial outputs", plan.path);
        const DistGeometry geo = tf32_dist_geometry(plan.dtype, plan.M, plan.N, plan.L);
        const size_t need_ws = (size_t)geo.nunits * geo.unit_elems * dtype_size(plan.dtype);
        const size_t need_fl = ((size_t)geo.nunits * dist->nranks + dist->nranks + 1) * sizeof(int);
        if (dist_ws_bytes < need_ws || dist_flag_bytes < need_fl)
            return fail(MB200_INVALID_ARGUMENT, "all-reduce workspace too small: %zu / %zu bytes given, %zu / %zu needed "
                                                "(mb200_allreduce_workspace)", dist_ws_bytes, dist_flag_bytes, need_ws, need_fl);
    }
    if (sc) {
        const bool tensor_path = (plan.path == MB200_PATH_GETT_F64 && plan.dtype == MB200_C128) ||
                                 (plan.path == MB200_PATH_TCGEN05_TF32 && tf32_available());
        if (!tensor_path)
            return fail(MB200_NOT_SUPPORTED, "fused reduce-scatter needs the ComplexF64 DMMA or ComplexF32 tcgen05 path "
                                             "(this contraction is planned on path %d); use an all-reduce", plan.path);
        if (dC.numel() != ((int64_t)sc->nranks << sc->shift))
            return fail(MB200_INVALID_ARGUMENT, "C has %lld elements, expected nranks << slab_shift = %lld",
                        (long long)dC.numel(), (long long)((int64_t)sc->nranks << sc->shift));
    }
    MB200_CUDA(cudaSetDevice(h->device));

    CachedPlan *cp;
    auto it = h->cache.find(plan.key);
    if (it != h->cache.end()) {
        cp = it->second.get();
        h->lru.erase(cp->lru);
        h->lru.push_front(plan.key);
        cp->lru = h->lru.begin();
        h->stats.plans_hit++;
    } else {
        if (h->capturing)
            return fail(MB200_NOT_SUPPORTED, "plan miss during graph capture: run the sequence once before mb200_graph_begin");
        while (h->cache.size() >= PLAN_CACHE_CAP) evict_one(h);
        auto fresh = std::make_shared<CachedPlan>();
        fresh->plan = plan;
        fresh->device = h->device;
        st = build_cached(h, *fresh);
        if (st != MB200_OK) return st;
        h->lru.push_front(plan.key);
        fresh->lru = h->lru.begin();
        cp = fresh.get();
        it = h->cache.emplace(plan.key, std::move(fresh)).first;
        h->stats.plans_built++;
    }
    if (h->capturing) h->captured.push_back(it->second);
    const Plan &p = cp->plan;
    cudaStream_t s = h->stream;

    // row / column operands, promoted to the compute dtype when the eltypes are mixed
    const void *R = p.swapped ? B : A;
    const void *Q = p.swapped ? A : B;
    const TensorDesc &dR = p.swapped ? dB : dA;
    const TensorDesc &dQ = p.swapped ? dA : dB;
    void *tmpR = nullptr, *tmpQ = nullptr;
    if (dR.dtype != p.dtype) {
        int64_t n = span_of(dR);
        MB200_CUDA(cudaMallocAsync(&tmpR, (size_t)std::max<int64_t>(n, 1) * dtype_size(p.dtype), s));
        MB200_CUDA(launch_convert(p.dtype, tmpR, dR.dtype, R, n, s));
        h->stats.launches_convert++; h->stats.launches_total++;
        R = tmpR;
    }
    if (dQ.dtype != p.dtype) {
        int64_t n = span_of(dQ);
        MB200_CUDA(cudaMallocAsync(&tmpQ, (size_t)std::max<int64_t>(n, 1) * dtype_size(p.dtype), s));
        MB200_CUDA(launch_convert(p.dtype, tmpQ, dQ.dtype, Q, n, s));
        h->stats.launches_convert++; h->stats.launches_total++;
        Q = tmpQ;
    }

    cudaError_t e;
    PermuteParams qa, qb;
    int64_t rows_a = 0, rows_b = 0;
    if (p.path == MB200_PATH_DIRECT) {
        if (cp->use_apply && !sc)
            e = launch_apply(p.dtype, cp->ap, cp->apply_big_is_row ? R : Q, cp->apply_big_is_row ? Q : R, C, s);
        else
            e = launch_direct(p.dtype, cp->dp, R, Q, C, s);
        h->stats.launches_direct++;
    } else if (p.path == MB200_PATH_TCGEN05_TF32 && tf32_available()) {
        // pack A, pack B (K1 with the split writer: a strided permutation when the layout allows it, else the table-driven
        // gather pack with K zero-padded to a multiple of 8), then the tcgen05 GEMM with the permuting epilogue
        const bool mixed = tf32_mixed(h);
        // MB200_PACK=permute: the K1 strided permutation with the 4-byte split writers (round 1); default: the table-driven
        // line-writer pack for every layout
        static const bool permute_pack = [] { const char *e = getenv("MB200_PACK"); return e && std::string(e) == "permute"; }();
        const bool by_permute = permute_pack && p.tc_permute_pack && build_pack_params(p, 0, mixed, qa, rows_a) &&
                                build_pack_params(p, 1, mixed, qb, rows_b);
        if (!by_permute) { rows_a = p.M; rows_b = p.N; }
        const int64_t Kp = by_permute ? p.K : (p.K + 7) / 8 * 8;
        void *pa = nullptr, *pb = nullptr;
        const size_t W = p.dtype == MB200_F32 ? 2 : 4;
        const size_t ba = (size_t)p.L * rows_a * W * Kp * sizeof(float), bb = (size_t)p.L * rows_b * W * Kp * sizeof(float);
        if (dist && !(phases & DIST_CONTRACT)) {
            e = cudaSuccess;
            if (phases & DIST_REDUCE) {
                e = launch_tf32_allreduce(p.dtype, cp->gp, *dist, s);
                h->stats.launches_reduce++; h->stats.launches_total++;
            }
            if (e == cudaSuccess && (phases & DIST_WAIT)) {
                e = launch_dist_wait_done(*dist, tf32_dist_geometry(p.dtype, p.M, p.N, p.L).nunits, s);
                h->stats.launches_reduce++; h->stats.launches_total++;
            }
            if (tmpR) cudaFreeAsync(tmpR, s);
            if (tmpQ) cudaFreeAsync(tmpQ, s);
            if (e != cudaSuccess) return cuda_fail(e, "all-reduce launch");
            return MB200_OK;
        }
        MB200_CUDA(cudaMallocAsync(&pa, ba, s));
        MB200_CUDA(cudaMallocAsync(&pb, bb, s));
        const GettParams &t = cp->gp;
        if (by_permute) {
            e = launch_permute(p.dtype, qa, R, pa, s);
            if (e == cudaSuccess) e = launch_permute(p.dtype, qb, Q, pb, s);
        } else {
            const PackRowOrder oa = pack_row_order(p, 0), ob = pack_row_order(p, 1);
            e = launch_pack_gather(p.dtype, R, t.rowA, t.kA, t.batA, p.M, p.K, Kp, p.L, p.a_kmajor, mixed ? 2 : 1, (float *)pa, s, &oa);
            if (e == cudaSuccess)
                e = launch_pack_gather(p.dtype, Q, t.colB, t.kB, t.batB, p.N, p.K, Kp, p.L, p.b_kmajor, mixed ? 3 : 1, (float *)pb, s, &ob);
        }
        h->stats.launches_permute += 2;
        h->stats.launches_total += 2;
        if (e == cudaSuccess) {
            GettParams g = cp->gp;
            g.C = C;
            g.K = Kp;
            if (sc) g.sc = *sc;
            bool pair = false;
            DistDesc dd{};
            if (dist) {
                dd = *dist;
                // Overlap policy (tools/diag_allreduce.py, globaltimer stamps). Every rank exports (N - 1) / N of its partial C over
                // NVLink whatever N is (117 MB of 134 MB at N = 8: ~0.26 ms at the ~450 GB/s the switch reduction sustains), so the
                // reduction is as long as the sliced GEMM and overlapping always pays: N = 2 1.14 ms overlapped vs 1.35 ms one after the
                // other, N = 8 0.47 vs 0.61. Reserving SMs costs the persistent GEMM whole rounds (512 pair tiles on 74 TPCs are 6.92
                // rounds: ANY reservation makes them 8), so the reducer gets the LARGEST reservation that adds no further round.
                // MB200_DIST_OVERLAP=0 runs the reducer after the GEMM on every SM; MB200_DIST_REDUCER_SMS fixes the reservation.
                static const int overlap = [] { const char *e = getenv("MB200_DIST_OVERLAP"); return e ? atoi(e) : 1; }();
                static const int rsms = [] { const char *e = getenv("MB200_DIST_REDUCER_SMS"); return e ? atoi(e) : 0; }();
                const bool fused_now = (phases & DIST_REDUCE) && (phases & DIST_CONTRACT) && overlap != 0;
                int reserve = 0;
                if (fused_now) {
                    const DistGeometry geo = tf32_dist_geometry(p.dtype, p.M, p.N, p.L);
                    const int64_t walkers = geo.pair ? 74 : 148, tiles = geo.pair ? geo.nunits / 2 : geo.nunits;
                    const int step = geo.pair ? 1 : 2;                       // reserve whole TPCs
                    auto rounds = [&](int64_t r) { return (tiles + (walkers - r) - 1) / (walkers - r); };
                    int64_t r = step * (dd.mc_ws ? 3 : 6);                    // at least 6 / 12 SMs
                    const int64_t cap = step * 16;                           // at most 32 SMs
                    while (r + step <= cap && rounds(r + step) == rounds(r)) r += step;
                    reserve = (int)(geo.pair ? 2 * r : r);
                    if (rsms > 0) reserve = std::min(96, std::max(2, rsms)) & ~1;
                }
                dd.reserve_sms = reserve;
                if (fused_now) {
                    // the reducer first, on the side stream (it must not start before everything the caller enqueued so far: the
                    // previous consumer of C and of the workspace); it occupies its reserved SMs and polls unit flags while the
                    // packs and the GEMM run on the others
                    if ((st = ensure_side_stream(h)) != MB200_OK) return st;
                    MB200_CUDA(cudaEventRecord(h->ev_fork, s));
                    MB200_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
                    MB200_CUDA(launch_tf32_allreduce(p.dtype, cp->gp, dd, h->side));
                    MB200_CUDA(cudaEventRecord(h->ev_join, h->side));
                    h->stats.launches_reduce++; h->stats.launches_total++;
                }
            }
            e = launch_tf32_gemm(p.dtype, pa, pb, g, mixed, s, &pair, dist ? &dd : nullptr);
            h->stats.launches_tcgen05++;
            if (pair) h->stats.launches_tcgen05_pair++;
            if (dist && e == cudaSuccess && (phases & DIST_REDUCE)) {
                if (dd.reserve_sms > 0) {
                    e = cudaStreamWaitEvent(s, h->ev_join, 0);   // join: the reducer has drained every owned unit
                } else {
                    e = launch_tf32_allreduce(p.dtype, cp->gp, dd, s);   // no overlap: after the GEMM, on every SM
                    h->stats.launches_reduce++; h->stats.launches_total++;
                }
            }
        }
        cudaFreeAsync(pa, s);
        cudaFreeAsync(pb, s);
        if (dist && e == cudaSuccess) {
            if (e == cudaSuccess && (phases & DIST_WAIT)) {
                e = launch_dist_wait_done(*dist, tf32_dist_geometry(p.dtype, p.M, p.N, p.L).nunits, s);
                h->stats.launches_reduce++; h->stats.launches_total++;
            }
        }
    } else {
        GettParams g = cp->gp;
        g.A = R; g.B = Q; g.C = C;
        if (sc) g.sc = *sc;
        if (dtype_is_double(p.dtype)) {
            e = launch_gett_f64(p.dtype, g, s);
            h->stats.launches_gett_f64++;
        } else {
            e = launch_simt_f32(p.dtype, g, s);
            h->stats.launches_simt_f32++;
        }
    }
    h->stats.launches_total++;
    if (tmpR) cudaFreeAsync(tmpR, s);
    if (tmpQ) cudaFreeAsync(tmpQ, s);
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return MB200_OK;
}

int make_c_desc(TensorDesc &d, int dtype, int nmode, const int32_t *modes) {
    if (!dtype_valid(dtype)) return fail(MB200_INVALID_ARGUMENT, "C: unknown dtype %d", dtype);
    if (nmode < 0 || nmode > MB200_MAX_MODES)
        return fail(MB200_INVALID_ARGUMENT, "C: nmode %d outside [0, %d]", nmode, MB200_MAX_MODES);
    if (nmode > 0 && !modes) return fail(MB200_INVALID_ARGUMENT, "C: modes must not be NULL when nmode > 0");
    d.dtype = dtype;
    d.n = nmode;
    for (int i = 0; i < nmode; i++) { d.modes[i] = modes[i]; d.ext[i] = 0; d.stride[i] = 0; }
    return MB200_OK;
}

}  // namespace

// unary_einsum with the handle's mutex already held (also the pre-reduce step of binary_einsum)
static int unary_locked(mb200_handle_t h, void *Y, int dtypeY, int nmodeY, const int32_t *modesY, const int64_t *stridesY,
                        const void *X, int dtypeX, int nmodeX, const int32_t *modesX, const int64_t *extentsX,
                        const int64_t *stridesX) {
    if (!dtype_valid(dtypeX) || dtypeY != dtypeX)
        return fail(MB200_INVALID_ARGUMENT, "unary_einsum: eltype(y) must equal eltype(x)");
    if (nmodeX < 0 || nmodeX > MB200_MAX_MODES || nmodeY < 0 || nmodeY > MB200_MAX_MODES)
        return fail(MB200_INVALID_ARGUMENT, "nmode out of range");
    if ((nmodeX > 0 && (!modesX || !extentsX)) || (nmodeY > 0 && !modesY))
        return fail(MB200_INVALID_ARGUMENT, "modes/extents are NULL");
    // distinct labels of x with the summed ("diagonal") stride of all their positions
    struct Lab { int32_t label; int64_t ext, sx, sy; bool in_y; };
    std::vector<Lab> labs;
    int64_t dense = 1;
    for (int i = 0; i < nmodeX; i++) {
        if (extentsX[i] < 0) return fail(MB200_INVALID_ARGUMENT, "x: negative extent");
        const int64_t st = stridesX ? stridesX[i] : dense;
        dense *= extentsX[i];
        bool found = false;
        for (Lab &l : labs)
            if (l.label == modesX[i]) {
                if (l.ext != extentsX[i])
                    return fail(MB200_DIMENSION_MISMATCH, "x: repeated mode %d has extents %lld and %lld", (int)modesX[i],
                                (long long)l.ext, (long long)extentsX[i]);
                l.sx += st;
                found = true;
            }
        if (!found) labs.push_back({modesX[i], extentsX[i], st, 0, false});
    }
    int64_t ydense = 1, total_y = 1;
    for (int i = 0; i < nmodeY; i++) {
        Lab *hit = nullptr;
        for (Lab &l : labs)
            if (l.label == modesY[i]) hit = &l;
        if (!hit) return fail(MB200_INVALID_ARGUMENT, "Output indices must be a subset of input indices (mode %d)", (int)modesY[i]);
        if (hit->in_y) return fail(MB200_INVALID_ARGUMENT, "y: mode %d is repeated", (int)modesY[i]);
        hit->in_y = true;
        hit->sy = stridesY ? stridesY[i] : ydense;
        ydense *= hit->ext;
        total_y *= hit->ext;
    }
    if (total_y == 0) return MB200_OK;
    if (!Y || !X) {
        bool x_empty = false;
        for (const Lab &l : labs) x_empty = x_empty || l.ext == 0;
        if (!Y || !x_empty) return fail(MB200_INVALID_ARGUMENT, "NULL data pointer");
    }
    UnaryParams p{};
    std::vector<Lab> outs, sums;
    for (const Lab &l : labs) {
        if (l.ext == 1) continue;
        (l.in_y ? outs : sums).push_back(l);
    }
    std::stable_sort(outs.begin(), outs.end(), [](const Lab &a, const Lab &b) { return a.sy < b.sy; });
    std::stable_sort(sums.begin(), sums.end(), [](const Lab &a, const Lab &b) { return a.sx < b.sx; });
    p.total_c = 1; p.total_k = 1;
    for (const Lab &l : outs) {
        p.total_c *= l.ext;
        if (p.nc > 0 && l.sx == p.c_sx[p.nc - 1] * p.c_ext[p.nc - 1] && l.sy == p.c_sy[p.nc - 1] * p.c_ext[p.nc - 1]) {
            p.c_ext[p.nc - 1] *= l.ext;
        } else {
            p.c_ext[p.nc] = l.ext; p.c_sx[p.nc] = l.sx; p.c_sy[p.nc] = l.sy; p.nc++;
        }
    }
    for (const Lab &l : sums) {
        p.total_k *= l.ext;
        if (p.nk > 0 && l.sx == p.k_sx[p.nk - 1] * p.k_ext[p.nk - 1]) {
            p.k_ext[p.nk - 1] *= l.ext;
        } else {
            p.k_ext[p.nk] = l.ext; p.k_sx[p.nk] = l.sx; p.nk++;
        }
    }
    MB200_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    // no reduction, no diagonal, dense operands: a plain permutation -> K1
    if (p.nk == 0 && (int)labs.size() == nmodeX && !stridesX && !stridesY && nmodeX == nmodeY) {
        PermuteParams q{};
        q.total = total_y;
        for (int i = 0; i < nmodeX; i++) {
            if (extentsX[i] == 1) continue;
            int64_t ds = 0;
            for (const Lab &l : labs)
                if (l.label == modesX[i]) ds = l.sy;
            if (q.n > 0 && ds == q.dst_stride[q.n - 1] * q.ext[q.n - 1]) {
                q.ext[q.n - 1] *= extentsX[i];
            } else {
                q.ext[q.n] = extentsX[i]; q.dst_stride[q.n] = ds; q.n++;
            }
        }
        MB200_CUDA(launch_permute(dtypeX, q, X, Y, s));
        h->stats.launches_permute++;
        h->stats.launches_total++;
        return MB200_OK;
    }
    const int nsplit = unary_nsplit(p);
    void *part = nullptr;
    if (nsplit > 1) MB200_CUDA(cudaMallocAsync(&part, (size_t)nsplit * p.total_c * dtype_size(dtypeX), s));
    cudaError_t e = launch_unary(dtypeX, p, X, Y, part, nsplit, s);
    if (part) cudaFreeAsync(part, s);
    if (e != cudaSuccess) return cuda_fail(e, "unary_einsum launch");
    h->stats.launches_unary += nsplit > 1 ? 2 : 1;
    h->stats.launches_total += nsplit > 1 ? 2 : 1;
    return MB200_OK;
}

// binary_einsum with modes that one operand carries and C does not ("dangling": cuTENSOR / OMEinsum semantics - they are
// summed; ext/MuscleCUDAExt.jl:30-38, ext/MuscleOMEinsumExt.jl:40-59). Extent-1 dangling modes are dropped from the
// descriptor. Launch-bound sizes (<= 2^20 MACs with the dangling extents counted) FOLD the sum into the contraction: the
// mode is handed to the planner as a summed mode that the other operand carries with stride 0 (a broadcast), one direct
// kernel launch, no temporary. Larger operands are pre-reduced by one HBM-bound unary_einsum pass (|X| read once) and the
// contraction then runs on the smaller tensor - ext(x) times fewer tensor-core flops than folding.
static int contract_entry(mb200_handle_t h, void *C, TensorDesc &dC, const int64_t *stridesC, const void *A, TensorDesc dA,
                          const void *B, TensorDesc dB, const ScatterDesc *sc = nullptr) {
    std::vector<int> da = dangling_modes(dA, dB, dC), db = dangling_modes(dB, dA, dC);
    if (da.empty() && db.empty()) return contract_device(h, C, dC, stridesC, A, dA, B, dB, sc);
    // extent-1 dangling modes never move an address
    auto drop_unit = [](TensorDesc &T, std::vector<int> &d) {
        std::vector<int> unit;
        for (int i : d)
            if (T.ext[i] == 1) unit.push_back(i);
        if (unit.empty()) return;
        T = without_modes(T, unit, false);
    };
    drop_unit(dA, da);
    drop_unit(dB, db);
    da = dangling_modes(dA, dB, dC);
    db = dangling_modes(dB, dA, dC);
    if (da.empty() && db.empty()) return contract_device(h, C, dC, stridesC, A, dA, B, dB, sc);

    double macs = 1.0;
    {
        std::vector<int32_t> seen;
        for (const TensorDesc *t : {&dA, &dB})
            for (int i = 0; i < t->n; i++)
                if (std::find(seen.begin(), seen.end(), t->modes[i]) == seen.end()) {
                    seen.push_back(t->modes[i]);
                    macs *= (double)t->ext[i];
                }
    }
    const bool room = dA.n + (int)db.size() <= MB200_MAX_MODES && dB.n + (int)da.size() <= MB200_MAX_MODES;
    if (macs <= (double)(1 << 20) && !sc && room) {
        TensorDesc fa = dA, fb = dB;
        for (int i : da) { fb.modes[fb.n] = dA.modes[i]; fb.ext[fb.n] = dA.ext[i]; fb.stride[fb.n] = 0; fb.n++; }
        for (int i : db) { fa.modes[fa.n] = dB.modes[i]; fa.ext[fa.n] = dB.ext[i]; fa.stride[fa.n] = 0; fa.n++; }
        const int saved = h->forced_path;
        h->forced_path = MB200_PATH_DIRECT;
        int st = contract_device(h, C, dC, stridesC, A, fa, B, fb, nullptr);
        h->forced_path = saved;
        return st;
    }
    cudaStream_t s = h->stream;
    void *tmpA = nullptr, *tmpB = nullptr;
    int st = MB200_OK;
    auto prereduce = [&](const void *&X, TensorDesc &dX, const std::vector<int> &drop, void *&tmp) -> int {
        if (drop.empty()) return MB200_OK;
        TensorDesc dY = without_modes(dX, drop, true);
        if (dY.numel() == 0) { dX = dY; return MB200_OK; }
        cudaError_t e = cudaSetDevice(h->device);
        if (e == cudaSuccess) e = cudaMallocAsync(&tmp, (size_t)dY.numel() * dtype_size(dY.dtype), s);
        if (e != cudaSuccess) return cuda_fail(e, "pre-reduce workspace");
        int r = unary_locked(h, tmp, dY.dtype, dY.n, dY.modes, nullptr, X, dX.dtype, dX.n, dX.modes, dX.ext, dX.stride);
        if (r != MB200_OK) return r;
        X = tmp;
        dX = dY;
        return MB200_OK;
    };
    st = prereduce(A, dA, da, tmpA);
    if (st == MB200_OK) st = prereduce(B, dB, db, tmpB);
    if (st == MB200_OK) st = contract_device(h, C, dC, stridesC, A, dA, B, dB, sc);
    if (tmpA) cudaFreeAsync(tmpA, s);
    if (tmpB) cudaFreeAsync(tmpB, s);
    return st;
}

// descriptors as the planner sees them: dangling modes already summed away (introspection / host-side validation)
static void strip_dangling_for_plan(TensorDesc &dA, TensorDesc &dB, const TensorDesc &dC) {
    std::vector<int> da = dangling_modes(dA, dB, dC), db = dangling_modes(dB, dA, dC);
    if (!da.empty()) dA = without_modes(dA, da, true);
    if (!db.empty()) dB = without_modes(dB, db, true);
}

extern "C" {

int mb200_version(void) { return 100; }

const char *mb200_last_error_string(void) { return last_error().c_str(); }

int mb200_device_count(int *count) {
    if (!count) return fail(MB200_INVALID_ARGUMENT, "count is NULL");
    *count = 0;
    MB200_CUDA(cudaGetDeviceCount(count));
    return MB200_OK;
}

int mb200_create(mb200_handle_t *handle, int device) {
    if (!handle) return fail(MB200_INVALID_ARGUMENT, "handle is NULL");
    *handle = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(MB200_CUDA_ERROR, "no usable CUDA device (%s); libmuscle_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(MB200_INVALID_ARGUMENT, "device %d outside [0, %d)", device, n);
    MB200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MB200_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(MB200_NOT_SUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    device, prop.major, prop.minor);
    MB200_CUDA(gett_configure());
    MB200_CUDA(permute_configure());
    MB200_CUDA(tf32_configure());
    // keep stream-ordered temporaries cached in the pool instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    auto *h = new mb200_handle_s();
    h->device = device;
    *handle = h;
    return MB200_OK;
}

int mb200_destroy(mb200_handle_t h) {
    MB200_CHECK_HANDLE(h);
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->svd_info) cudaFree(h->svd_info);
    if (h->side) { cudaStreamDestroy(h->side); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
    delete h;   // cached plans free their tables unless a captured graph still holds them
    return MB200_OK;
}

int mb200_set_stream(mb200_handle_t h, void *cuda_stream) {
    MB200_CHECK_HANDLE(h);
    std::lock_guard<std::mutex> lk(h->mu);
    h->stream = (cudaStream_t)cuda_stream;
    return MB200_OK;
}

int mb200_stream_sync(mb200_handle_t h) {
    MB200_CHECK_HANDLE(h);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaStreamSynchronize(h->stream));
    return MB200_OK;
}

int mb200_set_path(mb200_handle_t h, int path) {
    MB200_CHECK_HANDLE(h);
    if (path < MB200_PATH_AUTO || path > MB200_PATH_TCGEN05_TF32)
        return fail(MB200_INVALID_ARGUMENT, "unknown path %d", path);
    std::lock_guard<std::mutex> lk(h->mu);
    h->forced_path = path;
    return MB200_OK;
}

int mb200_set_compute_type(mb200_handle_t h, int compute_type) {
    MB200_CHECK_HANDLE(h);
    if (compute_type < MB200_COMPUTE_DEFAULT || compute_type > MB200_COMPUTE_3XTF32)
        return fail(MB200_INVALID_ARGUMENT, "unknown compute type %d", compute_type);
    std::lock_guard<std::mutex> lk(h->mu);
    h->compute_type = compute_type;
    return MB200_OK;
}

int mb200_get_compute_type(mb200_handle_t h, int *compute_type) {
    MB200_CHECK_HANDLE(h);
    if (!compute_type) return fail(MB200_INVALID_ARGUMENT, "compute_type is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    *compute_type = h->compute_type;
    return MB200_OK;
}

int mb200_malloc(mb200_handle_t h, void **dptr, size_t bytes) {
    MB200_CHECK_HANDLE(h);
    if (!dptr) return fail(MB200_INVALID_ARGUMENT, "dptr is NULL");
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return MB200_OK;
}

int mb200_free(mb200_handle_t h, void *dptr) {
    MB200_CHECK_HANDLE(h);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaFree(dptr));
    return MB200_OK;
}

int mb200_host_alloc(void **hptr, size_t bytes) {
    if (!hptr) return fail(MB200_INVALID_ARGUMENT, "hptr is NULL");
    MB200_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return MB200_OK;
}

int mb200_host_free(void *hptr) {
    MB200_CUDA(cudaFreeHost(hptr));
    return MB200_OK;
}

int mb200_memcpy_h2d(mb200_handle_t h, void *dst, const void *src, size_t bytes) {
    MB200_CHECK_HANDLE(h);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return MB200_OK;
}

int mb200_memcpy_d2h(mb200_handle_t h, void *dst, const void *src, size_t bytes) {
    MB200_CHECK_HANDLE(h);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return MB200_OK;
}

int mb200_memset(mb200_handle_t h, void *dptr, int value, size_t bytes) {
    MB200_CHECK_HANDLE(h);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaMemsetAsync(dptr, value, bytes, h->stream));
    return MB200_OK;
}

int mb200_binary_einsum(mb200_handle_t h, void *C, int dtypeC, int nmodeC, const int32_t *modesC,
                        const int64_t *stridesC, const void *A, int dtypeA, int nmodeA, const int32_t *modesA,
                        const int64_t *extentsA, const int64_t *stridesA, const void *B, int dtypeB, int nmodeB,
                        const int32_t *modesB, const int64_t *extentsB, const int64_t *stridesB) {
    MB200_CHECK_HANDLE(h);
    TensorDesc dA, dB, dC;
    int st;
    if ((st = make_desc(dA, dtypeA, nmodeA, modesA, extentsA, stridesA, "A")) != MB200_OK) return st;
    if ((st = make_desc(dB, dtypeB, nmodeB, modesB, extentsB, stridesB, "B")) != MB200_OK) return st;
    if ((st = make_c_desc(dC, dtypeC, nmodeC, modesC)) != MB200_OK) return st;
    std::lock_guard<std::mutex> lk(h->mu);
    return contract_entry(h, C, dC, stridesC, A, dA, B, dB);
}

int mb200_binary_einsum_host(mb200_handle_t h, void *C, int dtypeC, int nmodeC, const int32_t *modesC,
                             const void *A, int dtypeA, int nmodeA, const int32_t *modesA,
                             const int64_t *extentsA, const void *B, int dtypeB, int nmodeB,
                             const int32_t *modesB, const int64_t *extentsB) {
    MB200_CHECK_HANDLE(h);
    TensorDesc dA, dB, dC;
    int st;
    if ((st = make_desc(dA, dtypeA, nmodeA, modesA, extentsA, nullptr, "A")) != MB200_OK) return st;
    if ((st = make_desc(dB, dtypeB, nmodeB, modesB, extentsB, nullptr, "B")) != MB200_OK) return st;
    if ((st = make_c_desc(dC, dtypeC, nmodeC, modesC)) != MB200_OK) return st;
    std::lock_guard<std::mutex> lk(h->mu);
    {   // validate before touching the device so argument errors do not depend on a GPU
        Plan probe;
        TensorDesc tmp = dC, pa = dA, pb = dB;
        strip_dangling_for_plan(pa, pb, dC);
        if ((st = make_plan(pa, pb, tmp, nullptr, h->forced_path, probe)) != MB200_OK) return st;
        dC = tmp;
    }
    MB200_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t bA = (size_t)dA.numel() * dtype_size(dA.dtype);
    const size_t bB = (size_t)dB.numel() * dtype_size(dB.dtype);
    const size_t bC = (size_t)dC.numel() * dtype_size(dC.dtype);
    if (bC == 0) return MB200_OK;
    if (!C || (!A && bA) || (!B && bB)) return fail(MB200_INVALID_ARGUMENT, "NULL data pointer");
    void *dAp = nullptr, *dBp = nullptr, *dCp = nullptr;
    MB200_CUDA(cudaMallocAsync(&dAp, bA ? bA : 1, s));
    MB200_CUDA(cudaMallocAsync(&dBp, bB ? bB : 1, s));
    MB200_CUDA(cudaMallocAsync(&dCp, bC, s));
    if (bA) MB200_CUDA(cudaMemcpyAsync(dAp, A, bA, cudaMemcpyHostToDevice, s));
    if (bB) MB200_CUDA(cudaMemcpyAsync(dBp, B, bB, cudaMemcpyHostToDevice, s));
    st = contract_entry(h, dCp, dC, nullptr, dAp, dA, dBp, dB);
    if (st == MB200_OK) {
        cudaError_t e = cudaMemcpyAsync(C, dCp, bC, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) st = cuda_fail(e, "device-to-host copy of C");
    }
    cudaFreeAsync(dAp, s);
    cudaFreeAsync(dBp, s);
    cudaFreeAsync(dCp, s);
    return st;
}

int mb200_plan_describe(int dtypeC, int nmodeC, const int32_t *modesC, const int64_t *stridesC, int dtypeA,
                        int nmodeA, const int32_t *modesA, const int64_t *extentsA, const int64_t *stridesA,
                        int dtypeB, int nmodeB, const int32_t *modesB, const int64_t *extentsB,
                        const int64_t *stridesB, mb200_plan_info_t *info) {
    if (!info) return fail(MB200_INVALID_ARGUMENT, "info is NULL");
    TensorDesc dA, dB, dC;
    int st;
    if ((st = make_desc(dA, dtypeA, nmodeA, modesA, extentsA, stridesA, "A")) != MB200_OK) return st;
    if ((st = make_desc(dB, dtypeB, nmodeB, modesB, extentsB, stridesB, "B")) != MB200_OK) return st;
    if ((st = make_c_desc(dC, dtypeC, nmodeC, modesC)) != MB200_OK) return st;
    Plan plan;
    strip_dangling_for_plan(dA, dB, dC);   // the plan of the contraction that runs after the dangling modes are summed
    if ((st = make_plan(dA, dB, dC, stridesC, MB200_PATH_AUTO, plan)) != MB200_OK) return st;
    fill_info(plan, info);
    return MB200_OK;
}

int mb200_permute(mb200_handle_t h, void *dst, const void *src, int dtype, int nmode, const int64_t *extents,
                  const int32_t *perm, uint32_t flags) {
    MB200_CHECK_HANDLE(h);
    if (!dtype_valid(dtype)) return fail(MB200_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if (nmode < 0 || nmode > MB200_MAX_MODES) return fail(MB200_INVALID_ARGUMENT, "nmode %d out of range", nmode);
    if (nmode > 0 && (!extents || !perm)) return fail(MB200_INVALID_ARGUMENT, "extents/perm are NULL");
    bool seen[MB200_MAX_MODES] = {false};
    for (int d = 0; d < nmode; d++) {
        if (perm[d] < 0 || perm[d] >= nmode || seen[perm[d]])
            return fail(MB200_INVALID_ARGUMENT, "perm is not a permutation of 0..%d", nmode - 1);
        seen[perm[d]] = true;
        if (extents[d] < 0) return fail(MB200_INVALID_ARGUMENT, "negative extent");
    }
    if ((flags & MB200_PERMUTE_PLANAR) && !dtype_is_complex(dtype))
        return fail(MB200_INVALID_ARGUMENT, "planar output needs a complex dtype");
    // destination stride of every source mode
    int64_t dstride_src[MB200_MAX_MODES];
    int64_t st = 1, total = 1;
    for (int d = 0; d < nmode; d++) {
        dstride_src[perm[d]] = st;
        st *= extents[perm[d]];
    }
    for (int i = 0; i < nmode; i++) total *= extents[i];
    if (total == 0) return MB200_OK;
    if (!dst || !src) return fail(MB200_INVALID_ARGUMENT, "NULL data pointer");
    // canonical form: drop extent-1 modes, merge modes adjacent in both layouts
    PermuteParams q{};
    for (int i = 0; i < nmode; i++) {
        if (extents[i] == 1) continue;
        if (q.n > 0 && dstride_src[i] == q.dst_stride[q.n - 1] * q.ext[q.n - 1]) {
            q.ext[q.n - 1] *= extents[i];
        } else {
            q.ext[q.n] = extents[i];
            q.dst_stride[q.n] = dstride_src[i];
            q.n++;
        }
    }
    q.total = total;
    q.plane_stride = (flags & MB200_PERMUTE_PLANAR) ? total : 0;
    q.tma = (flags & MB200_PERMUTE_TMA) ? 1 : ((flags & MB200_PERMUTE_NO_TMA) ? -1 : 0);
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(launch_permute(dtype, q, src, dst, h->stream));
    h->stats.launches_permute++;
    h->stats.launches_total++;
    return MB200_OK;
}

int mb200_unary_einsum(mb200_handle_t h, void *Y, int dtypeY, int nmodeY, const int32_t *modesY, const int64_t *stridesY,
                       const void *X, int dtypeX, int nmodeX, const int32_t *modesX, const int64_t *extentsX,
                       const int64_t *stridesX) {
    MB200_CHECK_HANDLE(h);
    std::lock_guard<std::mutex> lk(h->mu);
    return unary_locked(h, Y, dtypeY, nmodeY, modesY, stridesY, X, dtypeX, nmodeX, modesX, extentsX, stridesX);
}

int mb200_hadamard(mb200_handle_t h, void *Cp, int dtypeC, const void *A, int dtypeA, int nmodeA, const int32_t *modesA,
                   const int64_t *extentsA, const void *B, int dtypeB, int nmodeB, const int32_t *modesB,
                   const int64_t *extentsB) {
    MB200_CHECK_HANDLE(h);
    if (!dtype_valid(dtypeA) || !dtype_valid(dtypeB)) return fail(MB200_INVALID_ARGUMENT, "unknown dtype");
    if (dtypeC != dtype_promote(dtypeA, dtypeB))
        return fail(MB200_INVALID_ARGUMENT, "eltype(c) = %s must be promote_eltype(a, b) = %s", dtype_name(dtypeC),
                    dtype_name(dtype_promote(dtypeA, dtypeB)));
    if (nmodeA < 0 || nmodeA > MB200_MAX_MODES || nmodeB < 0 || nmodeB > MB200_MAX_MODES)
        return fail(MB200_INVALID_ARGUMENT, "nmode out of range");
    if ((nmodeA > 0 && (!modesA || !extentsA)) || (nmodeB > 0 && (!modesB || !extentsB)))
        return fail(MB200_INVALID_ARGUMENT, "modes/extents are NULL");
    for (int i = 0; i < nmodeA; i++)
        for (int j = 0; j < i; j++)
            if (modesA[i] == modesA[j]) return fail(MB200_INVALID_ARGUMENT, "a: mode %d is repeated", (int)modesA[i]);
    // stride of every mode of a inside b (0 = b is broadcast along it)
    int64_t sb_of_a[MB200_MAX_MODES] = {0};
    int64_t bdense = 1, total_b = 1;
    for (int j = 0; j < nmodeB; j++) {
        for (int k = 0; k < j; k++)
            if (modesB[j] == modesB[k]) return fail(MB200_INVALID_ARGUMENT, "b: mode %d is repeated", (int)modesB[j]);
        int ia = -1;
        for (int i = 0; i < nmodeA; i++)
            if (modesA[i] == modesB[j]) ia = i;
        if (ia < 0) return fail(MB200_INVALID_ARGUMENT, "inds(b) must be a subset of inds(a) (mode %d)", (int)modesB[j]);
        if (extentsA[ia] != extentsB[j])
            return fail(MB200_DIMENSION_MISMATCH, "mode %d has extent %lld in a and %lld in b", (int)modesB[j],
                        (long long)extentsA[ia], (long long)extentsB[j]);
        sb_of_a[ia] = bdense;
        bdense *= extentsB[j];
        total_b *= extentsB[j];
    }
    HadamardParams p{};
    p.total = 1;
    for (int i = 0; i < nmodeA; i++) {
        if (extentsA[i] < 0) return fail(MB200_INVALID_ARGUMENT, "negative extent");
        p.total *= extentsA[i];
        if (extentsA[i] == 1) continue;
        if (p.n > 0 && sb_of_a[i] == p.sb[p.n - 1] * p.ext[p.n - 1]) {
            p.ext[p.n - 1] *= extentsA[i];
        } else {
            p.ext[p.n] = extentsA[i]; p.sb[p.n] = sb_of_a[i]; p.n++;
        }
    }
    if (p.total == 0) return MB200_OK;
    if (p.n == 0) { p.n = 1; p.ext[0] = 1; p.sb[0] = 0; }
    if (!Cp || !A || !B) return fail(MB200_INVALID_ARGUMENT, "NULL data pointer");
    if (Cp == A && dtypeA != dtypeC) return fail(MB200_INVALID_ARGUMENT, "c may alias a only when eltype(a) == eltype(c)");
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    void *tmpA = nullptr, *tmpB = nullptr;
    if (dtypeA != dtypeC) {
        MB200_CUDA(cudaMallocAsync(&tmpA, (size_t)p.total * dtype_size(dtypeC), s));
        MB200_CUDA(launch_convert(dtypeC, tmpA, dtypeA, A, p.total, s));
        h->stats.launches_convert++; h->stats.launches_total++;
        A = tmpA;
    }
    if (dtypeB != dtypeC) {
        MB200_CUDA(cudaMallocAsync(&tmpB, (size_t)std::max<int64_t>(total_b, 1) * dtype_size(dtypeC), s));
        MB200_CUDA(launch_convert(dtypeC, tmpB, dtypeB, B, total_b, s));
        h->stats.launches_convert++; h->stats.launches_total++;
        B = tmpB;
    }
    const int vec = (int)(16 / dtype_size(dtypeC));
    p.b_vec_aligned = (((uintptr_t)B) & 15) == 0;
    for (int i = 1; i < p.n; i++) p.b_vec_aligned = p.b_vec_aligned && p.sb[i] % vec == 0;
    cudaError_t e = launch_hadamard(dtypeC, p, A, B, Cp, s);
    if (tmpA) cudaFreeAsync(tmpA, s);
    if (tmpB) cudaFreeAsync(tmpB, s);
    if (e != cudaSuccess) return cuda_fail(e, "hadamard launch");
    h->stats.launches_hadamard++;
    h->stats.launches_total++;
    return MB200_OK;
}

int mb200_svd_thin(mb200_handle_t h, void *U, void *S, void *Vt, const void *A, int dtype, int64_t rows, int64_t cols,
                   double tol, int max_sweeps) {
    MB200_CHECK_HANDLE(h);
    if (!dtype_valid(dtype)) return fail(MB200_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if (rows < 0 || cols < 0 || rows >= ((int64_t)1 << 31) || cols >= ((int64_t)1 << 31))
        return fail(MB200_INVALID_ARGUMENT, "svd: bad shape %lld x %lld", (long long)rows, (long long)cols);
    const int64_t k = std::min(rows, cols), m = std::max(rows, cols);
    if (k == 0) return MB200_OK;
    if (!U || !S || !Vt || !A) return fail(MB200_INVALID_ARGUMENT, "NULL data pointer");
    if (tol <= 0) tol = std::sqrt((double)m) * (dtype_is_double(dtype) ? 2.220446049250313e-16 : 1.1920929e-7);
    if (max_sweeps <= 0) max_sweeps = 30;
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    void *G = nullptr, *V = nullptr;
    const size_t esz = dtype_size(dtype);
    if (!h->svd_info) MB200_CUDA(cudaMalloc((void **)&h->svd_info, 8 * sizeof(int)));
    MB200_CUDA(cudaMallocAsync(&G, (size_t)m * k * esz, s));
    MB200_CUDA(cudaMallocAsync(&V, (size_t)k * k * esz, s));
    cudaError_t e = launch_svd(dtype, A, (int)rows, (int)cols, U, S, Vt, G, V, h->svd_info, tol, max_sweeps, s);
    cudaFreeAsync(G, s);
    cudaFreeAsync(V, s);
    if (e != cudaSuccess) return cuda_fail(e, "svd launch");
    h->stats.launches_svd++;
    h->stats.launches_total++;
    return MB200_OK;
}

int mb200_svd_last_info(mb200_handle_t h, int *sweeps, int *converged, int *completed_columns) {
    MB200_CHECK_HANDLE(h);
    std::lock_guard<std::mutex> lk(h->mu);
    int info[8] = {0, 0, 0, 1, 0, 0, 0, 0};
    if (h->svd_info) {
        MB200_CUDA(cudaSetDevice(h->device));
        MB200_CUDA(cudaMemcpyAsync(info, h->svd_info, sizeof info, cudaMemcpyDeviceToHost, h->stream));
        MB200_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (sweeps) *sweeps = info[1];
    if (converged) *converged = info[3];
    if (completed_columns) *completed_columns = info[4];
    return MB200_OK;
}

int mb200_qr_thin(mb200_handle_t h, void *Q, void *R, const void *A, int dtype, int64_t rows, int64_t cols) {
    MB200_CHECK_HANDLE(h);
    if (!dtype_valid(dtype)) return fail(MB200_INVALID_ARGUMENT, "unknown dtype %d", dtype);
    if (rows < 0 || cols < 0 || rows >= ((int64_t)1 << 31) || cols >= ((int64_t)1 << 31))
        return fail(MB200_INVALID_ARGUMENT, "qr: bad shape %lld x %lld", (long long)rows, (long long)cols);
    const int64_t k = std::min(rows, cols);
    if (k == 0) return MB200_OK;
    if (!Q || !R || !A) return fail(MB200_INVALID_ARGUMENT, "NULL data pointer");
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t esz = dtype_size(dtype);
    void *W = nullptr, *rd = nullptr, *tau = nullptr;
    MB200_CUDA(cudaMallocAsync(&W, (size_t)rows * cols * esz, s));
    MB200_CUDA(cudaMallocAsync(&rd, (size_t)k * esz, s));
    MB200_CUDA(cudaMallocAsync(&tau, (size_t)k * sizeof(double), s));
    cudaError_t e = launch_qr(dtype, A, (int)rows, (int)cols, Q, R, W, rd, tau, s);
    cudaFreeAsync(W, s);
    cudaFreeAsync(rd, s);
    cudaFreeAsync(tau, s);
    if (e != cudaSuccess) return cuda_fail(e, "qr launch");
    h->stats.launches_svd++;
    h->stats.launches_total++;
    return MB200_OK;
}

int mb200_shard_plan(int nmodeC, const int32_t *modesC, int nmodeA, const int32_t *modesA,
                     const int64_t *extentsA, int nmodeB, const int32_t *modesB, const int64_t *extentsB,
                     int nranks, int rank, int prefer_sum, mb200_shard_info_t *info) {
    if (!info) return fail(MB200_INVALID_ARGUMENT, "info is NULL");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MB200_INVALID_ARGUMENT, "bad rank %d / %d", rank, nranks);
    if (nmodeA < 0 || nmodeB < 0 || nmodeC < 0 || nmodeA > MB200_MAX_MODES || nmodeB > MB200_MAX_MODES ||
        nmodeC > MB200_MAX_MODES)
        return fail(MB200_INVALID_ARGUMENT, "nmode out of range");
    std::memset(info, 0, sizeof *info);
    info->kind = MB200_SHARD_NONE;
    info->mode = -1;
    auto find = [](int n, const int32_t *m, int32_t x) {
        for (int i = 0; i < n; i++)
            if (m[i] == x) return i;
        return -1;
    };
    auto set = [&](int kind, int32_t mode, int64_t ext) {
        info->kind = kind;
        info->mode = mode;
        info->begin = ext * rank / nranks;
        info->end = ext * (rank + 1) / nranks;
        info->needs_allreduce = kind == MB200_SHARD_SUM;
    };
    if (nranks == 1) return MB200_OK;
    if (prefer_sum) {
        // slowest summed mode of A with extent >= nranks
        for (int i = nmodeA - 1; i >= 0; i--) {
            if (find(nmodeB, modesB, modesA[i]) >= 0 && find(nmodeC, modesC, modesA[i]) < 0 &&
                extentsA[i] >= nranks) {
                set(MB200_SHARD_SUM, modesA[i], extentsA[i]);
                return MB200_OK;
            }
        }
    }
    // slowest mode of C with extent >= nranks: free (one operand) or batch (both)
    for (int i = nmodeC - 1; i >= 0; i--) {
        int ia = find(nmodeA, modesA, modesC[i]), ib = find(nmodeB, modesB, modesC[i]);
        if (ia < 0 && ib < 0) return fail(MB200_INVALID_ARGUMENT, "mode %d of C is in neither operand", (int)modesC[i]);
        int64_t ext = ia >= 0 ? extentsA[ia] : extentsB[ib];
        if (ext >= nranks) {
            set((ia >= 0 && ib >= 0) ? MB200_SHARD_BATCH : MB200_SHARD_FREE, modesC[i], ext);
            return MB200_OK;
        }
    }
    if (!prefer_sum) {
        for (int i = nmodeA - 1; i >= 0; i--) {
            if (find(nmodeB, modesB, modesA[i]) >= 0 && find(nmodeC, modesC, modesA[i]) < 0 &&
                extentsA[i] >= nranks) {
                set(MB200_SHARD_SUM, modesA[i], extentsA[i]);
                return MB200_OK;
            }
        }
    }
    return MB200_OK;  // replicas only
}

int mb200_ipc_export(mb200_handle_t h, void *dptr, unsigned char *handle_out) {
    MB200_CHECK_HANDLE(h);
    if (!dptr || !handle_out) return fail(MB200_INVALID_ARGUMENT, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MB200_IPC_HANDLE_BYTES, "IPC handle size");
    MB200_CUDA(cudaSetDevice(h->device));
    cudaIpcMemHandle_t ih;
    MB200_CUDA(cudaIpcGetMemHandle(&ih, dptr));
    std::memcpy(handle_out, &ih, sizeof ih);
    return MB200_OK;
}

int mb200_ipc_import(mb200_handle_t h, const unsigned char *handle_in, void **peer_ptr) {
    MB200_CHECK_HANDLE(h);
    if (!handle_in || !peer_ptr) return fail(MB200_INVALID_ARGUMENT, "NULL argument");
    MB200_CUDA(cudaSetDevice(h->device));
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, handle_in, sizeof ih);
    MB200_CUDA(cudaIpcOpenMemHandle(peer_ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    return MB200_OK;
}

int mb200_ipc_release(mb200_handle_t h, void *peer_ptr) {
    MB200_CHECK_HANDLE(h);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return MB200_OK;
}

int mb200_binary_einsum_scatter(mb200_handle_t h, int dtypeC, int nmodeC, const int32_t *modesC, const void *A,
                                int dtypeA, int nmodeA, const int32_t *modesA, const int64_t *extentsA,
                                const int64_t *stridesA, const void *B, int dtypeB, int nmodeB, const int32_t *modesB,
                                const int64_t *extentsB, const int64_t *stridesB, void *const *staging, int nranks,
                                int rank, int slab_shift) {
    MB200_CHECK_HANDLE(h);
    if (!staging || nranks < 1 || nranks > MB200_MAX_PEERS || rank < 0 || rank >= nranks || slab_shift < 0 || slab_shift > 40)
        return fail(MB200_INVALID_ARGUMENT, "bad staging / nranks %d / rank %d / slab_shift %d", nranks, rank, slab_shift);
    ScatterDesc sc{};
    for (int r = 0; r < nranks; r++) {
        if (!staging[r]) return fail(MB200_INVALID_ARGUMENT, "staging[%d] is NULL", r);
        sc.peer[r] = staging[r];
    }
    sc.nranks = nranks; sc.rank = rank; sc.shift = slab_shift;
    TensorDesc dA, dB, dC;
    int st;
    if ((st = make_desc(dA, dtypeA, nmodeA, modesA, extentsA, stridesA, "A")) != MB200_OK) return st;
    if ((st = make_desc(dB, dtypeB, nmodeB, modesB, extentsB, stridesB, "B")) != MB200_OK) return st;
    if ((st = make_c_desc(dC, dtypeC, nmodeC, modesC)) != MB200_OK) return st;
    std::lock_guard<std::mutex> lk(h->mu);
    return contract_entry(h, nullptr, dC, nullptr, A, dA, B, dB, &sc);
}

static int make_dist(DistDesc &d, const mb200_comm_t *comm) {
    if (!comm) return fail(MB200_INVALID_ARGUMENT, "comm is NULL");
    if (comm->nranks < 1 || comm->nranks > MB200_MAX_PEERS || comm->rank < 0 || comm->rank >= comm->nranks)
        return fail(MB200_INVALID_ARGUMENT, "bad comm: rank %d of %d", comm->rank, comm->nranks);
    if (comm->epoch < 1) return fail(MB200_INVALID_ARGUMENT, "comm.epoch must be >= 1 and increase from call to call");
    d = DistDesc{};
    d.nranks = comm->nranks; d.rank = comm->rank; d.epoch = comm->epoch;
    for (int r = 0; r < comm->nranks; r++) {
        if (!comm->ws[r] || !comm->c[r] || !comm->flags[r]) return fail(MB200_INVALID_ARGUMENT, "comm: NULL buffer of rank %d", r);
        d.ws[r] = comm->ws[r]; d.c[r] = comm->c[r]; d.flags[r] = (int *)comm->flags[r];
    }
    if ((comm->mc_ws == nullptr) != (comm->mc_c == nullptr))
        return fail(MB200_INVALID_ARGUMENT, "comm: mc_ws and mc_c must both be set or both be NULL");
    d.mc_ws = comm->mc_ws; d.mc_c = comm->mc_c;
    return MB200_OK;
}

int mb200_allreduce_workspace(mb200_handle_t h, int dtypeC, int nmodeC, const int32_t *modesC,
                              int dtypeA, int nmodeA, const int32_t *modesA, const int64_t *extentsA,
                              int dtypeB, int nmodeB, const int32_t *modesB, const int64_t *extentsB,
                              int nranks, size_t *ws_bytes, size_t *flag_bytes) {
    MB200_CHECK_HANDLE(h);
    if (!ws_bytes || !flag_bytes || nranks < 1 || nranks > MB200_MAX_PEERS) return fail(MB200_INVALID_ARGUMENT, "bad arguments");
    TensorDesc dA, dB, dC;
    int st;
    if ((st = make_desc(dA, dtypeA, nmodeA, modesA, extentsA, nullptr, "A")) != MB200_OK) return st;
    if ((st = make_desc(dB, dtypeB, nmodeB, modesB, extentsB, nullptr, "B")) != MB200_OK) return st;
    if ((st = make_c_desc(dC, dtypeC, nmodeC, modesC)) != MB200_OK) return st;
    std::lock_guard<std::mutex> lk(h->mu);
    Plan plan;
    if ((st = make_plan(dA, dB, dC, nullptr, h->forced_path, plan)) != MB200_OK) return st;
    if (h->compute_type == MB200_COMPUTE_FP32 || plan.path != MB200_PATH_TCGEN05_TF32 || plan.empty_output)
        return fail(MB200_NOT_SUPPORTED, "the fused all-reduce runs on the ComplexF32 / Float32 tcgen05 path (this contraction is "
                                         "planned on path %d); use an NCCL all-reduce of the partial outputs", plan.path);
    const DistGeometry geo = tf32_dist_geometry(plan.dtype, plan.M, plan.N, plan.L);
    *ws_bytes = (size_t)geo.nunits * geo.unit_elems * dtype_size(plan.dtype);
    *flag_bytes = ((size_t)geo.nunits * nranks + nranks + 1) * sizeof(int);
    return MB200_OK;
}

int mb200_binary_einsum_allreduce(mb200_handle_t h, int dtypeC, int nmodeC, const int32_t *modesC,
                                  const void *A, int dtypeA, int nmodeA, const int32_t *modesA, const int64_t *extentsA,
                                  const int64_t *stridesA,
                                  const void *B, int dtypeB, int nmodeB, const int32_t *modesB, const int64_t *extentsB,
                                  const int64_t *stridesB, const mb200_comm_t *comm, int phases) {
    MB200_CHECK_HANDLE(h);
    if (phases < 1 || phases > 7) return fail(MB200_INVALID_ARGUMENT, "phases must be a non-empty subset of MB200_DIST_*");
    DistDesc d;
    int st;
    if ((st = make_dist(d, comm)) != MB200_OK) return st;
    TensorDesc dA, dB, dC;
    if ((st = make_desc(dA, dtypeA, nmodeA, modesA, extentsA, stridesA, "A")) != MB200_OK) return st;
    if ((st = make_desc(dB, dtypeB, nmodeB, modesB, extentsB, stridesB, "B")) != MB200_OK) return st;
    if ((st = make_c_desc(dC, dtypeC, nmodeC, modesC)) != MB200_OK) return st;
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->capturing) return fail(MB200_NOT_SUPPORTED, "the fused all-reduce cannot be captured into a graph");
    static const bool want_tl = [] { const char *e = getenv("MB200_DIST_TIMELINE"); return e && atoi(e) != 0; }();
    if (want_tl) {
        MB200_CUDA(cudaSetDevice(h->device));
        if (!h->timeline) MB200_CUDA(cudaMalloc((void **)&h->timeline, 8 * sizeof(unsigned long long)));
        if (phases & DIST_CONTRACT) {   // a new call: min-stamps start at +inf, max-stamps at 0
            const unsigned long long init[8] = {~0ull, 0, ~0ull, ~0ull, 0, 0, 0, 0};
            MB200_CUDA(cudaMemcpyAsync(h->timeline, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
        }
        d.timeline = h->timeline;
    }
    return contract_device(h, nullptr, dC, nullptr, A, dA, B, dB, nullptr, &d, phases, comm->ws_bytes, comm->flag_bytes);
}

int mb200_dist_timeline(mb200_handle_t h, unsigned long long *out8) {
    MB200_CHECK_HANDLE(h);
    if (!out8) return fail(MB200_INVALID_ARGUMENT, "out8 is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->timeline) return fail(MB200_NOT_SUPPORTED, "no timeline recorded (MB200_DIST_TIMELINE=1 and one fused all-reduce first)");
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaDeviceSynchronize());
    MB200_CUDA(cudaMemcpy(out8, h->timeline, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return MB200_OK;
}

int mb200_reduce_slots(mb200_handle_t h, void *out, const void *staging_local, int dtype, int64_t slab_elems, int nslots) {
    MB200_CHECK_HANDLE(h);
    if (!dtype_valid(dtype) || !out || !staging_local || slab_elems < 0 || nslots < 1)
        return fail(MB200_INVALID_ARGUMENT, "bad arguments to mb200_reduce_slots");
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(launch_reduce_slots(dtype, out, staging_local, slab_elems, nslots, h->stream));
    h->stats.launches_reduce++;
    h->stats.launches_total++;
    return MB200_OK;
}

int mb200_graph_begin(mb200_handle_t h) {
    MB200_CHECK_HANDLE(h);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->capturing) return fail(MB200_INVALID_ARGUMENT, "a capture is already in progress on this handle");
    if (h->stream == nullptr || h->stream == cudaStreamLegacy)
        return fail(MB200_NOT_SUPPORTED, "graph capture needs a non-default stream (mb200_set_stream)");
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    h->capturing = true;
    h->captured.clear();
    return MB200_OK;
}

int mb200_graph_end(mb200_handle_t h, mb200_graph_t *graph) {
    MB200_CHECK_HANDLE(h);
    if (!graph) return fail(MB200_INVALID_ARGUMENT, "graph is NULL");
    *graph = nullptr;
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->capturing) return fail(MB200_INVALID_ARGUMENT, "no capture in progress");
    h->capturing = false;
    MB200_CUDA(cudaSetDevice(h->device));
    auto g = std::make_unique<mb200_graph_s>();
    g->device = h->device;
    g->plans.swap(h->captured);
    MB200_CUDA(cudaStreamEndCapture(h->stream, &g->graph));
    cudaError_t e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g->graph);
        return cuda_fail(e, "cudaGraphInstantiate");
    }
    *graph = g.release();
    return MB200_OK;
}

int mb200_graph_launch(mb200_handle_t h, mb200_graph_t g) {
    MB200_CHECK_HANDLE(h);
    if (!g || !g->exec) return fail(MB200_INVALID_ARGUMENT, "graph is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(cudaGraphLaunch(g->exec, h->stream));
    h->stats.graph_launches++;
    return MB200_OK;
}

int mb200_graph_destroy(mb200_graph_t g) {
    if (!g) return MB200_OK;
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;   // releases its plans; tables of plans already evicted from their handle's cache are freed here
    return MB200_OK;
}

int mb200_signal_peers(mb200_handle_t h, void *const *flag_arrays, int nranks, int rank, int epoch) {
    MB200_CHECK_HANDLE(h);
    if (!flag_arrays || nranks < 1 || nranks > MB200_MAX_PEERS || rank < 0 || rank >= nranks)
        return fail(MB200_INVALID_ARGUMENT, "bad arguments to mb200_signal_peers");
    ScatterDesc f{};
    for (int r = 0; r < nranks; r++) {
        if (!flag_arrays[r]) return fail(MB200_INVALID_ARGUMENT, "flag_arrays[%d] is NULL", r);
        f.peer[r] = flag_arrays[r];
    }
    f.nranks = nranks; f.rank = rank;
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(launch_signal_peers(f, epoch, h->stream));
    h->stats.launches_reduce++;
    h->stats.launches_total++;
    return MB200_OK;
}

int mb200_reduce_slots_wait(mb200_handle_t h, void *out, const void *staging_local, int dtype, int64_t slab_elems, int nslots,
                            const void *flags_local, int epoch) {
    MB200_CHECK_HANDLE(h);
    if (!dtype_valid(dtype) || !out || !staging_local || !flags_local || slab_elems < 0 || nslots < 1 || nslots > MB200_MAX_PEERS)
        return fail(MB200_INVALID_ARGUMENT, "bad arguments to mb200_reduce_slots_wait");
    std::lock_guard<std::mutex> lk(h->mu);
    MB200_CUDA(cudaSetDevice(h->device));
    MB200_CUDA(launch_reduce_slots_wait(dtype, out, staging_local, slab_elems, nslots, (const int *)flags_local, epoch, h->stream));
    h->stats.launches_reduce++;
    h->stats.launches_total++;
    return MB200_OK;
}

int mb200_get_stats(mb200_handle_t h, mb200_stats_t *stats) {
    MB200_CHECK_HANDLE(h);
    if (!stats) return fail(MB200_INVALID_ARGUMENT, "stats is NULL");
    std::lock_guard<std::mutex> lk(h->mu);
    *stats = h->stats;
    return MB200_OK;
}

int mb200_reset_stats(mb200_handle_t h) {
    MB200_CHECK_HANDLE(h);
    std::lock_guard<std::mutex> lk(h->mu);
    h->stats = mb200_stats_t{};
    return MB200_OK;
}

}  // extern "C"
