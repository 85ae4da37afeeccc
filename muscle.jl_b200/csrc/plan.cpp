// Contraction planner (host only). See plan.hpp.
#include "plan.hpp"

#include <algorithm>
#include <cstring>

namespace mb200 {

std::string &last_error() {
    thread_local std::string e;
    return e;
}

int fail(int status, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return status;
}

int make_desc(TensorDesc &d, int dtype, int nmode, const int32_t *modes, const int64_t *extents,
              const int64_t *strides, const char *name) {
    if (!dtype_valid(dtype)) return fail(MB200_INVALID_ARGUMENT, "%s: unknown dtype %d", name, dtype);
    if (nmode < 0 || nmode > MB200_MAX_MODES)
        return fail(MB200_INVALID_ARGUMENT, "%s: nmode %d outside [0, %d]", name, nmode, MB200_MAX_MODES);
    if (nmode > 0 && (!modes || !extents))
        return fail(MB200_INVALID_ARGUMENT, "%s: modes/extents must not be NULL when nmode > 0", name);
    d.dtype = dtype;
    d.n = nmode;
    int64_t s = 1;
    for (int i = 0; i < nmode; i++) {
        d.modes[i] = modes[i];
        d.ext[i] = extents[i];
        if (extents[i] < 0) return fail(MB200_INVALID_ARGUMENT, "%s: negative extent", name);
        for (int j = 0; j < i; j++)
            if (modes[j] == modes[i])
                return fail(MB200_INVALID_ARGUMENT,
                            "%s: mode %d is repeated inside one tensor (traces/diagonals are unary_einsum, "
                            "not binary_einsum)", name, (int)modes[i]);
        d.stride[i] = strides ? strides[i] : s;
        s *= extents[i];
    }
    return MB200_OK;
}

bool k8_groupable(const std::vector<GroupMode> &sum) {
    int64_t ks = 1;
    for (const GroupMode &g : sum) {
        const int64_t need = 8 / ks;
        if (g.extent % need == 0) return true;      // this mode completes the first group of 8
        if (need % g.extent != 0) return false;     // straddles a group boundary
        ks *= g.extent;
    }
    return false;                                   // fewer than 8 summed elements
}

static int find_mode(const TensorDesc &t, int32_t m) {
    for (int i = 0; i < t.n; i++)
        if (t.modes[i] == m) return i;
    return -1;
}

std::vector<int> dangling_modes(const TensorDesc &T, const TensorDesc &other, const TensorDesc &C) {
    std::vector<int> out;
    for (int i = 0; i < T.n; i++)
        if (find_mode(other, T.modes[i]) < 0 && find_mode(C, T.modes[i]) < 0) out.push_back(i);
    return out;
}

TensorDesc without_modes(const TensorDesc &T, const std::vector<int> &drop, bool redense) {
    TensorDesc r;
    r.dtype = T.dtype;
    r.n = 0;
    int64_t s = 1;
    for (int i = 0; i < T.n; i++) {
        if (std::find(drop.begin(), drop.end(), i) != drop.end()) continue;
        r.modes[r.n] = T.modes[i];
        r.ext[r.n] = T.ext[i];
        r.stride[r.n] = redense ? s : T.stride[i];
        s *= T.ext[i];
        r.n++;
    }
    return r;
}

int make_plan(const TensorDesc &A, const TensorDesc &B, TensorDesc &C, const int64_t *stridesC,
              int forced_path, Plan &plan) {
    // ---- validation -------------------------------------------------------------------------
    if (C.dtype != dtype_promote(A.dtype, B.dtype))
        return fail(MB200_INVALID_ARGUMENT, "eltype of C (%s) must be promote_eltype(A, B) = %s",
                    dtype_name(C.dtype), dtype_name(dtype_promote(A.dtype, B.dtype)));
    for (int i = 0; i < C.n; i++)
        for (int j = 0; j < i; j++)
            if (C.modes[i] == C.modes[j])
                return fail(MB200_INVALID_ARGUMENT, "C: mode %d is repeated", (int)C.modes[i]);
    for (int i = 0; i < A.n; i++) {
        int j = find_mode(B, A.modes[i]);
        if (j >= 0 && A.ext[i] != B.ext[j])
            return fail(MB200_DIMENSION_MISMATCH, "mode %d has extent %lld in A but %lld in B",
                        (int)A.modes[i], (long long)A.ext[i], (long long)B.ext[j]);
        if (j < 0 && find_mode(C, A.modes[i]) < 0)
            return fail(MB200_INVALID_ARGUMENT,
                        "internal: mode %d appears only in A and not in the output (dangling modes are summed by "
                        "the entry points before planning)", (int)A.modes[i]);
    }
    for (int i = 0; i < B.n; i++)
        if (find_mode(A, B.modes[i]) < 0 && find_mode(C, B.modes[i]) < 0)
            return fail(MB200_INVALID_ARGUMENT,
                        "internal: mode %d appears only in B and not in the output (dangling modes are summed by "
                        "the entry points before planning)", (int)B.modes[i]);
    // C extents are implied by A / B
    {
        int64_t s = 1;
        for (int i = 0; i < C.n; i++) {
            int ia = find_mode(A, C.modes[i]), ib = find_mode(B, C.modes[i]);
            if (ia < 0 && ib < 0)
                return fail(MB200_INVALID_ARGUMENT, "mode %d of the output is found in neither operand",
                            (int)C.modes[i]);
            C.ext[i] = ia >= 0 ? A.ext[ia] : B.ext[ib];
            C.stride[i] = stridesC ? stridesC[i] : s;
            s *= C.ext[i];
        }
    }

    plan = Plan();
    plan.dtype = C.dtype;

    // ---- swap so that C's unit-stride mode is a row mode --------------------------------------
    // The kernels store C with consecutive rows (m) in consecutive lanes.
    int fastestC = -1;
    {
        int64_t best = INT64_MAX;
        for (int i = 0; i < C.n; i++)
            if (C.ext[i] > 1 && C.stride[i] < best) { best = C.stride[i]; fastestC = i; }
    }
    bool swap = false;
    if (fastestC >= 0) {
        int32_t m = C.modes[fastestC];
        if (find_mode(A, m) < 0 && find_mode(B, m) >= 0) swap = true;
    }
    const TensorDesc &R = swap ? B : A;   // row operand
    const TensorDesc &Q = swap ? A : B;   // column operand
    plan.swapped = swap;
    plan.dtype_row = R.dtype;
    plan.dtype_col = Q.dtype;

    // ---- classification -----------------------------------------------------------------------
    for (int i = 0; i < R.n; i++) {
        if (R.ext[i] == 1) continue;  // extent-1 modes never contribute to an address
        int32_t m = R.modes[i];
        int iq = find_mode(Q, m), ic = find_mode(C, m);
        GroupMode g{m, R.ext[i], R.stride[i], iq >= 0 ? Q.stride[iq] : 0, ic >= 0 ? C.stride[ic] : 0};
        if (iq >= 0 && ic >= 0) plan.batch.push_back(g);
        else if (iq >= 0) plan.sum.push_back(g);
        else plan.left.push_back(g);
    }
    for (int i = 0; i < Q.n; i++) {
        if (Q.ext[i] == 1) continue;
        int32_t m = Q.modes[i];
        if (find_mode(R, m) >= 0) continue;
        int ic = find_mode(C, m);
        plan.right.push_back(GroupMode{m, Q.ext[i], 0, Q.stride[i], C.stride[ic]});
    }
    plan.empty_output = false;
    for (int i = 0; i < C.n; i++)
        if (C.ext[i] == 0) plan.empty_output = true;

    // ---- walk order inside each group ----------------------------------------------------------
    auto by_sc = [](const GroupMode &x, const GroupMode &y) { return x.sc < y.sc; };
    auto by_sa = [](const GroupMode &x, const GroupMode &y) { return x.sa < y.sa; };
    auto by_sb = [](const GroupMode &x, const GroupMode &y) { return x.sb < y.sb; };
    std::stable_sort(plan.left.begin(), plan.left.end(), by_sc);
    std::stable_sort(plan.right.begin(), plan.right.end(), by_sc);
    std::stable_sort(plan.batch.begin(), plan.batch.end(), by_sc);
    // unit-stride mode of each operand decides whether lanes should run along k when loading it
    auto fastest_is_sum = [&](const TensorDesc &T) {
        int64_t best = INT64_MAX;
        int32_t m = -1;
        for (int i = 0; i < T.n; i++)
            if (T.ext[i] > 1 && T.stride[i] < best) { best = T.stride[i]; m = T.modes[i]; }
        if (m < 0) return false;
        for (auto &g : plan.sum)
            if (g.label == m) return true;
        return false;
    };
    plan.a_kmajor = fastest_is_sum(R);
    plan.b_kmajor = fastest_is_sum(Q);
    // summed modes follow the row operand's memory order (the reference uses A's label order,
    // binary_einsum.jl:77; any common order is valid) unless only the column operand is K-major.
    if (plan.b_kmajor && !plan.a_kmajor) std::stable_sort(plan.sum.begin(), plan.sum.end(), by_sb);
    else std::stable_sort(plan.sum.begin(), plan.sum.end(), by_sa);

    // ComplexF32: the tcgen05 operand format stores k in groups of 8 (one 8-k group = one 128 B TMA row), so the
    // leading summed modes must tile a group exactly: extents 8 | e, or 2*2*2, 4*2, 2*4*..., 2*16 (split) ...
    // (tensor-network tensors are mostly dim-2 / dim-4 indices). Any common order of the summed modes is valid, so
    // try the memory order first, then a multiple-of-8 mode in front, then power-of-two extents first.
    if ((plan.dtype == MB200_C64 || plan.dtype == MB200_F32) && !plan.sum.empty() && !k8_groupable(plan.sum)) {
        std::vector<GroupMode> cand = plan.sum;
        bool found = false;
        for (size_t i = 1; i < cand.size() && !found; i++)
            if (cand[i].extent % 8 == 0) {
                std::rotate(cand.begin(), cand.begin() + i, cand.begin() + i + 1);
                found = true;
            }
        if (!found) {
            cand = plan.sum;
            std::stable_partition(cand.begin(), cand.end(),
                                  [](const GroupMode &g) { return (g.extent & (g.extent - 1)) == 0; });
            found = k8_groupable(cand);
        }
        if (found) plan.sum = cand;
    }

    // merged walk groups: neighbours contiguous in every tensor that carries the group collapse into one mode
    auto merged = [](const std::vector<GroupMode> &g, bool in_a, bool in_b, bool in_c) {
        std::vector<GroupMode> out;
        for (const GroupMode &m : g) {
            if (!out.empty()) {
                GroupMode &t = out.back();
                const bool ok = (!in_a || m.sa == t.sa * t.extent) && (!in_b || m.sb == t.sb * t.extent) &&
                                (!in_c || m.sc == t.sc * t.extent);
                if (ok) { t.extent *= m.extent; continue; }
            }
            out.push_back(m);
        }
        return out;
    };
    plan.mleft = merged(plan.left, true, false, true);
    plan.mright = merged(plan.right, false, true, true);
    plan.msum = merged(plan.sum, true, true, false);
    plan.mbatch = merged(plan.batch, true, true, true);

    auto prod = [](const std::vector<GroupMode> &g) {
        int64_t p = 1;
        for (auto &x : g) p *= x.extent;
        return p;
    };
    plan.M = prod(plan.left);
    plan.N = prod(plan.right);
    plan.K = prod(plan.sum);
    plan.L = prod(plan.batch);
    double macs = (double)plan.M * (double)plan.N * (double)plan.K * (double)plan.L;
    plan.flops = (dtype_is_complex(plan.dtype) ? 8.0 : 2.0) * macs;
    plan.bytes = (double)dtype_size(plan.dtype) *
                 ((double)A.numel() + (double)B.numel() + (double)plan.M * plan.N * plan.L);

    // ---- tcgen05 eligibility: ComplexF32 / Float32 with enough rows, columns and k to fill 128-row tiles and 128-k chunks.
    // Operands that are dense in memory with leading summed modes tiling groups of 8 k are packed by a K1 permutation
    // (tc_permute_pack); everything else (K = 100, bond dimensions 3, 5, 6 ..., strided operands) by the table-driven gather
    // pack with K zero-padded to a multiple of 8.
    auto dense = [](const TensorDesc &T) {
        std::vector<std::pair<int64_t, int64_t>> v;
        for (int i = 0; i < T.n; i++)
            if (T.ext[i] != 1) v.push_back({T.stride[i], T.ext[i]});
        std::sort(v.begin(), v.end());
        int64_t expect = 1;
        for (auto &x : v) {
            if (x.first != expect) return false;
            expect *= x.second;
        }
        return true;
    };
    plan.tc_ok = (plan.dtype == MB200_C64 || plan.dtype == MB200_F32) && plan.M >= 64 && plan.N >= 32 && plan.K >= 64 &&
                 plan.M < ((int64_t)1 << 31) && plan.N < ((int64_t)1 << 31) && plan.L < ((int64_t)1 << 31) &&
                 plan.K < ((int64_t)1 << 27);
    plan.tc_permute_pack = plan.tc_ok && k8_groupable(plan.sum) && dense(A) && dense(B);

    // ---- kernel family ---------------------------------------------------------------------------
    const int64_t TABLE_LIMIT = (int64_t)1 << 26;
    bool tables_ok = plan.M <= TABLE_LIMIT && plan.N <= TABLE_LIMIT && plan.K <= TABLE_LIMIT &&
                     plan.L <= TABLE_LIMIT;
    int path;
    // few outputs with a long sum (inner products, norms): the direct kernel's split-K form parallelises over K
    const bool dot_like = plan.M * plan.N * plan.L <= 2048 && plan.M * plan.N <= 64 * 64 && plan.K >= 8192 &&
                          (plan.M < 16 || plan.N < 16);
    // a tiny operator applied to a big tensor (gate application): streaming "apply" form of the direct path
    // (for FP64 types an 8 x 8 operator is better served by the DMMA streaming kernel: 5.5 vs 4.3 TB/s)
    const int64_t apply_cap = dtype_is_double(plan.dtype) ? 4 : 8;
    plan.apply_like = plan.K <= apply_cap && std::min(plan.M, plan.N) <= apply_cap && plan.L == 1 && plan.K >= 1 &&
                      std::max(plan.M, plan.N) < ((int64_t)1 << 31);
    if (plan.empty_output || macs <= (double)(1 << 20) || plan.K <= 2 || !tables_ok || dot_like || plan.apply_like)
        path = MB200_PATH_DIRECT;
    else if (dtype_is_double(plan.dtype))
        path = MB200_PATH_GETT_F64;
    else
        path = (plan.tc_ok && macs >= (double)((int64_t)1 << 27)) ? MB200_PATH_TCGEN05_TF32 : MB200_PATH_SIMT_F32;
    if (forced_path != MB200_PATH_AUTO) {
        if (forced_path == MB200_PATH_DIRECT) path = MB200_PATH_DIRECT;
        else if (!plan.empty_output && tables_ok) {
            if (forced_path == MB200_PATH_GETT_F64 && dtype_is_double(plan.dtype)) path = forced_path;
            if (forced_path == MB200_PATH_SIMT_F32 && !dtype_is_double(plan.dtype)) path = forced_path;
            if (forced_path == MB200_PATH_TCGEN05_TF32 && plan.tc_ok) path = forced_path;
            if (forced_path == MB200_PATH_TCGEN05_TF32 && !plan.tc_ok && !dtype_is_double(plan.dtype)) path = MB200_PATH_SIMT_F32;
        }
    }
    plan.path = path;

    // ---- cache key ----------------------------------------------------------------------------------
    std::string &k = plan.key;
    k.clear();
    auto put = [&k](int64_t v) { k.append(reinterpret_cast<const char *>(&v), sizeof v); };
    put(forced_path);
    for (const TensorDesc *t : {&A, &B, (const TensorDesc *)&C}) {
        put(t->dtype);
        put(t->n);
        for (int i = 0; i < t->n; i++) { put(t->modes[i]); put(t->ext[i]); put(t->stride[i]); }
    }
    return MB200_OK;
}

void fill_info(const Plan &p, mb200_plan_info_t *info) {
    std::memset(info, 0, sizeof *info);
    info->M = p.M; info->N = p.N; info->K = p.K; info->L = p.L;
    info->swapped = p.swapped;
    info->path = p.path;
    info->compute_dtype = p.dtype;
    info->n_left = (int)p.left.size();
    info->n_right = (int)p.right.size();
    info->n_sum = (int)p.sum.size();
    info->n_batch = (int)p.batch.size();
    for (size_t i = 0; i < p.left.size(); i++) info->left[i] = p.left[i].label;
    for (size_t i = 0; i < p.right.size(); i++) info->right[i] = p.right[i].label;
    for (size_t i = 0; i < p.sum.size(); i++) info->sum[i] = p.sum[i].label;
    for (size_t i = 0; i < p.batch.size(); i++) info->batch[i] = p.batch[i].label;
    info->a_kmajor = p.a_kmajor;
    info->b_kmajor = p.b_kmajor;
    info->flops = p.flops;
    info->bytes = p.bytes;
    info->tc_eligible = p.tc_ok;
    info->tc_permute_pack = p.tc_permute_pack;
}

}  // namespace mb200
