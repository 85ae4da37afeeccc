// Contraction planner: index classification and canonical GEMM view of a pairwise contraction.
// Pure host code (no CUDA) so the bookkeeping is testable on a CPU-only box through
// mb200_plan_describe.
//
// Classification follows the union of the reference's backends:
//   batch/hyper = in A, B and C     (ext/MuscleReactantExt.jl:101-106, ext/MuscleOMEinsumExt.jl:40-49)
//   summed      = in A and B, not C (src/Operations/binary_einsum.jl:77)
//   free-left   = in A only, in C   (binary_einsum.jl:78)
//   free-right  = in B only, in C   (binary_einsum.jl:79)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "common.hpp"

namespace mb200 {

struct TensorDesc {
    int dtype = 0;
    int n = 0;
    int32_t modes[MB200_MAX_MODES];
    int64_t ext[MB200_MAX_MODES];
    int64_t stride[MB200_MAX_MODES];  // elements
    int64_t numel() const {
        int64_t t = 1;
        for (int i = 0; i < n; i++) t *= ext[i];
        return t;
    }
};

// One mode of a group with its strides in the row operand ("a"), the column operand ("b") and C.
// A stride of 0 means "absent from that tensor".
struct GroupMode {
    int32_t label;
    int64_t extent;
    int64_t sa, sb, sc;
};

struct Plan {
    int dtype = 0;          // compute (= C) dtype
    int dtype_row = 0;      // dtype of the row operand as given (before promotion)
    int dtype_col = 0;
    bool swapped = false;   // row operand is the caller's B
    std::vector<GroupMode> left, right, sum, batch;  // extent-1 modes removed, walk order (fastest first)
    // the same groups with neighbours merged when they are contiguous in every tensor that carries them (a run of
    // dim-2 indices q0..q12 untouched by a gate becomes one mode of extent 8192): what the kernels actually walk
    std::vector<GroupMode> mleft, mright, msum, mbatch;
    int64_t M = 1, N = 1, K = 1, L = 1;
    bool a_kmajor = false, b_kmajor = false;
    int path = MB200_PATH_DIRECT;
    bool empty_output = false;  // some C extent is 0
    bool apply_like = false;    // K <= 8 and one free side <= 8, no batch: the direct path uses its streaming apply kernel
    bool tc_ok = false;         // eligible for the tcgen05 path (ComplexF32 / Float32, M >= 64, N >= 32, K >= 64)
    bool tc_permute_pack = false;   // ... with operands a K1 permutation can pack (dense, leading summed modes tile groups of 8 k);
                                    // otherwise the table-driven gather pack is used (any strides, K zero-padded to 8)
    double flops = 0, bytes = 0;
    std::string key;  // cache key (all integers of the three descriptors + dtypes + forced path)
};

// Fills `desc` from raw ABI arguments; strides==NULL -> dense column-major. `extents` may be NULL
// only when nmode == 0.
int make_desc(TensorDesc &desc, int dtype, int nmode, const int32_t *modes, const int64_t *extents,
              const int64_t *strides, const char *name);

// Derives C's extents from A/B, validates everything the ABI promises to reject, classifies and
// orders the modes, and picks a kernel family. `C.ext` is written.
int make_plan(const TensorDesc &A, const TensorDesc &B, TensorDesc &C, const int64_t *stridesC,
              int forced_path, Plan &plan);

void fill_info(const Plan &plan, mb200_plan_info_t *info);

// "Dangling" modes: carried by ONE operand and absent from C. cuTENSOR.contract! (ext/MuscleCUDAExt.jl:30-38) and
// OMEinsum (ext/MuscleOMEinsumExt.jl:40-59) sum them; BackendBase rejects them (binary_einsum.jl:83). The library sums
// them before (or folded into) the contraction. Returns the positions inside T of its dangling modes.
std::vector<int> dangling_modes(const TensorDesc &T, const TensorDesc &other, const TensorDesc &C);
// T without the modes at `drop`; dense column-major strides in the remaining order when `redense`, else T's own strides
TensorDesc without_modes(const TensorDesc &T, const std::vector<int> &drop, bool redense);

// true when the leading summed modes tile a group of 8 k exactly (tcgen05 operand format, tf32.cu)
bool k8_groupable(const std::vector<GroupMode> &sum);

}  // namespace mb200
