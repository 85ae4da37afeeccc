"""muscle.jl_b200 — B200-native `binary_einsum` backend for Muscle.jl's hot path.

Only what the path needs lives here: the CUDA kernels + C ABI (`csrc/`, built into
`libmuscle_b200.so`) and a host-side mirror of the reference interface for that path
(`Tensor`, `Index`, `Backend`/`Domain` dispatch, `binary_einsum`, `binary_einsum_`).
Import it as `muscle_b200` (the repo-root shim maps the dotted directory name to a module).
"""
from ._lib import (ArgumentError, B200Error, DimensionMismatch, Handle, LIB_PATH, PATH_AUTO, PATH_DIRECT,
                   PATH_GETT_F64, PATH_NAMES, PATH_SIMT_F32, PATH_TCGEN05_TF32, lib, plan_describe, shard_plan)
from .backend import (Backend, BackendB200, BackendBase, BackendBlocks, BackendOMEinsum, Domain, DomainB200, DomainBlocks, DomainHost, choose_backend,
                      choose_backend_rule, domain, with_backend)
from .einsum import binary_einsum, binary_einsum_, binary_einsum_inplace, flatten_labels, frontend_inds_c
from .factorize import (AbsorbEqually, AbsorbU, AbsorbV, DontAbsorb, factorinds, simple_update, svd_last_info,
                        tensor_qr_thin, tensor_svd_thin, tensor_svd_trunc)
from .family import hadamard, hadamard_, unary_einsum, unary_einsum_, unary_frontend_inds_y
from .network import CapturedProgram, ContractionProgram, contract, find_path
from .tensor import B200Array, Index, Tensor, findperm
from .blocks import BlockArray, Blocks, blocked_binary_einsum, distribute
from . import dist

__all__ = [
    "ArgumentError", "B200Error", "DimensionMismatch", "Handle", "LIB_PATH", "lib", "plan_describe", "shard_plan",
    "PATH_AUTO", "PATH_DIRECT", "PATH_GETT_F64", "PATH_SIMT_F32", "PATH_TCGEN05_TF32", "PATH_NAMES",
    "Backend", "BackendB200", "BackendBase", "BackendOMEinsum",
    "AbsorbEqually", "AbsorbU", "AbsorbV", "DontAbsorb", "factorinds", "simple_update", "tensor_qr_thin", "tensor_svd_thin", "tensor_svd_trunc",
    "CapturedProgram", "ContractionProgram", "contract", "find_path",
    "hadamard", "hadamard_", "unary_einsum", "unary_einsum_", "unary_frontend_inds_y", "Domain", "DomainB200", "DomainHost", "choose_backend",
    "choose_backend_rule", "domain", "with_backend",
    "binary_einsum", "binary_einsum_", "binary_einsum_inplace", "flatten_labels", "frontend_inds_c",
    "B200Array", "Index", "Tensor", "findperm",
    "BackendBlocks", "DomainBlocks", "BlockArray", "Blocks", "blocked_binary_einsum", "distribute",
]
