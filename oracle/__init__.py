"""TEST INFRASTRUCTURE ONLY — CPU oracle for Muscle.jl's `binary_einsum` hot path (and its einsum-family
neighbours `unary_einsum` / `hadamard`).

Nothing in the product package (`muscle.jl_b200/`) may import this. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs use it,
and only as the checker / the CPU arm, never as the thing shipped.
"""
from .muscle_oracle import (  # noqa: F401
    ArgumentError,
    DimensionMismatch,
    frontend_inds_c,
    binary_einsum_base,
    binary_einsum_base_inplace,
    binary_einsum_general,
    binary_einsum,
    permutedims,
    rel_frobenius,
)
from .family_oracle import (  # noqa: F401
    contract_path_oracle,
    dagger_stage_oracle,
    simple_update_theta,
    tensor_svd_thin_base,
    hadamard_base,
    unary_einsum,
    unary_einsum_general,
    unary_frontend_inds_y,
)
