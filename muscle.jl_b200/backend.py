"""Mirror of Muscle's type-dispatched Backend / Domain plumbing (src/Backend.jl:4-36,
src/Domain.jl:4-14), with the one new backend this project adds: `BackendB200`.
"""
from __future__ import annotations

import contextvars

import numpy as np

from ._lib import ArgumentError
from .tensor import B200Array, Tensor


class Backend:
    """`abstract type Backend` (src/Backend.jl:4)."""

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return f"{type(self).__name__}()"


class BackendBase(Backend):
    """Muscle's own CPU TTGT backend (src/Operations/binary_einsum.jl:76-121). It lives in Muscle.jl,
    not in this package: selecting it here raises the reference's "not implemented or not loaded"
    ArgumentError (binary_einsum.jl:53-55)."""


class BackendOMEinsum(Backend):
    """The reference's default backend for `unary_einsum` on host arrays (src/Operations/unary_einsum.jl:20-24,
    ext/MuscleOMEinsumExt.jl). Like BackendBase it lives in Muscle.jl, not here."""


class BackendBlocks(Backend):
    """Stand-in for `BackendDagger` (src/Backend.jl:6-14, ext/MuscleDaggerExt): blocked operands, one `binary_einsum` per
    chunk pair (each re-enters the dispatch, so device chunks run on BackendB200) and an add-reduce over summed blocks."""


class BackendB200(Backend):
    """The new backend: ccall/ctypes → libmuscle_b200.so (CUDA for sm_100a)."""


class Domain:
    """`abstract type Domain` — memory-space trait (src/Domain.jl:4)."""

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __repr__(self):
        return f"{type(self).__name__}()"


class DomainHost(Domain):
    pass


class DomainB200(Domain):
    pass


class DomainBlocks(Domain):
    """Stand-in for `DomainDagger` (src/Domain.jl:6-9): a `blocks.BlockArray`."""


def domain(x) -> Domain:
    """`Domain(array)` (src/Domain.jl:11-14): numpy → host, B200Array → B200; Tensors are unwrapped."""
    if isinstance(x, Tensor):
        x = x.parent
    if isinstance(x, B200Array):
        return DomainB200()
    if getattr(x, "_is_block_array", False):
        return DomainBlocks()
    if isinstance(x, np.ndarray):
        return DomainHost()
    raise ArgumentError(f"no Domain for {type(x).__name__}")


# `const CURRENT_BACKEND = ScopedValue{Backend}()` (src/Backend.jl:16): task-local override
_CURRENT_BACKEND: contextvars.ContextVar = contextvars.ContextVar("CURRENT_BACKEND")


def with_backend(f, backend: Backend):
    """`with_backend(f, backend)` (src/Backend.jl:18)."""
    if not isinstance(backend, Backend):
        raise ArgumentError("backend must be a Backend")
    token = _CURRENT_BACKEND.set(backend)
    try:
        return f()
    finally:
        _CURRENT_BACKEND.reset(token)


# rule tables: (function name, domains...) -> backend   (binary_einsum.jl:20-31 pattern)
_RULES: dict = {}


def choose_backend_rule(fname: str, *domains):
    try:
        return _RULES[(fname,) + tuple(type(d) for d in domains)]
    except KeyError:
        raise ArgumentError(f"no backend rule for {fname} on {domains}") from None


def register_rule(fname: str, domains, backend: Backend):
    _RULES[(fname,) + tuple(domains)] = backend


def choose_backend(fname: str, *arrays) -> Backend:
    """`choose_backend(f, arrays...)` (src/Backend.jl:28-36)."""
    try:
        return _CURRENT_BACKEND.get()
    except LookupError:
        pass
    return choose_backend_rule(fname, *[domain(a) for a in arrays])


# binary_einsum rules. Host×Host keeps the reference's answer (BackendBase, binary_einsum.jl:20);
# device and mixed host/device operands select the new backend (pattern of binary_einsum.jl:21-24).
register_rule("binary_einsum", (DomainHost, DomainHost), BackendBase())
register_rule("binary_einsum", (DomainB200, DomainB200), BackendB200())
register_rule("binary_einsum", (DomainB200, DomainHost), BackendB200())
register_rule("binary_einsum", (DomainHost, DomainB200), BackendB200())
# blocked operands, also mixed with a plain array (binary_einsum.jl:25-31: Dagger rules incl. mixed-with-Host)
for _d in (DomainBlocks, DomainHost, DomainB200):
    register_rule("binary_einsum", (DomainBlocks, _d), BackendBlocks())
    register_rule("binary_einsum", (_d, DomainBlocks), BackendBlocks())
register_rule("binary_einsum!", (DomainHost, DomainHost, DomainHost), BackendBase())
register_rule("binary_einsum!", (DomainB200, DomainB200, DomainB200), BackendB200())

# unary_einsum / hadamard: host arrays keep the reference's answers (unary_einsum.jl:20-24,
# hadamard.jl:4-5); device (and mixed) operands select the new backend.
register_rule("unary_einsum", (DomainHost,), BackendOMEinsum())
register_rule("unary_einsum", (DomainB200,), BackendB200())
register_rule("unary_einsum!", (DomainHost, DomainHost), BackendOMEinsum())
register_rule("unary_einsum!", (DomainB200, DomainB200), BackendB200())
register_rule("hadamard", (DomainHost, DomainHost), BackendBase())
register_rule("hadamard", (DomainB200, DomainB200), BackendB200())
register_rule("hadamard", (DomainB200, DomainHost), BackendB200())
register_rule("hadamard", (DomainHost, DomainB200), BackendB200())
register_rule("hadamard!", (DomainHost, DomainHost, DomainHost), BackendBase())
register_rule("hadamard!", (DomainB200, DomainB200, DomainB200), BackendB200())
register_rule("hadamard!", (DomainB200, DomainB200, DomainHost), BackendB200())

# tensor_svd_thin / simple_update: tensor_svd.jl:38-39, simple_update.jl:4-5
register_rule("tensor_svd_thin", (DomainHost,), BackendBase())
register_rule("tensor_svd_thin", (DomainB200,), BackendB200())
register_rule("simple_update", (DomainHost, DomainHost, DomainHost), BackendBase())
register_rule("simple_update", (DomainB200, DomainB200, DomainB200), BackendB200())
register_rule("tensor_qr_thin", (DomainHost,), BackendBase())
register_rule("tensor_qr_thin", (DomainB200,), BackendB200())
