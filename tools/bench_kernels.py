"""Per-kernel numbers for DESIGN.md / profiles/: K1 permute GB/s and binary_einsum TFLOP/s on the BASELINE
configs (aligned and scrambled layouts). CUDA-event timed, device-resident, best-of and mean. Writes
gpurun_out/kernels.json."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import muscle_b200 as mb  # noqa: E402
from muscle_b200 import B200Array, Index, Tensor, _lib, binary_einsum  # noqa: E402

I = lambda s: [Index(c) for c in s]


def dev_rand(shape, dtype, seed=0):
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    n = int(np.prod(shape))
    cplx = np.dtype(dtype).kind == "c"
    real = torch.float64 if np.dtype(dtype).itemsize // (2 if cplx else 1) == 8 else torch.float32
    t = torch.rand((2 if cplx else 1) * n, dtype=real, device="cuda:0", generator=g) * 2 - 1
    return B200Array.from_torch(t, shape, dtype)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.min(ts)), float(np.mean(ts))


def permute_case(name, shape, perm, dtype, flags=0):
    src = dev_rand(shape, dtype)
    dst = B200Array([shape[p] for p in perm], dtype) if not flags else B200Array(list(shape) + [2], np.float64 if dtype == "complex128" else np.float32)
    h = _lib.Handle.get()
    fn = lambda: _lib.check(mb.lib().mb200_permute(h.ptr, C.c_void_p(dst.ptr), C.c_void_p(src.ptr), _lib.dtype_enum(dtype),
                                                    len(shape), _lib.i64(shape), _lib.i32(perm), flags))
    best, mean = timeit(fn)
    nbytes = 2 * src.nbytes
    return {"name": name, "shape": list(shape), "perm": list(perm), "dtype": dtype, "planar": bool(flags),
            "bytes": nbytes, "ms_best": best, "ms_mean": mean, "gbs_best": nbytes / best / 1e6, "gbs_mean": nbytes / mean / 1e6}


def family_cases():
    """unary_einsum / hadamard streaming kernels: GB/s over the algorithmic bytes (|x|+|y|, |a|+|b|+|c|)."""
    out = []
    h = _lib.Handle.get()
    L = mb.lib()

    def unary(name, shape, ix, iy, dtype):
        x = dev_rand(shape, dtype)
        m = {c: k for k, c in enumerate(dict.fromkeys(ix))}
        ext = {c: shape[ix.index(c)] for c in ix}
        y = B200Array([ext[c] for c in iy], dtype)
        fn = lambda: _lib.check(L.mb200_unary_einsum(h.ptr, C.c_void_p(y.ptr), _lib.dtype_enum(dtype), len(iy), _lib.i32([m[c] for c in iy]), None,
                                                     C.c_void_p(x.ptr), _lib.dtype_enum(dtype), len(ix), _lib.i32([m[c] for c in ix]), _lib.i64(shape), None))
        best, mean = timeit(fn)
        # algorithmic bytes: every DISTINCT-label element of x is read once (a repeated label reads only its diagonal)
        nb = int(np.prod(list(ext.values()))) * np.dtype(dtype).itemsize + y.nbytes
        out.append({"name": name, "op": "unary_einsum", "dtype": dtype, "bytes": nb, "ms_best": best, "gbs_best": nb / best / 1e6, "gbs_mean": nb / mean / 1e6})

    def had(name, sa, ia, sb, ib, dtype):
        a, b = dev_rand(sa, dtype, 1), dev_rand(sb, dtype, 2)
        c = B200Array(sa, dtype)
        m = {ch: k for k, ch in enumerate(ia)}
        fn = lambda: _lib.check(L.mb200_hadamard(h.ptr, C.c_void_p(c.ptr), _lib.dtype_enum(dtype),
                                                 C.c_void_p(a.ptr), _lib.dtype_enum(dtype), len(ia), _lib.i32([m[ch] for ch in ia]), _lib.i64(sa),
                                                 C.c_void_p(b.ptr), _lib.dtype_enum(dtype), len(ib), _lib.i32([m[ch] for ch in ib]), _lib.i64(sb)))
        best, mean = timeit(fn)
        nb = 2 * a.nbytes + b.nbytes
        out.append({"name": name, "op": "hadamard", "dtype": dtype, "bytes": nb, "ms_best": best, "gbs_best": nb / best / 1e6, "gbs_mean": nb / mean / 1e6})

    for dt in ("complex128", "complex64"):
        n = 1024 if dt == "complex128" else 2048
        unary(f"column sums  x[a,b]->y[a]  {n}x16384 {dt}", (n, 16384), "ab", "a", dt)
        unary(f"row sums     x[a,b]->y[b]  16384x{n} {dt}", (16384, n), "ab", "b", dt)
        unary(f"sum of all   x[a,b]->y[]   {dt}", (n, 16384), "ab", "", dt)
        unary(f"partial trace x[a,b,a]->y[b] (64,{n * 4},64) {dt}", (64, n * 4, 64), "aba", "b", dt)
        unary(f"axis sum rank-4 x[a,b,c,d]->y[a,c] (256,64,{n // 4},64) {dt}", (256, 64, n // 4, 64), "abcd", "ac", dt)
        unary(f"axis sum rank-4 x[a,b,c,d]->y[d,b] (256,64,{n // 4},64) {dt}", (256, 64, n // 4, 64), "abcd", "db", dt)
        had(f"hadamard bond weights a[l,p,r] .* s[r] (1024,16,{n}) {dt}", (1024, 16, n), "lpr", (n,), "r", dt)
        had(f"hadamard a[l,p,r] .* s[l] {dt}", (1024, 16, n), "lpr", (1024,), "l", dt)
        had(f"hadamard full a .* b {dt}", (1024, 16, n), "lpr", (1024, 16, n), "lpr", dt)
    return out


def einsum_case(name, ext, ia, ib, ic, dtype, iters=5):
    A = Tensor(dev_rand([ext[c] for c in ia], dtype, 1), I(ia))
    B = Tensor(dev_rand([ext[c] for c in ib], dtype, 2), I(ib))
    labels = set(ia) | set(ib)
    flops = (8.0 if np.dtype(dtype).kind == "c" else 2.0) * float(np.prod([ext[c] for c in labels], dtype=np.float64))
    best, mean = timeit(lambda: binary_einsum(A, B, out=I(ic)), iters=iters, warm=2)
    return {"name": name, "dtype": dtype, "a": ia, "b": ib, "c": ic, "flops": flops, "ms_best": best, "ms_mean": mean,
            "tflops_best": flops / best / 1e9, "tflops_mean": flops / mean / 1e9}


def main():
    out = {"permute": [], "einsum": []}
    P = out["permute"]
    P.append(permute_case("cfg1_A_pack kilj->ijkl", (64, 64, 64, 64), (1, 3, 0, 2), "complex128"))
    P.append(permute_case("cfg1_B_pack nlmk->klnm", (64, 64, 64, 64), (3, 1, 0, 2), "complex128"))
    P.append(permute_case("transpose 4096x4096", (4096, 4096), (1, 0), "complex128"))
    P.append(permute_case("transpose 8192x8192 f32", (8192, 8192), (1, 0), "float32"))
    P.append(permute_case("transpose 8192x4096 c64", (8192, 4096), (1, 0), "complex64"))
    P.append(permute_case("transpose 8192x4096 f64", (8192, 4096), (1, 0), "float64"))
    P.append(permute_case("cfg2b_out awsc->atvc-like (1024,8,2,1024)->(0,2,1,3)", (1024, 8, 2, 1024), (0, 2, 1, 3), "complex128"))
    P.append(permute_case("mps d=2 fastest (2,1024,1024,8)->(1,0,3,2)", (2, 1024, 1024, 8), (1, 0, 3, 2), "complex128"))
    P.append(permute_case("rank6 dim16 reverse", (16,) * 6, (5, 4, 3, 2, 1, 0), "complex128"))
    P.append(permute_case("rank8 dim8 c64 interleave", (8,) * 8, (4, 0, 5, 1, 6, 2, 7, 3), "complex64"))
    P.append(permute_case("cfg3 pack+planar c64 (256,8,8,256,8)->(0,2,3,1,4)", (256, 8, 8, 256, 8), (0, 2, 3, 1, 4), "complex64", 1))
    P.append(permute_case("copy (identity) c128 64^4", (64, 64, 64, 64), (0, 1, 2, 3), "complex128"))
    P.append(permute_case("rank8 dim8 f32 interleave", (8,) * 9, (4, 0, 5, 1, 6, 2, 7, 3, 8), "float32"))
    P.append(permute_case("c64 (256,8,8,256,8)->(0,2,3,1,4) plain", (256, 8, 8, 256, 8), (0, 2, 3, 1, 4), "complex64"))
    P.append(permute_case("c64 (256,8,8,256,8)->(3,1,0,2,4) plain", (256, 8, 8, 256, 8), (3, 1, 0, 2, 4), "complex64"))
    P.append(permute_case("f64 rank6 dim16 reverse", (16,) * 6 + (2,), (5, 4, 3, 2, 1, 0, 6), "float64"))
    P.append(permute_case("qubit rank-26 c64 bit reversal", (2,) * 26, tuple(range(25, -1, -1)), "complex64") if False else
             permute_case("qubit rank-16 (dim 2 x13, 4096 tail) c64 reversal", (2,) * 13 + (4096,), tuple(range(13, -1, -1)), "complex64"))
    if "--svd" in sys.argv:
        import time
        from muscle_b200 import tensor_svd_thin
        rows = []
        for n, dt in ((128, "complex128"), (256, "complex128"), (512, "complex128"), (1024, "complex128"), (512, "complex64")):
            A = Tensor(dev_rand((n, n), dt), I("ab"))
            fn = lambda: tensor_svd_thin(A, inds_u=I("a"), ind_s=Index("s"))
            best, mean = timeit(fn, iters=3, warm=1)
            host = A.to_host().data
            t0 = time.perf_counter(); np.linalg.svd(host, full_matrices=False); cpu = (time.perf_counter() - t0) * 1e3
            rows.append({"n": n, "dtype": dt, "ms_best": best, "ms_mean": mean, "numpy_lapack_ms": cpu})
            print(f"SVD     {n}x{n} {dt:<11s} jacobi {best:9.2f} ms   numpy/LAPACK on host {cpu:9.2f} ms")
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump({"svd": rows}, open("gpurun_out/kernels_svd.json", "w"), indent=1)
        return
    if "--family" in sys.argv:
        out["family"] = family_cases()
        for r in out["family"]:
            print(f"FAMILY  {r['name']:<72s} {r['gbs_best']:8.0f} GB/s best {r['gbs_mean']:8.0f} mean  {r['ms_best']:.3f} ms")
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/kernels_family.json", "w"), indent=1)
        return
    if "--permute-only" in sys.argv:
        for r in P:
            print(f"PERMUTE {r['name']:<60s} {r['gbs_best']:8.0f} GB/s best {r['gbs_mean']:8.0f} mean")
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/kernels_permute%s.json" % os.environ.get("MB200_PERMUTE_REGT", ""), "w"), indent=1)
        return
    E = out["einsum"]
    e64 = dict(i=64, j=64, k=64, l=64, m=64, n=64)
    E.append(einsum_case("cfg1 aligned  A[i,j,k,l] B[k,l,m,n] -> [i,j,m,n]", e64, "ijkl", "klmn", "ijmn", "complex128"))
    E.append(einsum_case("cfg1 scrambled A[k,i,l,j] B[n,l,m,k] -> [m,j,n,i]", e64, "kilj", "nlmk", "mjni", "complex128"))
    e2 = dict(a=1024, b=1024, c=1024, e=1024, w=8, v=8, s=2, t=2)
    E.append(einsum_case("cfg2a", e2, "awb", "bsc", "awsc", "complex128"))
    E.append(einsum_case("cfg2b", e2, "awsc", "wstv", "atvc", "complex128", iters=20))
    E.append(einsum_case("cfg2c", e2, "atvc", "ate", "evc", "complex128"))
    e3 = dict(l=256, k=8, b=8, m=256, q=8, r=256, z=8)
    E.append(einsum_case("cfg3 PEPS batched c64", e3, "lkbmz", "mkqrz", "lbqrz", "complex64"))
    e4 = {c: 16 for c in "abcdefghi"}
    E.append(einsum_case("cfg4a rank6 dim16", e4, "adbecf", "fgdhei", "abcghi", "complex128"))
    e5 = {c: 8 for c in "abcdefghpqrs"}
    E.append(einsum_case("cfg5 rank8 dim8 c64 (unsliced)", e5, "aebfcgdh", "hpgqfres", "srqpdcba", "complex64"))
    # quantum-circuit style: rank-24 x rank-24 over dim-2 indices, 12 shared (4096^3 GEMM-equivalent), interleaved labels
    la = "aAbBcCdDeEfFgGhHiIjJkKlL"                      # lower-case free, upper-case shared, alternating in memory
    lb = "AmBnCoDpEqFrGsHtIuJvKwLx"
    lc = "mabncdopefqrghstijuvklwx"
    eq = {c: 2 for c in set(la + lb)}
    E.append(einsum_case("qubit rank-24 dim-2 c128 (interleaved)", eq, la, lb, lc, "complex128"))
    E.append(einsum_case("qubit rank-24 dim-2 c64 (interleaved)", eq, la, lb, lc, "complex64"))
    E.append(einsum_case("dgemm 8192^3 f64", dict(i=8192, j=8192, k=8192), "ik", "kj", "ij", "float64", iters=3))
    E.append(einsum_case("sgemm 8192^3 f32", dict(i=8192, j=8192, k=8192), "ik", "kj", "ij", "float32", iters=3))
    # launch-bound tiny contraction (everything in the reference's own test-suite is this size): per-call latency
    # of the C ABI alone (ctypes, cached plan, one direct-kernel launch) and of the Python front-end
    import time
    h = _lib.Handle.get()
    a = Tensor(dev_rand((2, 3), "complex128"), I("ij")); b = Tensor(dev_rand((3, 4), "complex128"), I("jk"))
    c = B200Array((2, 4), "complex128")
    args = (h.ptr, C.c_void_p(c.ptr), _lib.C128, 2, _lib.i32([0, 2]), None,
            C.c_void_p(a.data.ptr), _lib.C128, 2, _lib.i32([0, 1]), _lib.i64((2, 3)), None,
            C.c_void_p(b.data.ptr), _lib.C128, 2, _lib.i32([1, 2]), _lib.i64((3, 4)), None)
    fn = mb.lib().mb200_binary_einsum
    for _ in range(100):
        fn(*args)
    torch.cuda.synchronize()
    n = 5000
    t0 = time.perf_counter()
    for _ in range(n):
        fn(*args)
    torch.cuda.synchronize()
    abi_us = (time.perf_counter() - t0) / n * 1e6
    t0 = time.perf_counter()
    for _ in range(1000):
        binary_einsum(a, b)
    torch.cuda.synchronize()
    py_us = (time.perf_counter() - t0) / 1000 * 1e6
    out["tiny_contraction_latency_us"] = {"c_abi_call": abi_us, "python_front_end": py_us,
                                          "what": "2x3 . 3x4 ComplexF64 matmul, device-resident, one direct_kernel launch per call"}
    print(f"TINY    2x3.3x4 c128: {abi_us:.2f} us per C-ABI call, {py_us:.1f} us through the Python front-end")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/kernels.json", "w"), indent=1)
    for r in P:
        print(f"PERMUTE {r['name']:<60s} {r['gbs_best']:8.0f} GB/s best {r['gbs_mean']:8.0f} mean")
    for r in E:
        print(f"EINSUM  {r['name']:<60s} {r['tflops_best']:8.2f} TF/s best {r['tflops_mean']:8.2f} mean  {r['ms_mean']:.3f} ms")


if __name__ == "__main__":
    main()
