#!/bin/bash
# compute-sanitizer passes over the kernels that are new in round 2 (line-writer pack, TMA-staged permute, dist-mode epilogue + reducer,
# folded dangling labels, SVD with W warps per pair). Small cases only: the sanitizer slows kernels 10-100x.
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { echo "== $1 :: $2"; timeout 420 $S --tool $1 --error-exitcode 7 --print-limit 5 python -m pytest $2 -q -x -p no:cacheprovider 2>&1 | grep -E "passed|failed|error|ERROR SUMMARY|Invalid|Race|hazard|=========" | tail -6; }
run memcheck "tests/test_gpu.py -k test_permute_tma_variant_bit_exact"
run memcheck "tests/test_at_size.py -k test_fused_allreduce_emulated_ranks"
run memcheck "tests/test_gpu.py -k test_tcgen05_gather_pack_parity"
run memcheck "tests/test_at_size.py -k dangling"
run memcheck "tests/test_factorize.py -k rank_deficient"
run racecheck "tests/test_gpu.py -k test_tcgen05_gather_pack_parity"
run racecheck "tests/test_factorize.py -k rank_deficient"
