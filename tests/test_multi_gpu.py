"""On-hardware parity of the multi-GPU data paths (needs >= 2 GPUs on the box; skipped otherwise): spawns one process per
GPU with torchrun and runs tests/_multi_gpu_worker.py, which compares every sharded / sliced result with the oracle."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_paths_against_oracle(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs on the box")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "_multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTI_GPU_REPORT ")]
    assert line, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.loads(line[-1][len("MULTI_GPU_REPORT "):])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_parity_n{world}.json"), "w"), indent=1)
    assert rep["ok"] and r.returncode == 0, json.dumps(rep, indent=1) + r.stderr[-2000:]
