"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2`): frames sharded over ranks, grid merged
by gv_grid_finalize_multi over NCCL, result bit-identical to the single-rank oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_finalize_multi_matches_oracle(world, mode):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + world + (50 if mode == "p2p" else 0)
    env = dict(os.environ, GV_MULTI_MODE=mode)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and f"MULTI_GPU_OK world={world} mode={mode}" in r.stdout, \
        r.stdout[-2000:] + r.stderr[-4000:]
