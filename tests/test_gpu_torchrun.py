"""GPU (-m gpu), >= 2 devices: the one-process-per-GPU deployment (torchrun; adept_cuda_join / share_tape /
collective adept_cuda_jacobian) that bench.py uses at N > 1 -- SURVEY.md 8e.  The worker is
tests/_torchrun_worker.py; this file only launches it and reports its output."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_process_per_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run with gpurun --gpus {world}; logs of such runs: profiles/r02_multi_*.log)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_torchrun_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert "comparisons bit-identical, error paths clean" in r.stdout
