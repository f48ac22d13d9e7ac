"""Multi-GPU SUMMA over NCCL (one process per GPU), checked tile by tile against the oracle.
Needs >= 2 GPUs on the box; the CPU (gloo, world size 2) counterpart of the host logic is
tests/test_summa_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    from tiledarray_b200 import device_count
    return device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(nproc, args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "_multi_gpu_worker.py")] + [str(a) for a in args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("MULTI_GPU_RESULT")]
    assert out.returncode == 0 and line, out.stdout[-2000:] + out.stderr[-4000:]
    return line[0]


@pytest.mark.parametrize("case", [(6, 5, 6, 128, 1.0), (8, 8, 8, 64, 0.4), (5, 7, 3, 100, 1.0)])
def test_summa_nccl(case):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 8 if n >= 8 else (4 if n >= 4 else 2)
    res = _run(nproc, list(case) + [0])
    assert "ok=True" in res, res


@pytest.mark.parametrize("case", [(6, 5, 6, 128, 1.0), (8, 8, 8, 64, 0.4)])
def test_summa_nccl_host_resident(case):
    """Same, with operands and result in pinned host memory (streamed panels + row blocks)."""
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 8 if n >= 8 else (4 if n >= 4 else 2)
    res = _run(nproc, list(case) + [2, "host"])
    assert "ok=True" in res, res
