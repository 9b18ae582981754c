"""N > 1: x-slab decomposition over several GPUs of one box (one process per GPU, NCCL send/recv of
the face-crossing populations).  Needs >= 2 devices; on a 1-GPU box the same exchange logic is
covered by test_gpu_parity.py::test_slab_decomposition_is_bit_identical (several slabs in one process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(nproc, spr):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    env = dict(os.environ, PLB_TEST_SLABS_PER_RANK=str(spr))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc * 10 + spr), os.path.join(HERE, "mgpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_ranks():
    _run(2, 1)


def test_two_ranks_two_slabs_each():
    _run(2, 2)


def test_four_ranks():
    _run(4, 1)
