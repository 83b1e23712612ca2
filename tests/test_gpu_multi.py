"""Multi-GPU tests (need >= 2 B200 on the box: `gpurun --gpus 2`): the NCCL gradient all-reduce of the training step, run
through the product path under torchrun."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_fit_two_ranks_nccl():
    """Model.fit(distributed=DataParallel) on 2 GPUs: weights bit-identical across ranks, <= 2e-5 from the single-GPU run on
    the same global batches, final batch shorter than the number of ranks handled (tools/train_dp_check.py)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "train_dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["adam"]["ranks_identical"] and out["sgd"]["ranks_identical"] and out["dropout_ranks_identical"]
    assert out["p2p"] and out["adam_nccl"]["ranks_identical"] and out["adam_nccl"]["max_rel_dev_vs_single_gpu"] <= 2e-5 and out["status_flags"] == 0
    assert out["sgd"]["max_rel_dev_vs_single_gpu"] <= 2e-5 and out["adam_one_step_max_rel_dev"] <= 2e-5
