"""-m gpu, needs >= 2 GPUs on the box (skipped otherwise): the one-kernel exchange step (csrc/exchange.cu) against NCCL."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_allreduce_matches_nccl_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "exchange_probe.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300).stdout
    lines = [l for l in out.splitlines() if l.startswith(("multimem:", "p2p:", "p2p-oddP:", "cpp-multimem:", "cpp-p2p:"))]
    assert any(l.startswith("p2p:") for l in lines) and any(l.startswith("p2p-oddP:") for l in lines), out
    if os.path.exists(os.path.join(ROOT, "adapter", "gsb_adapter.so")):   # the C++ (libtorch) host of the kernel, adapter/Exchange.{h,cc}
        assert any(l.startswith("cpp-p2p:") for l in lines), out
    for l in lines:   # the sum is computed once per element and broadcast: equal to NCCL up to summation order, identical on all ranks
        err = float(l.split("max|err| vs NCCL")[1].split(",")[0])
        assert err <= 1e-5 and "identical across ranks True" in l, l
