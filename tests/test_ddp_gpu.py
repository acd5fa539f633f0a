"""Data parallelism on real GPUs (needs >= 2 visible devices; skipped on a 1-GPU box): tools/check_ddp_gpu.py under
torch.distributed.run -- bucketed NCCL all-reduce with the wgrad kernels writing straight into the buckets == the plain model
on the whole batch, for ViT, TiTok, VideoGPT and the blocks.py tokenizer; gradient accumulation after no_sync(); the
bf16-compressed buckets."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_ddp_gradients_equal_single_process_on_two_gpus():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", os.path.join(ROOT, "tools", "check_ddp_gpu.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DDP gradient check OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
