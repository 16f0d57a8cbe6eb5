"""Pieces over several B200s: the fused tail-GEMM -> NVLink peer store -> flag -> head-GEMM
handoff, wavefront micro-batch groups and the token return path must reproduce the
single-GPU result bit-for-bit (same kernels, same split-K, only the transport differs)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, model, groups, batch, port):
    env = dict(os.environ, B2B_MODEL=model, B2B_GROUPS=str(groups), B2B_BATCH=str(batch), MASTER_ADDR="127.0.0.1")
    if world == 1:
        cmd = [sys.executable, os.path.join(ROOT, "tools", "mp_check.py")]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mp_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])["tokens"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("model", ["tiny-llama", "tiny-gemma2", "tiny-gpt2"])
def test_two_gpu_pipeline_matches_single_gpu(model):
    # same total batch; 2 wavefront groups on the 2-rank mesh, and the same grouping on one rank
    ref = _run(1, model, 2, 4, 0)
    got = _run(2, model, 2, 4, 29611)
    assert got == ref


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs >= 4 GPUs")
def test_four_gpu_pipeline_matches_single_gpu():
    assert _run(4, "tiny-llama", 4, 2, 29613) == _run(1, "tiny-llama", 4, 2, 0)
