"""Real multi-process, multi-GPU runs (torchrun, one process per GPU).  Skipped on boxes with a single GPU; the
same algorithms are covered there by the virtual-rank GPU test and the 2-rank gloo CPU test."""
import json
import os
import subprocess
import sys

import pytest
import torch

from tests.util import ROOT

pytestmark = pytest.mark.gpu
needs2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")


def _torchrun(n, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, script), *args]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


@needs2
def test_row_sharded_gram_over_nvlink_matches_fp64():
    lines = _torchrun(2, "tools/gram_dist.py", "--check-only")
    d = json.loads(lines[-1])
    assert d["world"] == 2 and max(d["small_rel_fro_err"]) < 1e-5
    assert all(d["nccl_allgather_same_bits"])      # library-collective baseline == planes ring, bit for bit


@needs2
def test_sharded_gram_of_real_per_sample_gradients_two_gpus():
    """config 5b end to end on 2 GPUs at a small shard size: rollout -> replay ring -> per-sample gradients -> one
    snk_gram_shard_run per rank; every rank checks >= 1000 off-diagonal entries against Float64 dot products"""
    lines = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "20", "--warmup", "3", "--envs", "65536", "--skip-config2",
                      "--skip-config4", "--skip-variants", "--e2e-steps", "2", "--gram-rows", "640")
    assert len(lines) == 1
    g = json.loads(lines[0])["gram_5b_sharded"]
    assert g["K_total"] == 1280 and g["nccl_allgather_same_bits_rank0"] is True
    v = g["verify_per_rank"]
    assert v["entries_per_rank"] >= 1000 and len(v["max_err_over_sqrt_GiiGjj"]) == 2
    assert max(v["max_err_over_sqrt_GiiGjj"]) < 1e-5


@needs2
def test_env_sharded_bench_line_two_gpus():
    lines = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "50", "--warmup", "5", "--envs", "65536", "--skip-gram",
                      "--skip-config2", "--skip-config4", "--e2e-steps", "2")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["config"]["global_envs"] == 131072
    assert d["config"]["env_errors"] == 0 and d["gpu_launches"] == 100
