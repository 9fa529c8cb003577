"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): per-rank shards + NCCL all-reduce
(Q1/Q6) and hash-partitioned all-to-all exchange (Q14) equal the single-GPU result."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_queries_match_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_multigpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines[0]["ok"] and lines[0]["planned_runs"] >= 1
    assert lines[1]["overflow_rerun_ok"] and lines[1]["replanned"]


@pytest.mark.parametrize("world", [2])
def test_exchange_operators_match_oracle(world):
    """Multi-fragment plans through B200PartitionedOutput / B200Exchange on `world` GPUs vs the oracle."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "check_multigpu_ops.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines[0]["ok"] and lines[0]["exchange_rows_sent"] == lines[0]["exchange_rows_received"] > 0


@pytest.mark.parametrize("world", [2])
def test_config5_partitioned_aggregation_matches_oracle(world):
    """BASELINE.json configs[4] at test size (10 M rows, 1 M distinct BIGINT keys): rows hash-partitioned by key
    through B200PartitionedOutput / B200Exchange, every rank aggregating its key slice — every group equals the
    CPU oracle's single-process result and every key lives on exactly one rank."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "scripts", "bench_config5_multi.py"), "--rows", "1e7", "--keys", "1e6",
           "--iters", "1", "--check"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")][-1]
    assert line["ok"] and line["groups_equal_oracle"] and line["keys_on_one_rank"] and line["distinct"] == line["oracle_groups"]
