"""Multi-GPU check (needs >= 2 GPUs; skipped on the single-GPU test box): two ranks shard a population, exchange the
fitness values through the fused peer-memory gather (stito_eval_population_gather) -- and through NCCL with
STITO_PEER_GATHER=0 -- and must reproduce, bit for bit, what one rank computes for the same shards (bench.py's shard_check)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_ranks_reproduce_one_rank_bitwise(peer):
    env = dict(os.environ, STITO_PEER_GATHER=peer)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200 + int(peer)), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--config", "1",
           "--pop", "7", "--steps", "1", "--warmup", "1", "--iters", "3", "--no-cpu-baseline"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    assert line["shard_check"]["gathered_equals_single_rank_bitwise"] is True
    assert ("peer memory" in line["config"]["parallelism"]) == (peer == "1")
