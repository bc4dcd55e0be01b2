"""GPU, N > 1: the sharded path (query-read shards; index built partition-wise and replicated; adjacency rows pushed to the peers
chunk by chunk behind the verification; NCCL allgather of the overflow entries, verdict bits and final edges) gives every rank the
oracle's graph -- on the small adversarial sets, edge for edge, and at scale (config 2 / config 3 x world, config 4 at full size on 8
ranks) through counters and the order-independent checksum of the lean oracle's goldens. Needs >= 2 GPUs (`gpurun --gpus N`); the cases a
box cannot run are skipped. Logs of the round's runs: profiles/r2/tests_2gpu_mid.log, profiles/r2/tests_8gpu.log."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(world, script, *args, timeout=1800):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", script)] + [str(a) for a in args]
    return subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)


# (world, config, scale): the weak-scaled bench workloads (config 3 x world) and config 2 x world -- sub-partitioned index,
# rows of hundreds of MB through C1, the retry paths -- and, on 8 GPUs, BASELINE.json configs[3] at full size (50 M x 150 bp)
BIG = [(2, 2, 2.0), (2, 3, 2.0), (4, 2, 4.0), (4, 3, 4.0), (8, 2, 8.0), (8, 3, 8.0), (8, 4, 1.0)]


@pytest.mark.parametrize("world,config,scale", BIG, ids=[f"{w}gpu-config{c}@{s}" for w, c, s in BIG])
def test_sharded_build_matches_golden_at_scale(world, config, scale):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _torchrun(world, "dist_worker_big.py", config, scale, timeout=3000)
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.count("identical to the golden") == world, r.stdout[-2000:]
    print(r.stdout[-1500:])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_build_matches_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.count("identical to the oracle") == world, r.stdout[-2000:]
