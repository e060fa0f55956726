"""GPU (needs >= 2 devices; skipped on a one-GPU box): DistributedTubeSection's exchange over NVLink peer memory
(sharded.PeerExchange: symmetric memory; copy engines, or the hb_peer_put kernel for small shards) delivers the same hit records and end states to rank 0 as the padded
NCCL gather on the same launch -- two ranks under torchrun, tools/gpu_probe_peer.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("cap", [128, 8])
def test_peer_exchange_equals_nccl_gather_on_two_ranks(cap):
    """cap = 8 recorded steps per trajectory: most trajectories outgrow the step scratch and are rerun with the fused kernel;
    their records (kept on the host by the runner) must reach rank 0 as well, through every form of the exchange (the kernel
    form then takes its host-sized second round)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "..", "tools", "gpu_probe_peer.py"), "40000", str(cap)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "peer exchange == nccl gather: OK" in res.stdout
