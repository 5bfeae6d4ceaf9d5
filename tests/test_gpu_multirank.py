"""Multi-rank GPU test (`-m gpu`, needs >= 2 GPUs on the box; skipped otherwise): one process per GPU under
torch.distributed.run, NCCL.  Runs tools/check_sharded_nccl.py, which checks the sharded paths -- the callback forms
and the collectives inside libshiftedprox (spx_comm_*) -- against the single-device results on the same vectors."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("peer", ["1", "0"], ids=["scalars_over_peer_memory", "scalars_over_nccl"])
def test_sharded_paths_over_nccl(peer):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (one process per GPU); covered by the world_size-2 gloo tests on the CPU")
    world = 2 if ngpu < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "check_sharded_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=580, env=dict(os.environ, SPX_PEER=peer))
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARDED CHECK OK" in r.stdout
    assert f"peer-memory path {peer == '1'}" in r.stdout  # the path asked for is the one that ran
