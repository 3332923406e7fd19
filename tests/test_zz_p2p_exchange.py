"""GPU (>= 2 devices): the peer-memory exchange of the sequence-sharded partial attention results, fused with their
merge (kvq_attend_exchange_merge, kvq_p2p.cu), against NCCL all_gather + kvq_attend_merge on the same inputs.

One process per GPU under torchrun (tests/_p2p_check.py): every rank stores its (out[H,128], lse[H]) straight into its
peers' IPC-mapped buffers over NVLink, waits for theirs and merges -- 200 rounds, so both buffer parities and the
device-resident sequence counter are exercised.  Skipped on a single-GPU box (a blocking N-rank exchange cannot be
emulated on one stream: the round-1 "loopback" version of this test could only time out)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_memory_exchange_equals_nccl_all_gather_merge(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(HERE, "_p2p_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("P2P_CHECK")]
    assert r.returncode == 0 and line, (r.stdout[-1500:], r.stderr[-1500:])
    assert float(line[0].split()[-1]) <= 1e-6, line[0]
