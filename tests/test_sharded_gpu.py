"""GPU, >= 2 devices: one 256-frame clip frame-sharded over 2 ranks against the reference golden — both exchange
back-ends: the NCCL callback (eager) and the device-side peer-memory exchange (csrc/p2p_exchange.cu, CUDA graph).

The fixture is the "decisive" cfg-4 one: 40 of 256 frames are chosen in pass 1 and 20 in pass 2, so the selection counts, the
masked classifier / seed means and the temporal-attention K|V rows all really cross the ranks.  Each case spawns its ranks with
torch.distributed.run (tools/run_sharded.py prints PASS / FAIL on rank 0)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, p2p, port, golden=None):
    env = dict(os.environ, VGQA_SHARD_P2P="1" if p2p else "0", MASTER_ADDR="127.0.0.1")
    if golden:
        env["VGQA_SHARD_GOLDEN"] = golden
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "run_sharded.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    line = [ln for ln in out.splitlines() if ln.startswith("SHARDED")]
    assert line, out[-3000:]
    return line[-1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="frame sharding needs at least 2 GPUs")
@pytest.mark.parametrize("p2p", [False, True], ids=["nccl_callback", "peer_memory_graph"])
def test_sharded_clip_two_ranks_matches_reference_golden(p2p):
    line = _run(2, p2p, 29541 + int(p2p))
    assert "PASS" in line and "p2p_error=0" in line, line
    assert "selection_identical=True" in line and "sted_argmax_identical=True" in line, line


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_sharded_clip_four_ranks_peer_memory():
    line = _run(4, True, 29545)
    assert "PASS" in line and "p2p_error=0" in line, line
