"""N > 1 ranks on real GPUs (skipped on a one-GPU box): the sharded host path of
DeviceBasis.jk_direct -- row-sliced uploads + NCCL all-gather, row-sliced downloads into the host
buffer the ranks share (pychem_b200/dist.py NodeShare) -- against device-tensor inputs and the
one-rank result, for closed-shell, open-shell and general densities (tools/check_share_ngpu.py).
The 2- and 8-rank records of this check are under profiles/ (r2_share_check_8gpu.json)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_host_path_matches_plain_path_on_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "check_share_ngpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert rec["world"] == 2 and rec["shared_host_buffer"] is True
    for per_rank in rec["max_abs_diff_sharded_vs_device_inputs_per_rank"]:
        assert max(per_rank.values()) < 1.0e-12
    assert rec["max_abs_diff_vs_one_rank"] < 1.0e-11
