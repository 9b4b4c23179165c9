"""The peer-memory exchange kernels (csrc/p2p_allreduce.cu: vn_p2p_allreduce, vn_p2p_allreduce_small, vn_p2p_reduce_adam,
vn_p2p_step) need >= 2 GPUs with peer access: tools/p2p_test.py is launched under torchrun with 2 ranks and checks them
against NCCL and against allreduce + the dense optimiser kernels (bit-identical parameters on every rank, incl. steps with
an injected inf / nan).  Skipped on a 1-GPU box; the protocol itself is covered on CPU by tests/test_dp_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs with peer access")
def test_p2p_kernels_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "p2p_test.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("world 2")]
    assert any("one-kernel step 5" in l and "params == reference True" in l for l in lines), out.stdout[-3000:]
    assert not any("False" in l for l in lines), "\n".join(lines)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs with peer access")
def test_dp_engine_two_ranks_strong_scaling_matches_one_rank():
    """bench.py --scaling strong: one globally seeded batch split over 2 ranks trains like the 1-rank run (same losses
    within the gradient tolerance of the atomics' summation order)"""
    import json
    env = {**os.environ, "VN_P2P_TIMEOUT_MS": "5000"}
    common = ["--steps", "6", "--warmup", "3", "--windows", "1", "--no-extra", "--no-cpu-baseline", "--no-e2e", "--scaling", "strong",
              "--rays", "4096"]
    one = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + common, capture_output=True, text=True, timeout=600,
                         cwd=ROOT, env={**env, "WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0"})
    assert one.returncode == 0, one.stderr[-3000:]
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29632", os.path.join(ROOT, "bench.py"), "--gpus", "2"] + common,
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert two.returncode == 0, two.stderr[-3000:]
    l1 = json.loads([l for l in one.stdout.splitlines() if l.startswith('{"metric"')][-1])
    l2 = json.loads([l for l in two.stdout.splitlines() if l.startswith('{"metric"')][-1])
    assert l2["n_gpus"] == 2 and l2["scaling"] == "strong" and l2["replicas_bit_identical"]
    assert l2["config"]["global_rays_per_step"] == l1["config"]["global_rays_per_step"] == 4096
    assert abs(l1["final_loss"] - l2["final_loss"]) <= 2e-3 * abs(l1["final_loss"]), (l1["final_loss"], l2["final_loss"])
