"""-m gpu, needs >= 2 GPUs on the box (skipped otherwise; the driver's single-GPU test box skips it, `gpurun --gpus 2`
runs it): the data-parallel train step over the peer-memory gradient exchange (csrc/peer_allreduce.cu).
tools/dp_check.py is launched with torchrun on 2 ranks and checks that (1) after backward every replica's gradient arena
is bit-identical and equals the SUM of the per-replica gradients computed by an independent model copy without exchange
(gathered with NCCL) to weight-gradient summation-order accuracy, (2) after six optimisation steps (eager, then
CUDA-graph replays with the exchange kernels inside the graph) all replicas hold bit-identical parameters."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("multicast", ["1", "0"])
def test_data_parallel_peer_exchange_two_replicas(multicast):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, DP_TIMING="0", CRIMAC_AR_MULTICAST=multicast)
    port = str(29600 + os.getpid() % 300 + (7 if multicast == "1" else 0))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "dp_check.py")],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    d = json.loads(lines[-1])
    assert d["ok"] and d["world"] == 2 and d["exchange"] == "peer"
    assert d["grad_arenas_bit_identical"] and d["params_bit_identical_after_6_steps"] and d["graph_replay"]
    assert d["grad_sum_rel_err"] < 1e-5
