"""N>1 on the GPU: the sharded meta-step of ``MetaOptimizer(_distributed=True)`` (kernels + one all-reduce of
[dtheta | fx] + identical Adam) against the single-GPU optimizer.  Runs on ANY box: NCCL with one rank per GPU when two
GPUs are visible, otherwise both ranks on cuda:0 with gloo as the transport (scripts/meta_dist_check.py prints which)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
def test_sharded_meta_step_matches_single_gpu():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547",
                        os.path.join(ROOT, "scripts", "meta_dist_check.py")], capture_output=True, text=True, timeout=800)
    print(r.stdout[-1500:])
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
