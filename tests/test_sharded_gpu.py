"""2-GPU NCCL run of the batch-sharded path (skipped with fewer than 2 GPUs): scatter -> engine forward on
each rank -> gather equals the single-GPU batch result bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ESR_ROOT"])
from oracle import esr_oracle as O
from ntire2022_esr_b200 import build_model
from ntire2022_esr_b200.sharded import forward_sharded
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
w = O.load_weights(os.path.join(os.environ["ESR_ROOT"], "tests", "golden", "weights", "rfdn_baseline.npz"))
m = build_model(0, state_dict=w).eval().to(f"cuda:{local}")
n = 5
g = torch.Generator().manual_seed(3)
imgs = (torch.rand(n, 3, 48, 40, generator=g) * 255).half()
out = forward_sharded(m, imgs.cuda() if rank == 0 else None, n, (3, 48, 40), torch.float16, torch.device(f"cuda:{local}"))
if rank == 0:
    ref = m(imgs.cuda())
    assert torch.equal(out, ref), "sharded result differs from the single-GPU batch"
    print("SHARDED_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_forward_sharded_nccl_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ESR_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK" in r.stdout
