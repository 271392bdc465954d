"""world_size-2 gloo test of the launcher glue: the communicator id travels
from rank 0 to every rank through torch.distributed exactly as bench.py does
under torchrun (with NCCL there, gloo here)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from tmr_b200 import dist as tdist
dist.init_process_group("gloo")
rank, size = dist.get_rank(), dist.get_world_size()
t = torch.zeros(tdist.ID_BYTES, dtype=torch.uint8)
if rank == 0:
    t.copy_(torch.arange(tdist.ID_BYTES, dtype=torch.uint8))
dist.broadcast(t, src=0)
assert t.tolist() == list(range(tdist.ID_BYTES))
counts = tdist.partition_counts(1001, size)
assert sum(counts) == 1001 and counts[0] - counts[-1] in (0, 1)
gathered = [None] * size
dist.all_gather_object(gathered, counts[rank])
assert gathered == counts
dist.destroy_process_group()
print("rank", rank, "ok")
""" % ROOT


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "gloo_check.py"
    script.write_text(SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port",
           "29541", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
