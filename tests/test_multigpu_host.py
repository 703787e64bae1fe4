"""Host-side multi-GPU logic (gbnns_dim_red_b200/multigpu.py) on CPU: world_size 2 and 3 over gloo."""
import os
import socket
import subprocess
import sys

import pytest

from gbnns_dim_red_b200 import multigpu as mg

HERE = os.path.dirname(os.path.abspath(__file__))


def test_partition_is_contiguous_balanced_and_complete():
    for n in (0, 1, 7, 10, 10000, 100_000_007):
        for world in (1, 2, 3, 8):
            parts = mg.partitions(n, world)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_workers(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "_mg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0 and "MG_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
