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


def test_sharded_dataset_parts_are_independent_streams_of_one_law():
    """synth.make_part (bench.py --workload deep-sharded): every rank draws its shard on its own, all from one law."""
    import numpy as np

    from gbnns_dim_red_b200 import synth

    a = synth.make_part(4000, 24, 0, seed=5)
    assert np.array_equal(a, synth.make_part(4000, 24, 0, seed=5))          # deterministic per (seed, part)
    b = synth.make_part(4000, 24, 1, seed=5)
    assert not np.array_equal(a, b)
    # same latent map: the covariance spectra agree (8 strong directions + isotropic noise), and part 0's principal
    # subspace explains part 1 equally well
    ua, sa, _ = np.linalg.svd(a - a.mean(0), full_matrices=False)
    sb = np.linalg.svd(b - b.mean(0), compute_uv=False)
    assert np.allclose(sa[:8], sb[:8], rtol=0.1) and sa[8] < 0.2 * sa[7]
    va = np.linalg.svd(a - a.mean(0), full_matrices=False)[2][:8]
    resid = b - b.mean(0) - (b - b.mean(0)) @ va.T @ va
    assert (resid ** 2).sum() / ((b - b.mean(0)) ** 2).sum() < 0.05
    c = synth.make_part(4000, 24, 0, seed=6)                                 # another seed: another law
    vc = np.linalg.svd(c - c.mean(0), full_matrices=False)[2][:8]
    assert np.linalg.norm(va @ vc.T) < 0.95 * np.sqrt(8)
