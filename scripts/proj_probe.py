"""Prints projection errors / timings per mode (GPU debugging aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import capi, synth
for (d, dh, dl, nq) in [(128, 256, 32, 10000), (96, 128, 16, 777), (960, 1024, 32, 1000), (100, 72, 20, 131)]:
    rng = np.random.default_rng(1)
    q = rng.standard_normal((nq, d), dtype=np.float32)
    net = synth.make_net(d, dh, dl, seed=3)
    exact = synth.project_numpy(*net, q)
    ix = capi.Index(0)
    ix.set_net(*net)
    for mode in (2, 0, 1):
        ix.set_projection_mode(mode)
        try:
            got = ix.project(q)
            t0 = time.time()
            for _ in range(5):
                got = ix.project(q)
            dt = (time.time() - t0) / 5
            print(f"d={d} dh={dh} dl={dl} nq={nq} mode={mode} maxerr={np.abs(got-exact).max():.3e} t={dt*1e3:.3f} ms", flush=True)
        except Exception as e:
            print(f"d={d} mode={mode} FAILED: {e}", flush=True)
    ix.close()
