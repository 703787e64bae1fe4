"""The drop-in `final_test` binary on the bench's SIFT-1M workload, laid out on disk the way the reference expects it:
its printed work_time (seconds per query, number_exper repetitions per ef) against the bench's end-to-end numbers.
usage: python scripts/final_test_probe.py [out.json]   (GBDR_DEVICES=0,1,.. splits every batch over several GPUs)"""
import json, os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, ".")
from gbnns_dim_red_b200 import build, workload, xvecs

out_path = sys.argv[1] if len(sys.argv) > 1 else None
build.build_host()
w = workload.build_workload("sift1m", device=0, cache_dir=os.environ.get("GBDR_BENCH_CACHE", "/tmp/gbdr_bench_cache"))
sh = w["shape"]
ds, lat = "sift", "lat"
root = tempfile.mkdtemp(prefix="gbdr_ft_")
data, models, results = (os.path.join(root, x, ds) for x in ("data", "models", "results"))
for p in (data, models, results):
    os.makedirs(p)
t0 = time.time()
xvecs.write_fvecs(os.path.join(data, f"{ds}_base.fvecs"), w["base"])
xvecs.write_fvecs(os.path.join(data, f"{ds}_query.fvecs"), w["queries"])
xvecs.write_ivecs(os.path.join(data, f"{ds}_groundtruth.ivecs"), w["truth"])
xvecs.write_fvecs(os.path.join(data, f"{ds}_base_{lat}.fvecs"), w["db_low"])
for i, m in enumerate(w["net"], 1):
    xvecs.write_fvecs(os.path.join(models, f"{ds}_net_as_matrix_{lat}_{i}.fvecs"), m)
xvecs.write_edges(os.path.join(models, "gd_low.ivecs"), *w["graph"])
params = os.path.join(root, "params.txt")
efs = "20,40,53,60,100"
with open(params, "w") as f:
    f.write(f"{ds} n {sh['n']}\n{ds} n_q {sh['n_q']}\n{ds} n_tr {w['truth'].shape[1]}\n{ds} d {sh['d']}\n{ds} d_low {sh['d_low']}\n"
            f"{ds} d_hidden {sh['d_hidden']}\n{ds} efs {efs}\n{ds} efs_hnsw 10\n{ds} hnsw_name x\n")
print(f"files written in {time.time() - t0:.1f}s", flush=True)
env = dict(os.environ, GBDR_PARAMS=params, GBDR_DATA_ROOT=os.path.join(root, "data"), GBDR_MODELS_ROOT=os.path.join(root, "models"),
           GBDR_RESULTS_ROOT=os.path.join(root, "results"), GBDR_LAT_NAME=lat, GBDR_GRAPH_ORIG="none", GBDR_GRAPH_LOW="gd_low",
           GBDR_GRAPH_LOW_NAME="gd_low", GBDR_SEED="1234", GBDR_NUM_EXPER="5")
exe = os.path.join("gbnns_dim_red_b200", "host", "bin", "final_test")
t0 = time.time()
r = subprocess.run([exe, ds], env=env, capture_output=True, text=True)
print(r.stdout[-1500:], r.stderr[-500:], f"final_test: {time.time() - t0:.1f}s rc={r.returncode}", flush=True)
rows = []
for line in r.stdout.splitlines():
    t = line.split(" ")
    if t[0] == "graph_type" and len(t) == 10:
        rows.append(dict(graph=t[1], acc=float(t[3]), hops=int(t[5]), dist_calc=int(t[7]), work_time=float(t[9]),
                         qps=1.0 / float(t[9])))
for ef, row in zip(efs.split(","), rows):
    row["ef"] = int(ef)
    print(f"ef {ef}: acc {row['acc']:.4f} work_time {row['work_time']:.3e} s/query = {row['qps'] / 1e6:.2f} M QPS")
if out_path:
    json.dump(dict(devices=os.environ.get("GBDR_DEVICES", "0"), number_exper=5, rows=rows), open(out_path, "w"), indent=1)
subprocess.run(["rm", "-rf", root])
