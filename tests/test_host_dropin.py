"""The C++ host layer (gbnns_dim_red_b200/host): the reference's drivers must build against the
drop-in header, and on a GPU the drop-in binaries must reproduce the oracle's numbers on files laid
out exactly as the reference expects them."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from gbnns_dim_red_b200 import build, synth, xvecs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "gbnns_dim_red_b200", "host")
REF = "/root/reference/search"


def test_drivers_are_built_and_linked_against_the_library():
    build.build_host()
    for name in ("final_test", "prepare_graph"):
        exe = os.path.join(HOST, "bin", name)
        assert os.access(exe, os.X_OK)
        out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "libgbdr.so" in out and "not found" not in out
        # wrong argc: message + exit code 1, like the reference (final_test.cpp:10-15)
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 1 and "Need to specify parameters" in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree absent")
@pytest.mark.parametrize("name", ["final_test", "prepare_graph"])
def test_unmodified_reference_driver_builds_against_dropin_header(tmp_path, name):
    """Source-level drop-in: the reference's own main(), copied to a scratch directory so that its
    `#include "search_function.h"` resolves to host/search_function.h, compiles and links."""
    src = tmp_path / f"{name}.cpp"
    shutil.copy(os.path.join(REF, f"{name}.cpp"), src)
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++11", "-w", "-fopenmp", "-I", HOST, "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(tmp_path / name), "-L", os.path.join(ROOT, "gbnns_dim_red_b200"), "-lgbdr"],
                   check=True)


def _write_dataset(root, ds, lat, c, base_graph):
    data = os.path.join(root, "data", ds)
    models = os.path.join(root, "models", ds)
    results = os.path.join(root, "results", ds)
    for p in (data, models, results):
        os.makedirs(p, exist_ok=True)
    xvecs.write_fvecs(os.path.join(data, f"{ds}_base.fvecs"), c["base"])
    xvecs.write_fvecs(os.path.join(data, f"{ds}_query.fvecs"), c["queries"])
    xvecs.write_ivecs(os.path.join(data, f"{ds}_groundtruth.ivecs"), c["truth"])
    xvecs.write_fvecs(os.path.join(data, f"{ds}_base_{lat}.fvecs"), c["db_low"])
    xvecs.write_ivecs(os.path.join(models, f"{ds}_knn_1k_{lat}.ivecs"), c["knn_ids"])
    for i, m in enumerate(c["net"], 1):
        xvecs.write_fvecs(os.path.join(models, f"{ds}_net_as_matrix_{lat}_{i}.fvecs"), m)
    xvecs.write_edges(os.path.join(models, "orig_graph.ivecs"), *base_graph)
    params = os.path.join(root, "params.txt")
    with open(params, "w") as f:
        f.write(f"{ds} n {c['n']}\n{ds} n_q {c['n_q']}\n{ds} n_tr {c['truth'].shape[1]}\n{ds} d {c['d']}\n"
                f"{ds} d_low {c['d_low']}\n{ds} d_hidden {c['dh']}\n{ds} efs 4,16,40\n{ds} efs_hnsw 8,30\n{ds} hnsw_name x\n")
    env = dict(os.environ, GBDR_PARAMS=params, GBDR_DATA_ROOT=os.path.join(root, "data"),
               GBDR_MODELS_ROOT=os.path.join(root, "models"), GBDR_RESULTS_ROOT=os.path.join(root, "results"),
               GBDR_LAT_NAME=lat, GBDR_NUM_EXPER="2")
    return env, models, results


def _parse(line):
    t = line.split(" ")
    assert t[0] == "graph_type" and len(t) == 10
    return dict(name=t[1], acc=float(t[3]), hops=int(t[5]), dist_calc=int(t[7]), work_time=float(t[9]))


@pytest.mark.gpu
def test_prepare_graph_and_final_test_end_to_end(tmp_path):
    from . import _oracle as O
    from ._data import small_case

    build.build_host()
    c = small_case()
    ds, lat = "toy", "lat"
    base_knn, _ = O.orc_knn(c["base"], c["base"], 40)
    bg = O.orc_gd_prune(*xvecs.adjacency_from_matrix(base_knn), c["base"], M=8, reverse=True)
    env, models, results = _write_dataset(str(tmp_path), ds, lat, c, bg)

    # ---- prepare_graph: GD graph file must equal the oracle's graph, byte for byte
    env_pg = dict(env, GBDR_GD_M=str(c["M"]))
    r = subprocess.run([os.path.join(HOST, "bin", "prepare_graph"), ds, lat], env=env_pg, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    goff, ged = xvecs.read_edges(os.path.join(models, f"{ds}_gd_knn_{lat}.ivecs"), n=c["n"])
    assert np.array_equal(goff, c["graph"][0]) and np.array_equal(ged, c["graph"][1])
    assert f"GD_knn {int(ged.size / c['n'])}" in r.stdout

    # ---- prepare_graph with the kNN file missing: kNN lists and GD graph built in one HBM-resident chain
    #      (gbdr_build_graph), both files identical to the oracle's
    os.remove(os.path.join(models, f"{ds}_knn_1k_{lat}.ivecs"))
    os.remove(os.path.join(models, f"{ds}_gd_knn_{lat}.ivecs"))
    r = subprocess.run([os.path.join(HOST, "bin", "prepare_graph"), ds, lat],
                       env=dict(env_pg, GBDR_KNN_K=str(c["knn_ids"].shape[1])), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert np.array_equal(xvecs.read_ivecs(os.path.join(models, f"{ds}_knn_1k_{lat}.ivecs")), c["knn_ids"])
    goff, ged = xvecs.read_edges(os.path.join(models, f"{ds}_gd_knn_{lat}.ivecs"), n=c["n"])
    assert np.array_equal(goff, c["graph"][0]) and np.array_equal(ged, c["graph"][1])
    assert f"GD_knn {int(ged.size / c['n'])}" in r.stdout

    # ---- final_test: both sweeps, entry vertex 0 (graph labels starting with "hnsw"), result lines
    env_ft = dict(env, GBDR_GRAPH_ORIG="orig_graph", GBDR_GRAPH_LOW=f"{ds}_gd_knn_{lat}", GBDR_GRAPH_LOW_NAME="hnsw_gd")
    r = subprocess.run([os.path.join(HOST, "bin", "final_test"), ds], env=env_ft, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [_parse(x) for x in open(os.path.join(results, f"final_results_{ds}.txt")).read().splitlines()]
    assert [x["name"] for x in lines] == ["hnsw", "hnsw", "hnsw_gd", "hnsw_gd", "hnsw_gd"]
    assert [x for x in r.stdout.splitlines() if x.startswith("graph_type")] == \
        open(os.path.join(results, f"final_results_{ds}.txt")).read().splitlines()
    entry = np.zeros(c["n_q"], np.uint32)
    n_q, truth = c["n_q"], c["truth"]
    for line, ef in zip(lines[:2], (8, 30)):   # original-dimension search (d == d_low branch)
        o = O.orc_search(c["queries"], None, c["base"], None, bg[0], bg[1], ef, 1, 2, entry)
        assert line["hops"] == int(o["hops"].sum()) // n_q
        assert line["dist_calc"] == int(o["dist_calc"].sum()) // n_q
        assert abs(line["acc"] - float((o["ids"][:, 0] == truth[:, 0]).mean())) < 1e-6
    q_low = O.orc_project(*c["net"], c["queries"])
    for line, ef in zip(lines[2:], (4, 16, 40)):  # projected search + re-rank (performNetTest)
        o = O.orc_search(c["queries"], q_low, c["base"], c["db_low"], c["graph"][0], c["graph"][1], ef, 1, 0, entry)
        # the GPU projects with 3xTF32 (<= 1e-5 from the oracle's fp32): allow a query or two to walk differently
        assert abs(line["hops"] - int(o["hops"].sum()) // n_q) <= 1
        assert abs(line["dist_calc"] - int(o["dist_calc"].sum()) // n_q) <= 2
        assert abs(line["acc"] - float((o["ids"][:, 0] == truth[:, 0]).mean())) <= 2.0 / n_q
        assert line["work_time"] > 0


@pytest.mark.gpu
def test_second_graph_through_dropin_header(tmp_path):
    """performRealTests(..., use_second_graph = true, llf) as naive_test.cpp:102-105 calls it, through the drop-in
    header: result lines must carry the oracle's hops / dist_calc / accuracy."""
    from . import _oracle as O
    from ._data import long_link_graph, small_case

    c = small_case()
    aux = long_link_graph(c["n"])
    t = str(tmp_path)
    xvecs.write_fvecs(f"{t}/base.fvecs", c["base"])
    xvecs.write_fvecs(f"{t}/query.fvecs", c["queries"])
    xvecs.write_ivecs(f"{t}/truth.ivecs", c["truth"])
    xvecs.write_fvecs(f"{t}/base_low.fvecs", c["db_low"])
    xvecs.write_fvecs(f"{t}/query_low.fvecs", c["q_low"])
    xvecs.write_edges(f"{t}/main.edges", *c["graph"])
    xvecs.write_edges(f"{t}/aux.edges", *aux)
    exe = f"{t}/second_graph_driver"
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++11", "-w", "-fopenmp", "-I", HOST, "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "second_graph_driver.cpp"), "-o", exe, "-L",
                    os.path.join(ROOT, "gbnns_dim_red_b200"), "-lgbdr",
                    "-Wl,-rpath," + os.path.join(ROOT, "gbnns_dim_red_b200")], check=True)
    r = subprocess.run([exe, f"{t}/base.fvecs", f"{t}/query.fvecs", f"{t}/truth.ivecs", f"{t}/base_low.fvecs",
                        f"{t}/query_low.fvecs", f"{t}/main.edges", f"{t}/aux.edges", f"{t}/out.txt", str(c["n"]),
                        str(c["d"]), str(c["d_low"]), str(c["n_q"]), str(c["truth"].shape[1])],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [_parse(x) for x in open(f"{t}/out.txt").read().splitlines()]
    assert [x["name"] for x in lines] == ["hnsw_lk_low"] * 2 + ["hnsw_lk_low_nollf"] * 2 + ["hnsw_low"] * 2
    entry = np.zeros(c["n_q"], np.uint32)
    n_q, truth = c["n_q"], c["truth"]
    goff, ged = c["graph"]
    want = [(dict(aux=aux, llf=True), 6), (dict(aux=aux, llf=True), 30), (dict(aux=aux, llf=False), 6),
            (dict(aux=aux, llf=False), 30), ({}, 6), ({}, 30)]
    for line, (kw, ef) in zip(lines, want):
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, entry, hops_bound=50, **kw)
        assert line["hops"] == int(o["hops"].sum()) // n_q
        assert line["dist_calc"] == int(o["dist_calc"].sum()) // n_q
        assert abs(line["acc"] - float((o["ids"][:, 0] == truth[:, 0]).mean())) < 1e-6
    assert lines[0]["dist_calc"] != lines[4]["dist_calc"]


@pytest.mark.gpu
def test_wrap_c_support_dropin(tmp_path, monkeypatch, capsys):
    """wrap.c_support.get_graphs_and_search_tests on the file layout the trainers write
    (dim_red/triplet.py:142-153): same GD graph and search numbers as the oracle, returns 0."""
    from gbnns_dim_red_b200.wrap import c_support

    from . import _oracle as O
    from ._data import small_case

    c = small_case()
    n, n_q = c["n"], c["n_q"]
    data = tmp_path / "data" / "sift"
    models = tmp_path / "models" / "sift"
    data.mkdir(parents=True)
    models.mkdir(parents=True)
    truth, _ = O.orc_knn(c["queries"], c["base"], 100)
    xvecs.write_fvecs(data / "sift_base_valid.fvecs", c["base"])
    xvecs.write_fvecs(data / "sift_query_valid.fvecs", c["queries"])
    xvecs.write_ivecs(data / "sift_groundtruth_valid.ivecs", truth)
    xvecs.write_fvecs(data / "sift_base_triplet_wrap_valid.fvecs", c["db_low"])
    xvecs.write_fvecs(data / "sift_query_triplet_wrap_valid.fvecs", c["q_low"])
    xvecs.write_ivecs(models / "knn_1k_triplet_wrap_valid.ivecs", c["knn_ids"])
    monkeypatch.setenv("GBDR_DATA_ROOT", str(tmp_path / "data"))
    monkeypatch.setenv("GBDR_MODELS_ROOT", str(tmp_path / "models"))
    monkeypatch.setenv("GBDR_TRAIN_RESULTS_ROOT", str(tmp_path / "results"))
    monkeypatch.setenv("GBDR_SEED", "5")
    rc = c_support.get_graphs_and_search_tests("t", "s", c["d"], c["d_low"], n_q, "v", n, False, "ignored-9th-arg")
    assert rc == 0
    (ef, acc, hops, dist_calc, work), = c_support.last_results()
    assert ef == 150
    goff, ged = O.orc_gd_prune(*c["knn"], c["db_low"], M=20, reverse=False)
    entry = np.random.default_rng(5).integers(0, n, size=n_q, dtype=np.uint32)
    o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, 150, 1, 0, entry)
    assert hops == int(o["hops"].sum()) // n_q and dist_calc == int(o["dist_calc"].sum()) // n_q
    assert acc == float((o["ids"][:, 0] == truth[:, 0]).mean())
    line = open(tmp_path / "results" / "sift" / "train_results_triplet_wrap.txt").read().strip()
    assert line.startswith("graph_type gd_knn_20 acc ") and len(line.split(" ")) == 10
    assert f"GD_knn_low {int(ged.size / n)}" in capsys.readouterr().out


@pytest.mark.gpu
def test_training_loop_hooks_in_memory(tmp_path):
    """the in-memory form of the wrap.c_support check and the kNN helpers of dim_red/support_func.py."""
    from gbnns_dim_red_b200.wrap import c_support, support_func

    from . import _oracle as O
    from ._data import small_case

    c = small_case()
    n, n_q = c["n"], c["n_q"]
    # kNN helpers: int64 ids like faiss' I; blocks + ivecs dump; every row answered (no len % 500 remainder dropped)
    knn = support_func.get_nearestneighbors(c["db_low"], c["db_low"], 100, "cuda")
    assert knn.dtype == np.int64 and np.array_equal(knn.astype(np.uint32), c["knn_ids"])
    path = str(tmp_path / "knn.ivecs")
    part = support_func.get_nearestneighbors_partly(c["db_low"][:777], c["db_low"], 100, "cuda", bs=250, path=path)
    assert np.array_equal(part, knn[:777]) and np.array_equal(xvecs.read_ivecs(path), part.astype(np.uint32))
    gt = support_func.get_nearestneighbors(c["queries"], c["base"], 10, "cuda", needs_exact=False)
    assert np.array_equal(gt.astype(np.uint32), c["truth"])

    # search check straight from arrays: same numbers as the oracle, and what the early-stopping code wants back
    res = c_support.search_tests(c["base"], c["queries"], c["truth"], c["db_low"], c["q_low"], knn.astype(np.uint32),
                                 [20, 60], M=12, reverse_gd=True, seed=9, output_txt=str(tmp_path / "res.txt"))
    assert [r[0] for r in res] == [20, 60] and res == c_support.last_results()
    goff, ged = O.orc_gd_prune(*c["knn"], c["db_low"], M=12, reverse=True)
    entry = np.random.default_rng(9).integers(0, n, size=n_q, dtype=np.uint32)
    for (ef, acc, hops, dist_calc, work) in res:
        o = O.orc_search(c["queries"], c["q_low"], c["base"], c["db_low"], goff, ged, ef, 1, 0, entry)
        assert hops == int(o["hops"].sum()) // n_q and dist_calc == int(o["dist_calc"].sum()) // n_q
        assert acc == float((o["ids"][:, 0] == c["truth"][:, 0]).mean()) and work > 0
    assert len(open(tmp_path / "res.txt").read().splitlines()) == 2


WALK_SRC = r"""
// a caller-driven walk over makeStep / VisitedListPool, the way getOneSearchResults uses them (search_function.h:43-102)
#include "search_function.h"
int main(int argc, char** argv) {
    const int n = atoi(argv[1]), d = atoi(argv[2]), n_q = atoi(argv[3]), ef = atoi(argv[4]);
    vector<float> db = loadXvecs<float>(argv[5], d, n);
    vector<float> q = loadXvecs<float>(argv[6], d, n_q);
    vector<vector<uint32_t>> g = loadEdges(argv[7], n, "graph");
    L2Metric l2;
    VisitedListPool pool(1, n);
    for (int i = 0; i < n_q; ++i) {
        VisitedList* vl = pool.getFreeVisitedList();
        priority_queue<pair<float, int>> top, cand;
        const float* query = q.data() + (size_t)i * d;
        int dist_calc = 1, hops = 0, k = 1, e = ef;
        const int entry = (i * 7919) % n;
        const float d0 = l2.Dist(query, db.data() + (size_t)entry * d, d);
        top.emplace(d0, entry);
        cand.emplace(-d0, entry);
        vl->mass[entry] = vl->curV;
        while (!cand.empty()) {
            pair<float, int> c = cand.top();
            if (-c.first > top.top().first) break;
            cand.pop();
            bool found = false;
            makeStep(g[c.second], query, db.data(), top, cand, &l2, d, dist_calc, found, e, k, vl);
            ++hops;
        }
        pool.releaseVisitedList(vl);
        printf("%d %d %d", i, hops, dist_calc);
        while (!top.empty()) { printf(" %d:%.9g", top.top().second, top.top().first); top.pop(); }
        printf("\n");
    }
    return 0;
}
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree absent")
def test_make_step_and_visited_pool_match_the_reference(tmp_path):
    """makeStep / VisitedList / VisitedListPool of the drop-in header are host code (no device needed): the same
    caller-driven walk compiled against the reference's headers and against ours prints the same lines."""
    from ._data import small_case

    c = small_case()
    n, d = c["db_low"].shape
    n_q = 300
    xvecs.write_fvecs(str(tmp_path / "db.fvecs"), c["db_low"])
    xvecs.write_fvecs(str(tmp_path / "q.fvecs"), np.tile(c["q_low"], (2, 1))[:n_q])
    xvecs.write_edges(str(tmp_path / "g.ivecs"), *c["graph"])
    (tmp_path / "walk.cpp").write_text(WALK_SRC)
    outs = []
    for tag, inc, link in (("ref", ["-I", REF], []),
                           ("ours", ["-I", HOST, "-I", os.path.join(ROOT, "include")],
                            ["-L", os.path.join(ROOT, "gbnns_dim_red_b200"), "-lgbdr",
                             "-Wl,-rpath," + os.path.join(ROOT, "gbnns_dim_red_b200")])):
        exe = tmp_path / f"walk_{tag}"
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++11", "-w", "-fopenmp", "-march=x86-64-v3", "-fno-fast-math",
                        "-ffp-contract=off", *inc, str(tmp_path / "walk.cpp"), "-o", str(exe), *link], check=True)
        r = subprocess.run([str(exe), str(n), str(d), str(n_q), "24", str(tmp_path / "db.fvecs"), str(tmp_path / "q.fvecs"),
                            str(tmp_path / "g.ivecs")], capture_output=True, text=True, check=True)
        outs.append([l for l in r.stdout.splitlines() if l and l[0].isdigit()])
    assert len(outs[0]) == n_q and outs[0] == outs[1]


def test_trainer_import_path_resolves_unedited():
    """dim_red/triplet.py:143 and dim_red/angular.py:180 say `import wrap.c_support`: with the repository root on sys.path
    that name resolves to the top-level wrap/ shim, which forwards to the package's module (same callables)."""
    import importlib
    import sys

    sys.modules.pop("wrap", None)
    sys.modules.pop("wrap.c_support", None)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    shim = importlib.import_module("wrap.c_support")
    from gbnns_dim_red_b200.wrap import c_support as impl

    assert shim.get_graphs_and_search_tests is impl.get_graphs_and_search_tests
    assert shim.last_results is impl.last_results and shim.search_tests is impl.search_tests


@pytest.mark.gpu
def test_dropin_binaries_over_several_gpus(tmp_path):
    """GBDR_DEVICES=0,1: prepare_graph builds the graph row-block sharded (gbdr_group_build_graph) and final_test splits
    every batch over the GPUs (gbdr_group_search on a replicated group).  Same files and same result lines as one GPU."""
    from gbnns_dim_red_b200 import capi

    from . import _oracle as O
    from ._data import small_case

    if capi.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    build.build_host()
    c = small_case()
    ds, lat = "toy", "lat"
    base_knn, _ = O.orc_knn(c["base"], c["base"], 40)
    bg = O.orc_gd_prune(*xvecs.adjacency_from_matrix(base_knn), c["base"], M=8, reverse=True)
    env, models, results = _write_dataset(str(tmp_path), ds, lat, c, bg)
    os.remove(os.path.join(models, f"{ds}_knn_1k_{lat}.ivecs"))
    env2 = dict(env, GBDR_DEVICES="0,1", GBDR_GD_M=str(c["M"]), GBDR_KNN_K=str(c["knn_ids"].shape[1]))
    r = subprocess.run([os.path.join(HOST, "bin", "prepare_graph"), ds, lat], env=env2, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "(2 GPU)" in r.stdout
    assert np.array_equal(xvecs.read_ivecs(os.path.join(models, f"{ds}_knn_1k_{lat}.ivecs")), c["knn_ids"])
    goff, ged = xvecs.read_edges(os.path.join(models, f"{ds}_gd_knn_{lat}.ivecs"), n=c["n"])
    assert np.array_equal(goff, c["graph"][0]) and np.array_equal(ged, c["graph"][1])

    lines = {}
    for tag, e in (("one", env), ("two", dict(env, GBDR_DEVICES="0,1"))):
        e = dict(e, GBDR_GRAPH_ORIG="orig_graph", GBDR_GRAPH_LOW=f"{ds}_gd_knn_{lat}", GBDR_GRAPH_LOW_NAME="gd", GBDR_SEED="7")
        r = subprocess.run([os.path.join(HOST, "bin", "final_test"), ds], env=e, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        lines[tag] = [_parse(x) for x in open(os.path.join(results, f"final_results_{ds}.txt")).read().splitlines()]
    assert len(lines["one"]) == len(lines["two"]) == 5
    for a, b in zip(lines["one"], lines["two"]):
        assert (a["name"], a["acc"], a["hops"], a["dist_calc"]) == (b["name"], b["acc"], b["hops"], b["dist_calc"])
