#!/usr/bin/env python
"""bench.py — QPS at recall@1 = 0.95 on the SIFT-1M shape (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload sift1m|c1|deep1m|gist1m|deep-sharded]

A "step" is one pass of the hot path (query projection -> low-dim beam search -> original-dim
re-rank, top-1) over one batch of n_q synthetic queries.

  value     device-resident leg: queries already in HBM, K steps, CUDA events, max over ranks.  With
            --in-flight 4 (default) the steps rotate over the index and views of it
            (gbdr_index_create_view: same resident data, own stream + workspaces), so the drain of one batch's
            persistent search kernel overlaps the start of the next batch; `single_stream` holds the same K
            steps issued back to back on one stream
  e2e       the same step through the host-facing C ABI with pinned HOST buffers (H2D of queries + entry
            points and D2H of ids/dists/hops/dist_calc inside the timed region, every step): gbdr_search_submit /
            gbdr_search_wait with --in-flight batches outstanding; `sync` holds the blocking gbdr_search loop
  roofline  beam-search kernel (dominant): algorithmic bytes per launch / its mean CUDA-event duration in the
            single-stream leg (per-launch durations of overlapped kernels would include waiting for SMs)
  cpu_baseline  the reference's own performTest (OpenMP, all host threads) on a bounded query sample

`--impl reference` times the reference's CPU code (oracle/_ref, built from /root/reference) on the
same workload/ef rule; under torchrun only rank 0 runs it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "qps_at_recall1_0.95_sift1m_dlow32"
UNIT = "queries/s"
EFS = [1, 3, 8, 15, 20, 25, 40, 60, 80, 100, 120, 140, 160, 180, 300, 500]  # parameters_of_databases.txt:7 (+300,500)
TARGET_RECALL = 0.95
E2E_REPS = 5  # repetitions of the K-step host-timed loops (median reported)
VALUE_REPS = 5  # repetitions of the K-step device-timed loop behind `value` (median over repetitions of the max over ranks)


def metric_name(workload):
    """BASELINE.json's metric on its own workload; the other configs are labelled by theirs."""
    return METRIC if workload == "sift1m" else f"qps_at_recall1_0.95_{workload}"


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------- ef rule
def pick_ef(recall_of, efs=EFS, target=TARGET_RECALL):
    """Smallest ef with recall@1 >= target: coarse sweep over the reference's ef list, then integer
    bisection inside the bracket.  Returns (ef, recall, bracket)."""
    prev = None
    for ef in efs:
        r = recall_of(ef)
        log(f"  ef={ef:4d} recall@1={r:.4f}")
        if r >= target:
            lo, hi, rhi = (prev[0] if prev else 0), ef, r
            bracket = [prev, (ef, r)]
            while hi - lo > 1:
                mid = (lo + hi) // 2
                rm = recall_of(mid)
                log(f"  ef={mid:4d} recall@1={rm:.4f} (bisect)")
                if rm >= target:
                    hi, rhi = mid, rm
                else:
                    lo = mid
            return hi, rhi, bracket
        prev = (ef, r)
    return efs[-1], prev[1], [prev, None]


def workload_desc(name, shape):
    """One description of the workload for both arms (the driver compares their `config`)."""
    graph = "GD graph M=30 (kNN-1000 + hnswlikeGD)" if shape.get("graph", "gd") == "gd" else "fixed-degree graph: cutKNNbyK(32) of the kNN lists (self included, as the reference keeps it)"
    return (f"{name}: {shape['n']}x{shape['d']} base, {shape['n_q']} queries, net {shape['d']}-{shape['d_hidden']}-"
            f"{shape['d_hidden']}-{shape['d_low']}, {graph}, projection + beam search + top-1 re-rank")


def host_threads():
    """Host threads the CPU arm may use: the cores this process may run on.  (Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which is what made round 1's N>1 reference numbers single-threaded.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------------------- reference arm
def reference_workload(args):
    from gbnns_dim_red_b200 import workload

    return workload.build_workload(args.workload, device=0, cache_dir=args.cache, log=log,
                                   n=args.n or None, n_q=args.n_q or None)


def run_reference(args, w=None, quiet=False):
    """The reference's own search code on the host CPU (all threads).  Returns the JSON dict."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests import _oracle as O
    from gbnns_dim_red_b200 import workload

    if O.ref("fast") is None:
        return {"impl": "reference", "unavailable": "oracle/_ref/libgbdr_ref_fast.so not built (needs /root/reference at build time)"}
    if w is None:
        w = reference_workload(args)
    shape = w["shape"]
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)  # performTest sets the team size itself (search_function.h:147); belt and braces
    goff, gedges = w["graph"]
    # low-dim queries with the reference's own GetLowQueryFromNet (untimed, as performTest expects them precomputed)
    q_low = O.ref_project(*w["net"], w["queries"], kind="fast")
    ctx = O.RefContext(w["base"], w["queries"], w["db_low"], q_low, w["truth"], goff, gedges, kind="fast")
    sample = min(shape["n_q"], args.ref_sample)

    def recall_of(ef):  # every query of the workload, scored by the reference's own loop (search_function.h:190-203)
        return ctx.perform_test(ef, w["entry"], n_q_use=shape["n_q"], number_exper=1, threads=threads)["acc"]

    if args.ef:
        ef, rec = args.ef, recall_of(args.ef)
        bracket = None
    else:
        ef, rec, bracket = pick_ef(recall_of)
    # size the sample so that one step is ~1-2 s of wall time
    probe = ctx.perform_test(ef, w["entry"], n_q_use=min(sample, 1000), number_exper=1, threads=threads)
    per_q = probe["work_time"]
    sample = int(max(500, min(shape["n_q"], 1.5 / max(per_q, 1e-9))))
    for _ in range(args.warmup):
        ctx.perform_test(ef, w["entry"], n_q_use=sample, number_exper=1, threads=threads)
    t_total = 0.0
    stats = None
    for _ in range(args.steps):
        stats = ctx.perform_test(ef, w["entry"], n_q_use=sample, number_exper=1, threads=threads)
        t_total += stats["work_time"] * sample  # StopW region of performTest (search_function.h:151,188)
    qps = sample * args.steps / t_total
    one_thread = ctx.perform_test(ef, w["entry"], n_q_use=min(sample, 1000), number_exper=1, threads=1)
    # recall@10 of the reference pipeline: exact re-rank of its own ef survivors, top 10 by (dist, id)
    rec10 = None
    try:
        from tests._data import exact_rerank_topk

        nq10 = min(shape["n_q"], 2000)
        r = O.ref_search(w["queries"][:nq10], q_low[:nq10], w["base"], w["db_low"], goff, gedges, ef, 1, 0, w["entry"][:nq10],
                         kind="fast", threads=threads)
        rec10 = workload.recall_at_k(exact_rerank_topk(r["low_ids"], w["queries"][:nq10], w["base"], 10), w["truth"][:nq10], 10)
    except Exception as e:  # a statistic, never a reason to lose the timing
        log(f"reference recall@10 unavailable: {e}")
    ctx.close()
    out = {
        "impl": "reference", "metric": metric_name(args.workload), "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, shape), "ef": ef, "recall_at_1": rec, "recall_at_10": rec10,
                   "ef_bracket": bracket, "queries_per_step": sample,
                   "setup": "dataset/graph built untimed by the GPU pipeline; timed region = reference performTest "
                            "(search_function.h:128-210) with precomputed low-dim queries, OpenMP over queries",
                   "build_flags": "README.md:33 flags (-Ofast -fopenmp -ftree-vectorize) with -march=x86-64-v3 in place of "
                                  "-march=native, so that the object built in the build container runs on this host "
                                  "(AVX2 + FMA; no AVX-512 code paths)"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{sample} of {shape['n_q']} queries per step, {args.steps} steps, ef={ef}; "
                                   f"1-thread QPS {1.0 / one_thread['work_time']:.0f}",
                         "recall_at_1": stats["acc"], "hops": stats["hops"], "dist_calc": stats["dist_calc"]},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    return out


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch

    from gbnns_dim_red_b200 import capi, workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"note: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE")
    n_gpus = world
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a host-only group: ranks that have nothing to do during rank 0's single-process legs must not sit in an NCCL
    # barrier (its kernel spins on their GPUs, which rank 0 is then driving through the C-ABI group)
    side = dist.new_group(backend="gloo") if dist is not None else None

    # ---- workload: rank 0 builds (or loads) and caches, the others load the cache ----
    kw = dict(cache_dir=args.cache, log=log, n=args.n or None, n_q=args.n_q or None)
    if world > 1:
        if rank == 0:
            w = workload.build_workload(args.workload, device=local, **kw)
        dist.barrier()
        if rank != 0:
            w = workload.build_workload(args.workload, device=local, **kw)
    else:
        w = workload.build_workload(args.workload, device=local, **kw)
    shape = w["shape"]
    n_q, d, d_low = shape["n_q"], shape["d"], shape["d_low"]
    goff, gedges = w["graph"]

    ix = capi.Index(local)
    ix.set_base(w["base"])
    ix.set_low(w["db_low"])
    ix.set_graph(goff, gedges)
    ix.set_net(*w["net"])
    if args.proj_mode is not None:
        ix.set_projection_mode(args.proj_mode)

    # ---- operating point: smallest ef with recall@1 >= 0.95 (projection on the fly, as performNetTest) ----
    def recall_of(ef):
        r = ix.search(w["queries"], None, ef, 1, w["entry"], flags=capi.SEARCH_RERANK)
        return workload.recall_at_1(r["ids"], w["truth"], w["base"])

    if args.ef:
        ef, rec, bracket = args.ef, recall_of(args.ef), None
    else:
        ef, rec, bracket = pick_ef(recall_of)
    r10 = ix.search(w["queries"], None, max(ef, 10), 10, w["entry"], flags=capi.SEARCH_RERANK)
    rec10 = workload.recall_at_k(r10["ids"], w["truth"], 10)
    log(f"operating point: ef={ef} recall@1={rec:.4f} recall@10={rec10:.4f}")

    # ---- device-resident leg (`value`) ----
    dev = torch.device("cuda", local)
    d_q = torch.from_numpy(w["queries"]).to(dev)
    d_entry = torch.from_numpy(w["entry"].astype(np.int32)).to(dev)
    d_ids = torch.empty((n_q, 1), dtype=torch.int32, device=dev)
    d_dists = torch.empty((n_q, 1), dtype=torch.float32, device=dev)
    d_hops = torch.empty(n_q, dtype=torch.int32, device=dev)
    d_dc = torch.empty(n_q, dtype=torch.int32, device=dev)
    d_sc = torch.empty(n_q, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_dev():
        ix.search_dev(d_q.data_ptr(), 0, n_q, ef, 1, d_entry.data_ptr(), d_ids.data_ptr(), d_dists.data_ptr(),
                      d_hops.data_ptr(), d_dc.data_ptr(), d_sc.data_ptr(), flags=capi.SEARCH_RERANK, stream=stream)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    ms_single = e0.elapsed_time(e1)
    kms = ix.last_kernel_ms(min(args.steps, 256))
    assert ix.status() & 6 == 0, "search reported a capacity failure"
    ids_dev = d_ids.cpu().numpy().astype(np.uint32).reshape(-1)
    rec_dev = workload.recall_at_1(ids_dev, w["truth"], w["base"])
    dc = d_dc.cpu().numpy().astype(np.int64) - ef  # low-dim evaluations (dist_calc minus the +ef of the re-rank)
    sc = d_sc.cpu().numpy().astype(np.int64)
    hops_mean = float(d_hops.float().mean().item())
    hops_np = d_hops.cpu().numpy()
    hops_pct = {f"p{q}": float(np.percentile(hops_np, q)) for q in (1, 50, 90, 99)}
    hops_pct["max"] = float(hops_np.max())
    log(f"hops per query: mean {hops_mean:.1f} " + " ".join(f"{k} {v:.0f}" for k, v in hops_pct.items()))

    # ---- device-resident leg with several batches in flight (`value` when --in-flight > 1) ----
    # batches outstanding per GPU: enough of them to fill the ~5000 walk slots of a B200 (148 SMs x 34 warps): four
    # 10 000-query batches, six 1000-query ones (GIST's n_q); `--in-flight N` fixes it
    nfl = args.in_flight if args.in_flight > 0 else (4 if n_q >= 5000 else min(8, -(-5032 // max(n_q, 1)) + 1))
    handles = [ix] + [ix.view() for _ in range(nfl - 1)]
    hstreams = [torch.cuda.ExternalStream(h.stream(), device=dev) for h in handles]
    obufs = [dict(ids=torch.empty((n_q, 1), dtype=torch.int32, device=dev),
                  dists=torch.empty((n_q, 1), dtype=torch.float32, device=dev),
                  hops=torch.empty(n_q, dtype=torch.int32, device=dev), dc=torch.empty(n_q, dtype=torch.int32, device=dev))
             for _ in handles]

    def step_flight(i):
        j = i % nfl
        o = obufs[j]
        handles[j].search_dev(d_q.data_ptr(), 0, n_q, ef, 1, d_entry.data_ptr(), o["ids"].data_ptr(),
                              o["dists"].data_ptr(), o["hops"].data_ptr(), o["dc"].data_ptr(), 0,
                              flags=capi.SEARCH_RERANK, stream=hstreams[j].cuda_stream)

    launches = 0
    ms_total = ms_single
    value_reps = [ms_single]
    if nfl > 1:
        for i in range(max(3, args.warmup) * nfl):
            step_flight(i)
        barrier()
        value_reps = []
        for rep in range(VALUE_REPS):  # the K-step loop is ~10 ms: repeat it and report the median (as `e2e` does)
            launches0 = capi.launch_count()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            f0.record(hstreams[0])
            for sst in hstreams[1:]:
                sst.wait_event(f0)
            for i in range(args.steps):
                step_flight(i)
            for sst in hstreams[1:]:
                fe = torch.cuda.Event()
                fe.record(sst)
                hstreams[0].wait_event(fe)
            f1.record(hstreams[0])
            barrier()
            launches = capi.launch_count() - launches0
            value_reps.append(f0.elapsed_time(f1))
        for h in handles:
            assert h.status() & 6 == 0, "search reported a capacity failure"
        for o in obufs[: min(nfl, args.steps)]:
            got = o["ids"].cpu().numpy().astype(np.uint32).reshape(-1)
            assert np.array_equal(got, ids_dev), "batches in flight changed the results"
    else:
        launches0 = capi.launch_count()
        step_dev()
        barrier()
        launches = (capi.launch_count() - launches0) * args.steps

    # ---- end-to-end leg (`e2e`): host-facing calls, pinned host buffers, copies inside the timed region ----
    h_q = capi.pinned_empty((n_q, d), np.float32)
    h_q[:] = w["queries"]
    h_entry = capi.pinned_empty((n_q,), np.uint32)
    h_entry[:] = w["entry"]
    outs = [dict(ids=capi.pinned_empty((n_q, 1), np.uint32), dists=capi.pinned_empty((n_q, 1), np.float32),
                 hops=capi.pinned_empty((n_q,), np.int32), dist_calc=capi.pinned_empty((n_q,), np.int32))
            for _ in handles]
    out = outs[0]
    for _ in range(max(3, args.warmup)):
        ix.search(h_q, None, ef, 1, h_entry, flags=capi.SEARCH_RERANK, out=out)
    barrier()
    # the host-timed legs are short (K steps of well under a millisecond): each is repeated E2E_REPS times and the
    # median K-step time is reported, so that one scheduling hiccup of the host thread does not decide the number
    sync_times = []
    for _ in range(E2E_REPS):
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ix.search(h_q, None, ef, 1, h_entry, flags=capi.SEARCH_RERANK, out=out)
        torch.cuda.synchronize()
        sync_times.append(time.perf_counter() - t0)
    e2e_sync_s = float(np.median(sync_times))
    rec_e2e = workload.recall_at_1(out["ids"], w["truth"], w["base"])
    ids_sync = out["ids"].copy()
    e2e_s = e2e_sync_s
    if nfl > 1:
        def run_pipelined(steps):
            busy = [False] * nfl
            for i in range(steps):
                j = i % nfl
                if busy[j]:
                    handles[j].search_wait()
                handles[j].search_submit(h_q, None, ef, 1, h_entry, flags=capi.SEARCH_RERANK, out=outs[j])
                busy[j] = True
            for j in range(nfl):
                if busy[j]:
                    handles[j].search_wait()

        run_pipelined(max(3, args.warmup) * nfl)
        barrier()
        pipe_times = []
        for _ in range(E2E_REPS):
            t0 = time.perf_counter()
            run_pipelined(args.steps)
            torch.cuda.synchronize()
            pipe_times.append(time.perf_counter() - t0)
        e2e_s = float(np.median(pipe_times))
        for o in outs[: min(nfl, args.steps)]:
            assert np.array_equal(o["ids"], ids_sync), "pipelined host calls changed the results"
    clocks = sampler.stop() if rank == 0 else None

    # ---- ef sweep (BASELINE config: "ef sweep recall/QPS curve"), outside the timed legs: device-resident,
    # one batch at a time, 3 steps per point after one warm-up, recall@1 scored like performTest ----
    ef_curve = []
    if rank == 0 and not args.no_ef_curve:
        for e in ([int(x) for x in args.efs.split(",")] if args.efs else EFS):
            if e > 500:
                continue
            def one(e=e):
                ix.search_dev(d_q.data_ptr(), 0, n_q, e, 1, d_entry.data_ptr(), d_ids.data_ptr(), d_dists.data_ptr(),
                              d_hops.data_ptr(), d_dc.data_ptr(), d_sc.data_ptr(), flags=capi.SEARCH_RERANK, stream=stream)
            one()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(3):
                one()
            c1.record()
            torch.cuda.synchronize()
            r = workload.recall_at_1(d_ids.cpu().numpy().astype(np.uint32).reshape(-1), w["truth"], w["base"])
            ef_curve.append({"ef": e, "recall_at_1": round(r, 4), "qps": round(3 * n_q / (c0.elapsed_time(c1) * 1e-3)),
                             "dist_calc": round(float(d_dc.float().mean().item()), 1)})
        log("ef curve: " + ", ".join(f"{c['ef']}:{c['recall_at_1']:.3f}@{c['qps'] / 1e6:.2f}M" for c in ef_curve))
    h2d = n_q * d * 4 + n_q * 4
    d2h = n_q * (4 + 4 + 4 + 4) + 4

    # ---- max over ranks ----
    tt = torch.tensor([e2e_s * 1e3, ms_single, e2e_sync_s * 1e3] + list(value_reps), dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms, ms_single, e2e_sync_ms = tt.tolist()[:3]
    value_reps = tt.tolist()[3:]
    ms_total = float(np.median(value_reps))
    qps = n_gpus * n_q * args.steps / (ms_total * 1e-3)
    e2e_qps = n_gpus * n_q * args.steps / (e2e_ms * 1e-3)
    qps_single = n_gpus * n_q * args.steps / (ms_single * 1e-3)
    e2e_sync_qps = n_gpus * n_q * args.steps / (e2e_sync_ms * 1e-3)

    # ---- roofline of the dominant kernel (beam search), SURVEY §8d accounting ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    k_low = ef  # ids handed to the re-rank
    bytes_search = 4.0 * (dc.sum() * d_low + sc.sum() + n_q * (d_low + 1 + k_low))
    bytes_rerank = 4.0 * (n_q * ef * d + n_q * (d + ef + 2))
    achieved = bytes_search / (kms["search"] * 1e-3) / 1e9
    # measured DRAM bytes of the same kernel/launch shape from one `ncu --set full` capture (scripts/ncu_traffic.py)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if tj.get("workload") == args.workload and int(tj.get("ef", -1)) == int(ef) and not (args.n or args.n_q):
            traffic = float(tj["dram_bytes_per_launch"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "beam_search_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_launch": bytes_search, "kernel_ms": kms["search"],
                "measured_in": "single-stream leg (one batch at a time; overlapped launches would include SM waiting)",
                "other_kernels_ms": {"project": kms["project"], "rerank": kms["rerank"]},
                "rerank_achieved_gbs": bytes_rerank / max(kms["rerank"], 1e-9) / 1e6,
                "per_query": {"low_dim_evals": float(dc.mean()), "adjacency_ids": float(sc.mean()), "hops": hops_mean,
                              "hops_percentiles": hops_pct}}

    build = measure_build(capi, w, local) if rank == 0 else {}
    if dist is not None:
        dist.barrier()
    result = {
        "metric": metric_name(args.workload), "value": qps, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.workload, shape), "ef": ef, "recall_at_1": rec_dev, "recall_at_10": rec10,
                   "recall_at_1_e2e": rec_e2e, "ef_bracket": bracket, "queries_per_step_per_gpu": n_q,
                   "graph_avg_degree": round(gedges.size / shape["n"], 1),
                   "parallelism": f"replicated index, {n_q} queries per step on each of {n_gpus} GPU(s)",
                   "batches_in_flight": nfl,
                   "l2_policy": f"inputs ({(w['base'].nbytes + w['db_low'].nbytes + 4 * shape['n'] * 64) / 1e9:.1f} GB of "
                                "db/db_low/graph gathers) larger than the 126 MB L2; no flush",
                   "projection": {0: "3xTF32 tcgen05", 1: "TF32 tcgen05", 2: "fp32 CUDA cores"}.get(args.proj_mode, "default")},
        "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps,
                "api": "gbdr_search_submit/gbdr_search_wait" if nfl > 1 else "gbdr_search", "batches_in_flight": nfl,
                "timing": f"host clock around K = {args.steps} steps, median of {E2E_REPS} repetitions",
                "repetitions_qps_rank0": [round(n_q * args.steps / t) for t in (pipe_times if nfl > 1 else sync_times)],
                "sync": {"value": e2e_sync_qps, "ms_per_step": e2e_sync_ms / args.steps, "api": "gbdr_search"}},
        "single_stream": {"value": qps_single, "ms_per_step": ms_single / args.steps},
        "value_timing": {"what": f"CUDA events around K = {args.steps} steps, max over ranks per repetition, median of {len(value_reps)} repetitions",
                         "repetitions_qps": [round(n_gpus * n_q * args.steps / (t * 1e-3)) for t in value_reps]},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "build": build,
        "knn_graph_build_sec": build.get("knn_build_s"),
    }
    if ef_curve:
        result["ef_curve"] = {"note": "device-resident, one batch at a time, projection + search + top-1 re-rank",
                              "points": ef_curve}
    if clocks is not None:
        result["clocks"] = clocks

    # ---- the multi-GPU paths the north star names beyond replicated queries (N > 1) ----
    if n_gpus > 1 and not args.no_multi_blocks:
        result["strong"] = strong_block(args, torch, dist, capi, ix, w, ef, rank, n_gpus, dev)
        for h in handles[1:]:
            h.close()
        ix.close()
        ix = None
        del d_q, d_entry, d_ids, d_dists
        torch.cuda.empty_cache()
        # (a leg that fails must not take the headline down with it; a failure inside a collective still stalls the
        #  other ranks until NCCL's watchdog fires, so the legs validate their inputs before the first collective)
        for name, leg in (("build_sharded", lambda: build_sharded_block(args, torch, dist, rank, n_gpus, local, w)),
                          ("sharded", lambda: sharded_block(args, dist, rank, n_gpus, local))):
            try:
                result[name] = leg()
            except Exception as e:
                result[name] = {"failed": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:  # one process, all the GPUs, through the C-ABI group; the other ranks wait on the host
            try:
                result["group"] = group_block(args, n_gpus, w, ef)
            except Exception as e:
                result["group"] = {"failed": str(e)}
        dist.barrier(group=side)

    # ---- CPU baseline beside it (rank 0) ----
    if rank == 0 and not args.no_cpu_baseline:
        try:
            ra = argparse.Namespace(**vars(args))
            ra.ef = ef
            ra.steps, ra.warmup = 5, 2
            ref = run_reference(ra, w=w, quiet=True)
            if "cpu_baseline" in ref:
                result["cpu_baseline"] = ref["cpu_baseline"]
            else:
                result["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                          "sample": ref.get("unavailable", "unavailable")}
        except Exception as e:  # the baseline must never take the GPU number down with it
            result["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                      "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(result), flush=True)
    if dist is not None:
        dist.barrier(group=side)  # (host-only wait: the CPU baseline above must not compete with spinning NCCL waits)
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- graph build seconds
def measure_build(capi, w, device):
    """The graph build of this workload, timed in THIS process (the workload itself may come from the cache, built by
    another process): after a warm-up call on a 65 536-row subset (kernel images, memory pools), one gbdr_build_graph of
    the whole low-dimensional base — kNN-1000 self-join + hnswlikeGD(M = 30) without leaving HBM, graph downloaded, kNN
    lists not — or, for the fixed-degree workloads, gbdr_knn (k = 33) + gbdr_knn_cut (32).  Seconds per stage as the
    library reports them (host clock around each stage, stream synchronised)."""
    Y = w["db_low"]
    t = dict(w["timings"])
    out = {"ground_truth_s": t.get("ground_truth_s"), "project_base_s": t.get("project_base_s")}
    try:
        if w["shape"].get("graph", "gd") == "gd":
            kk = min(1000, Y.shape[0])
            capi.build_graph(Y[: 1 << 16], knn_k=min(64, kk), M=30, device=device)
            t0 = time.perf_counter()
            off, ed, bt = capi.build_graph(Y, knn_k=kk, M=30, reverse=True, device=device)
            out.update({"knn_build_s": bt["knn_s"], "gd_prune_s": bt["prune_s"] + bt["finish_s"], "upload_s": bt["upload_s"],
                        "forward_prune_s": bt["prune_s"], "reverse_pass_and_output_s": bt["finish_s"],
                        "wall_s": time.perf_counter() - t0, "knn_k": kk, "M": 30,
                        "graph_identical_to_the_workload_graph": bool(np.array_equal(off, w["graph"][0]) and
                                                                      np.array_equal(ed, w["graph"][1])),
                        "what": "gbdr_build_graph (kNN self-join + hnswlikeGD from HBM), this process, after a warm-up call"})
        else:
            from gbnns_dim_red_b200 import xvecs

            capi.knn(Y[: 1 << 16], Y[: 1 << 16], 33, device=device)
            t0 = time.perf_counter()
            ids, knn_s = capi.knn(Y, Y, 33, device=device)
            koff, ked = xvecs.adjacency_from_matrix(ids)
            off, ed, cut_s = capi.knn_cut(koff, ked, Y, 32, device=device)
            out.update({"knn_build_s": knn_s, "gd_prune_s": cut_s, "wall_s": time.perf_counter() - t0, "knn_k": 33,
                        "graph_identical_to_the_workload_graph": bool(np.array_equal(ed, w["graph"][1])),
                        "what": "gbdr_knn (k = 33, CUDA events) + gbdr_knn_cut(32), this process, after a warm-up call"})
    except Exception as e:  # never take the headline down
        out["failed"] = f"{type(e).__name__}: {e}"
        out["knn_build_s"] = t.get("knn_build_s")
    return out


# ----------------------------------------------------------------------------- N > 1: strong scaling
def strong_block(args, torch, dist, capi, ix, w, ef, rank, world, dev):
    """ONE 10 000-query batch cut into N contiguous slices (rank r answers slice r on its replica): the latency floor
    of a single batch, where the weak-scaling headline gives every GPU its own batch.  Device-resident (CUDA events,
    max over ranks) and through the blocking host call (pinned host buffers, host clock, max over ranks)."""
    from gbnns_dim_red_b200 import multigpu as mg

    n_q = w["shape"]["n_q"]
    b, e = mg.partition(n_q, world, rank)
    m = e - b
    q = torch.from_numpy(np.ascontiguousarray(w["queries"][b:e])).to(dev)
    en = torch.from_numpy(w["entry"][b:e].astype(np.int32)).to(dev)
    ids = torch.empty((m, 1), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ix.search_dev(q.data_ptr(), 0, m, ef, 1, en.data_ptr(), ids.data_ptr(), flags=capi.SEARCH_RERANK, stream=stream)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(5):
        step()
    reps = []
    for _ in range(VALUE_REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        reps.append(e0.elapsed_time(e1))
    h_q = capi.pinned_empty((m, w["shape"]["d"]), np.float32)
    h_q[:] = w["queries"][b:e]
    h_e = capi.pinned_empty((m,), np.uint32)
    h_e[:] = w["entry"][b:e]
    out = dict(ids=capi.pinned_empty((m, 1), np.uint32), dists=capi.pinned_empty((m, 1), np.float32),
               hops=capi.pinned_empty((m,), np.int32), dist_calc=capi.pinned_empty((m,), np.int32))
    for _ in range(3):
        ix.search(h_q, None, ef, 1, h_e, flags=capi.SEARCH_RERANK, out=out)
    host = []
    for _ in range(E2E_REPS):
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ix.search(h_q, None, ef, 1, h_e, flags=capi.SEARCH_RERANK, out=out)
        host.append((time.perf_counter() - t0) * 1e3)
    tt = torch.tensor(reps + host, dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    reps, host = tt.tolist()[: len(reps)], tt.tolist()[len(reps):]
    ms, hms = float(np.median(reps)) / args.steps, float(np.median(host)) / args.steps
    return {"what": f"one {n_q}-query batch per step, split over {world} GPUs (replicated index)", "scaling": "strong",
            "value": n_q / (ms * 1e-3), "ms_per_step": ms, "unit": UNIT,
            "e2e": {"value": n_q / (hms * 1e-3), "ms_per_step": hms, "api": "gbdr_search (blocking) on each rank's slice"}}


# ----------------------------------------------------------------------------- N > 1: sharded graph build
def build_sharded_block(args, torch, dist, rank, world, local, w):
    """kNN-1000 self-join of the workload's 1M x 32 low-dimensional base + hnswlikeGD(M = 30), row-block sharded
    (multigpu.sharded_build_graph: own block up over PCIe, NCCL all-gather of the vectors, per-block kNN and forward
    prune from HBM, NCCL all-gather of the forward lists, reverse pass on rank 0).  Seconds, max over ranks; the graph
    is compared with the workload's (built on one GPU)."""
    from gbnns_dim_red_b200 import multigpu as mg

    Y = w["db_low"]
    kk = min(1000, Y.shape[0])
    mg.sharded_build_graph(Y[: 1 << 16], min(64, kk), 30, local)   # warm-up: workspaces, NCCL channels
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    off, edges, t = mg.sharded_build_graph(Y, kk, 30, local)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev = torch.device("cuda", local)
    tt = torch.tensor([t["upload_allgather_s"], t["knn_s"], t["prune_s"], t["finish_s"], wall], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    up, knn_s, prune_s, finish_s, wall = tt.tolist()
    same = None
    if rank == 0 and w["shape"].get("graph", "gd") == "gd":
        same = bool(np.array_equal(off, w["graph"][0]) and np.array_equal(edges, w["graph"][1]))
    one = w["timings"]
    return {"rows": int(Y.shape[0]), "d_low": int(Y.shape[1]), "knn_k": int(kk), "M": 30,
            "knn_build_sharded_s": knn_s, "gd_prune_sharded_s": prune_s + finish_s,
            "upload_allgather_s": up, "forward_prune_s": prune_s, "reverse_pass_s": finish_s, "wall_s": wall,
            "one_gpu": {"knn_build_s": one.get("knn_build_s"), "gd_prune_s": one.get("gd_prune_gpu_s"),
                        "note": "as timed when the workload was built; the line's `build` block is this process's measurement"},
            "graph_identical_to_one_gpu_build": same,
            "note": "max over ranks; kNN / forward prune = device time of each rank's row block (CUDA events)"}


# ----------------------------------------------------------------------------- N > 1: C-ABI group (one process)
def group_block(args, world, w, ef):
    """The same modes driven by ONE process through gbdr_group_* (include/gbdr.h): what the drop-in final_test binary
    uses when GBDR_DEVICES lists several GPUs.  Host buffers (pinned), host clock around K blocking calls, median of
    E2E_REPS repetitions.  Runs on rank 0 while the other ranks wait on the host."""
    from gbnns_dim_red_b200 import capi, synth, xvecs

    out = {}
    n_q, d = w["shape"]["n_q"], w["shape"]["d"]
    devs = list(range(world))

    def timed(call, n_per_call):
        for _ in range(3):
            call()
        ts = []
        for _ in range(E2E_REPS):
            t0 = time.perf_counter()
            for _ in range(args.steps):
                call()
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        return {"value": n_per_call * args.steps / t, "ms_per_step": 1e3 * t / args.steps, "unit": UNIT}

    # replicated: weak (N batches per call) and strong (one batch per call)
    g = capi.Group(devs, capi.GROUP_REPLICATED)
    try:
        g.set_net(*w["net"])
        g.set_base(w["base"])
        g.set_low(w["db_low"])
        g.set_graph(*w["graph"])
        for label, reps in (("replicated_weak", world), ("replicated_strong", 1)):
            m = n_q * reps
            hq = capi.pinned_empty((m, d), np.float32)
            he = capi.pinned_empty((m,), np.uint32)
            for r in range(reps):
                hq[r * n_q:(r + 1) * n_q] = w["queries"]
                he[r * n_q:(r + 1) * n_q] = w["entry"]
            ob = dict(ids=capi.pinned_empty((m, 1), np.uint32), dists=capi.pinned_empty((m, 1), np.float32),
                      hops=capi.pinned_empty((m,), np.int32), dist_calc=capi.pinned_empty((m,), np.int32))
            out[label] = timed(lambda: g.search(hq, None, ef, 1, he, flags=capi.SEARCH_RERANK, out=ob), m)
            out[label]["queries_per_call"] = m
            out[label]["recall_at_1"] = workload_recall(ob["ids"][:n_q], w)
        # the sharded build through the group (gbdr_group_build_graph), both exchanges
        if w["shape"].get("graph", "gd") == "gd":
            Y = w["db_low"]
            g.build_graph(Y[: 1 << 16], knn_k=64, M=30)  # warm-up
            for ex, name in ((capi.EXCHANGE_PEER, "peer"), (capi.EXCHANGE_NCCL, "nccl")):
                try:
                    g.set_exchange(ex)
                    t0 = time.perf_counter()
                    off, ed, t = g.build_graph(Y, knn_k=min(1000, Y.shape[0]), M=30)
                    t["wall_s"] = time.perf_counter() - t0
                    t["graph_identical_to_one_gpu_build"] = bool(np.array_equal(off, w["graph"][0]) and np.array_equal(ed, w["graph"][1]))
                    out["build_" + name] = t
                except capi.GbdrError as e:
                    out["build_" + name] = {"failed": str(e)}
    finally:
        g.close()

    # sharded: Deep-shaped shards (96 -> 16 dims, kNN-32 graph per shard), fused peer-load merge vs NCCL all-gather + merge
    shard_n, k, seed = args.group_shard_n, 10, 1234
    dd, d_low, dh = 96, 16, 128
    net = synth.make_net(dd, dh, d_low, seed=seed)
    queries = synth.make_part(n_q, dd, 100_003, seed=seed)
    g = capi.Group(devs, capi.GROUP_SHARDED)
    try:
        g.set_net(*net)
        base = np.concatenate([synth.make_part(shard_n, dd, r, seed=seed) for r in range(world)])
        pj = capi.Index(0)
        pj.set_net(*net)
        low = np.concatenate([pj.project(base[i:i + (1 << 20)]) for i in range(0, base.shape[0], 1 << 20)])
        pj.close()
        g.set_base(base)
        g.set_low(low)
        entry = np.empty((world, n_q), np.uint32)
        for r in range(world):
            b, e = g.shard_rows(r)
            ids, _ = capi.knn(low[b:e], low[b:e], 33, device=r)
            g.set_shard_graph(r, *xvecs.adjacency_from_matrix(np.ascontiguousarray(ids[:, 1:])))
            entry[r] = np.random.default_rng([seed, 31 + r]).integers(0, e - b, size=n_q, dtype=np.uint32)
        truth, _ = capi.knn(queries, base, 1)
        hq = capi.pinned_empty((n_q, dd), np.float32)
        hq[:] = queries
        he = capi.pinned_empty((world, n_q), np.uint32)
        he[:] = entry
        ob = dict(ids=capi.pinned_empty((n_q, k), np.uint32), dists=capi.pinned_empty((n_q, k), np.float32),
                  hops=capi.pinned_empty((n_q,), np.int32), dist_calc=capi.pinned_empty((n_q,), np.int32))
        gef = args.ef or ef
        res = {}
        for ex, name in ((capi.EXCHANGE_PEER, "peer"), (capi.EXCHANGE_NCCL, "nccl")):
            try:
                g.set_exchange(ex)
                res[name] = timed(lambda: g.search(hq, None, max(gef, k), k, he, flags=capi.SEARCH_RERANK, out=ob), n_q)
                res[name]["recall_at_1"] = float((ob["ids"][:, 0] == truth[:, 0]).mean())
                res[name + "_ids"] = ob["ids"].copy()
            except capi.GbdrError as e:
                res[name] = {"failed": str(e)}
        if "peer_ids" in res and "nccl_ids" in res:
            res["exchanges_agree"] = bool(np.array_equal(res.pop("peer_ids"), res.pop("nccl_ids")))
        res.pop("peer_ids", None)
        res.pop("nccl_ids", None)
        res["what"] = (f"{shard_n} x {dd} rows per shard x {world} shards, {n_q} queries per call searched on every shard at ef "
                       f"{max(gef, k)}, top-{k} re-rank per shard, merged on device 0")
        out["sharded"] = res
    finally:
        g.close()
    return out


def workload_recall(ids, w):
    from gbnns_dim_red_b200 import workload

    return workload.recall_at_1(ids, w["truth"], w["base"])


# ----------------------------------------------------------------------------- sharded arm (BASELINE config 5)
def sharded_block(args, dist, rank, world, local):
    """Deep-shaped database too large for one GPU, SURVEY.md §8e / BASELINE config 5: rank r holds `--shard-n` rows (96-dim
    base, 16-dim projection, a fixed-degree kNN-32 graph built INSIDE the shard; 12.5 M rows per GPU x 8 GPUs = the
    Deep-100M shape), every rank searches every query on its shard (on-the-fly projection, beam `ef`, top-k re-rank in
    the original dimension, global ids), the per-shard top-k lists are all-gathered over NCCL and merged by (dist, id)
    on the GPU (K5).  A step = one 10 000-query batch through all of that.  Returns the result dict (every rank)."""
    import torch

    from gbnns_dim_red_b200 import capi, multigpu as mg, synth, xvecs

    dev = torch.device("cuda", local)
    d, d_low, dh, n_q, k, seed = 96, 16, 128, args.n_q or 10_000, 10, 1234
    shard_n = args.shard_n
    n_total = shard_n * world
    t0 = time.time()
    base = synth.make_part(shard_n, d, rank, seed=seed)
    queries = synth.make_part(n_q, d, 100_003, seed=seed)
    net = synth.make_net(d, dh, d_low, seed=seed)
    log(f"  shard generated at {time.time() - t0:.1f}s")
    ix = capi.Index(local)
    ix.set_net(*net)
    db_low = np.empty((shard_n, d_low), np.float32)
    for i in range(0, shard_n, 1 << 20):
        db_low[i:i + (1 << 20)] = ix.project(base[i:i + (1 << 20)])
    log(f"  shard projected at {time.time() - t0:.1f}s")
    knn_ids, knn_s = capi.knn(db_low, db_low, 33, device=local)
    log(f"  shard kNN-33 at {time.time() - t0:.1f}s ({knn_s:.2f}s on the GPU)")
    goff, gedges = xvecs.adjacency_from_matrix(np.ascontiguousarray(knn_ids[:, 1:]))
    del knn_ids
    ix.set_base(base)
    ix.set_low(db_low)
    ix.set_graph(goff, gedges)
    ix.set_id_offset(rank * shard_n)
    # ground truth: exact nearest neighbour over ALL shards = (dist, id)-minimum of the per-shard exact answers
    t_ids, t_d, _ = capi.knn(queries, base, 1, device=local, return_dists=True)
    g_ids = mg.all_gather_equal(torch.from_numpy(t_ids.astype(np.int64) + rank * shard_n).to(dev))  # [world, n_q, 1]
    g_d = mg.all_gather_equal(torch.from_numpy(t_d).to(dev))
    best = torch.argmin(g_d[:, :, 0], dim=0)
    truth = g_ids[:, :, 0].gather(0, best.unsqueeze(0))[0].cpu().numpy().astype(np.uint32)
    entry = np.random.default_rng([seed, 31 + rank]).integers(0, shard_n, size=n_q, dtype=np.uint32)
    build_s = time.time() - t0
    log(f"shard built in {build_s:.1f}s: {shard_n} rows/GPU x {world} GPUs, shard kNN-33 {knn_s:.2f}s")

    d_q = torch.from_numpy(queries).to(dev)
    d_entry = torch.from_numpy(entry.astype(np.int32)).to(dev)
    d_ids = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    d_dists = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    d_hops = torch.empty(n_q, dtype=torch.int32, device=dev)
    d_dc = torch.empty(n_q, dtype=torch.int32, device=dev)
    d_sc = torch.empty(n_q, dtype=torch.int32, device=dev)
    merge = mg.gpu_merge(local)
    stream = torch.cuda.current_stream().cuda_stream

    def step(ef):
        ix.search_dev(d_q.data_ptr(), 0, n_q, ef, k, d_entry.data_ptr(), d_ids.data_ptr(), d_dists.data_ptr(),
                      d_hops.data_ptr(), d_dc.data_ptr(), d_sc.data_ptr(), flags=capi.SEARCH_RERANK, stream=stream)
        return merge(mg.all_gather_equal(d_ids), mg.all_gather_equal(d_dists), k)

    def recall_of(ef):
        if ef < k:
            return 0.0
        ids, _ = step(ef)
        torch.cuda.synchronize()
        return float((ids[:, 0].cpu().numpy().view(np.uint32) == truth).mean())

    fixed_ef = args.ef if args.workload == "deep-sharded" else args.shard_ef
    if fixed_ef:
        ef, rec, bracket = fixed_ef, recall_of(fixed_ef), None
    else:
        ef, rec, bracket = pick_ef(recall_of, efs=[e for e in EFS if e >= k])
    log(f"operating point: ef={ef} recall@1={rec:.4f}")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(ef)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(ef)
    e1.record()
    barrier()
    launches = capi.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    kms = ix.last_kernel_ms(min(args.steps, 256))
    assert ix.status() & 6 == 0, "search reported a capacity failure"
    dc = d_dc.cpu().numpy().astype(np.int64) - ef
    sc = d_sc.cpu().numpy().astype(np.int64)

    # end to end: queries from pinned host memory every step, merged ids/dists back to pinned host memory
    h_q = torch.from_numpy(queries).pin_memory()
    h_entry = torch.from_numpy(entry.astype(np.int32)).pin_memory()
    h_ids = torch.empty((n_q, k), dtype=torch.int32).pin_memory()
    h_d = torch.empty((n_q, k), dtype=torch.float32).pin_memory()

    def step_e2e():
        d_q.copy_(h_q, non_blocking=True)
        d_entry.copy_(h_entry, non_blocking=True)
        ids, dd = step(ef)
        h_ids.copy_(ids, non_blocking=True)
        h_d.copy_(dd, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    rec_e2e = float((h_ids[:, 0].numpy().view(np.uint32) == truth).mean())

    tt = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = tt.tolist()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_search = 4.0 * (dc.sum() * d_low + sc.sum() + n_q * (d_low + 1 + ef))
    achieved = bytes_search / (kms["search"] * 1e-3) / 1e9
    result = {
        "metric": "qps_at_recall1_0.95_deep_sharded", "value": n_q * args.steps / (ms_total * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"deep-sharded: {shard_n} x {d} rows per GPU ({n_total} in all), {n_q} queries/step searched on "
                               f"every shard, net {d}-{dh}-{dh}-{d_low}, per-shard fixed kNN-32 graph, projection + beam "
                               f"search + top-{k} re-rank per shard, NCCL all-gather + (dist,id) merge",
                   "ef": ef, "recall_at_1": rec, "recall_at_1_e2e": rec_e2e, "ef_bracket": bracket,
                   "parallelism": f"rows sharded x{world}, every GPU answers every query",
                   "l2_policy": "per-shard gathers larger than the 126 MB L2; no flush"},
        "e2e": {"value": n_q * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_q * d * 4 + n_q * 4,
                "d2h_bytes_per_step": n_q * k * 8, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "beam_search_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "algorithmic_bytes_per_launch": bytes_search,
                     "kernel_ms": kms["search"], "other_kernels_ms": {"project": kms["project"], "rerank": kms["rerank"]},
                     "per_query": {"low_dim_evals": float(dc.mean()), "adjacency_ids": float(sc.mean())}},
        "cpu_baseline": {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                         "sample": "the reference has no sharded search; see the sift1m workload for the CPU arm"},
        "shard_knn33_build_s": knn_s,
    }
    if clocks is not None:
        result["clocks"] = clocks
    result["shard_build_s"] = build_s
    ix.close()
    return result


def run_sharded(args):
    """`--workload deep-sharded`: the sharded-index leg on its own (one JSON line)."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    result = sharded_block(args, dist, rank, world, local)
    if rank == 0:
        print(json.dumps(result), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sift1m")
    ap.add_argument("--n", type=int, default=0, help="override base size (debug)")
    ap.add_argument("--n-q", dest="n_q", type=int, default=0, help="override query count (debug)")
    ap.add_argument("--ef", type=int, default=0, help="fix ef instead of searching for recall@1 >= 0.95")
    ap.add_argument("--proj-mode", dest="proj_mode", type=int, default=None)
    ap.add_argument("--cache", default=os.environ.get("GBDR_BENCH_CACHE", "/tmp/gbdr_bench_cache"))
    ap.add_argument("--ref-sample", dest="ref_sample", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-blocks", dest="no_multi_blocks", action="store_true",
                    help="N > 1: skip the strong-scaling, sharded-index, sharded-build and C-ABI group legs")
    ap.add_argument("--group-shard-n", dest="group_shard_n", type=int, default=1_000_000, help="rows per shard of the C-ABI group's sharded leg")
    ap.add_argument("--no-ef-curve", dest="no_ef_curve", action="store_true")
    ap.add_argument("--efs", default="", help="comma list: ef points of the curve instead of the reference's list")
    ap.add_argument("--shard-n", dest="shard_n", type=int, default=12_500_000,
                    help="sharded index: rows per GPU (12.5 M x 8 GPUs = the Deep-100M shape of BASELINE.json)")
    ap.add_argument("--shard-ef", dest="shard_ef", type=int, default=0, help="sharded leg of an N > 1 run: fix its ef")
    ap.add_argument("--in-flight", dest="in_flight", type=int, default=0,
                    help="batches outstanding per GPU (1 = one stream, blocking host calls)")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        if args.workload == "deep-sharded":
            print(json.dumps({"impl": "reference", "unavailable": "the reference has no sharded search (SURVEY.md §8e)"}))
            return 0
        out = run_reference(args)
        print(json.dumps(out), flush=True)
        return 0
    if args.workload == "deep-sharded":
        run_sharded(args)
        return 0
    run_ours(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
