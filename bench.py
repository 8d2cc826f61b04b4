#!/usr/bin/env python
"""Benchmark of the retrieval hot path: exact hybrid top-30 (flat inner product over N x 1024 fp32
rows + BM25Okapi over the same N chunks + score fusion) in queries/s, on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the CPU restatement of faiss / rank_bm25, timed)

A step = one batch of `--nq` hybrid queries (default 1: the single-query latency path the north star
targets) over the whole corpus.  The corpus (default 10M chunks, fixed as GPUs are added: strong
scaling) is row-sharded over the ranks; every step is local scan + BM25 -> one all-gather -> merge ->
fusion, identical answers on all ranks.  Inputs are far larger than L2 (41 GB of rows per query), so
no flush is needed between steps.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "exact top-30 queries/s, 10M x 1024 IP + BM25 hybrid"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="chunks in the corpus (whole job)")
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--vocab", type=int, default=50_000)
    ap.add_argument("--nq", type=int, default=1, help="queries per step")
    ap.add_argument("--k", type=int, default=30)
    ap.add_argument("--pool", type=int, default=64, help="distinct queries cycled through")
    ap.add_argument("--cpu-rows", type=int, default=1_000_000, help="rows of the CPU-baseline dense sample")
    ap.add_argument("--cpu-docs", type=int, default=50_000, help="documents of the CPU-baseline BM25 sample")
    ap.add_argument("--bm25-weight-bits", type=int, default=32, choices=[32, 64],
                    help="stored BM25 weight width: 32 = 8 B/posting (scores within 1e-6 of rank_bm25), 64 = bit-exact")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run BM25 after the scan instead of beside it")
    ap.add_argument("--overlap", default="pdl", choices=["pdl", "streams"],
                    help="how BM25 runs beside the scan: programmatic dependent launch (one stream) or two streams")
    ap.add_argument("--no-scan-priority", action="store_true", help="scan on the caller's stream instead of a high-priority one (experiment)")
    ap.add_argument("--no-bm25-events", action="store_true", help="do not time the BM25 kernel inside the step (experiment)")
    ap.add_argument("--tunable", action="append", default=[], help="library tunable key=value (experiments)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="multi-GPU result exchange: one fused kernel over NVLink peer memory, or NCCL all-gather + merges")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []  # (t, sm_mhz, max_mhz, reasons)
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:  # noqa: BLE001
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            reasons = [n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")]
            self.samples.append((time.time(), sm, mx, reasons))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside and self.samples:  # region shorter than the sampling period: nearest samples
            inside = sorted(self.samples, key=lambda s: min(abs(s[0] - t0), abs(s[0] - t1)))[:2]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sms = sorted(s[1] for s in inside)
        reasons = sorted({r for s in inside for r in s[3]})
        return {"sm_mhz": sms[len(sms) // 2], "sm_max_mhz": inside[0][2], "reasons": reasons, "samples": len(inside)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes(rows_local, dim):
    """dram read+write bytes per scan launch from the committed ncu capture, if one exists for this shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get(f"flat_scan_{rows_local}x{dim}")
    except Exception:  # noqa: BLE001
        return None


# --------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (port of faiss / rank_bm25 semantics) on a bounded sample, scaled to N
# --------------------------------------------------------------------------------------------------
def cpu_baseline(args, steps, rows_sample=None, verbose=False):
    """Times the CPU restatement of the reference path on a bounded sample and extrapolates linearly
    in the corpus size (a flat scan and rank_bm25's per-document loops are both O(N))."""
    import numpy as np

    import oracle
    from rag_challenge_2_b200 import synth

    n_dense = min(args.cpu_rows, args.rows)
    n_docs = min(args.cpu_docs, args.rows)
    if rows_sample is None:
        rows_sample = synth.dense_rows_np(n_dense, args.dim, seed=1234)
    rng = np.random.default_rng(99)
    xq = rng.standard_normal((max(steps, 1), args.dim)).astype(np.float32)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    docs = synth.token_corpus_np(n_docs, vocab=args.vocab, seed=4242)
    bm = oracle.BM25OkapiRef(docs)
    queries = [[f"t{r}" for r in q] for q in synth.token_queries_np(max(steps, 1), vocab=args.vocab, seed=4243)]
    # The real third-party packages, if this machine has them (this image does not: SURVEY 8c): the reference arm
    # then times faiss / rank_bm25 themselves and reports how far the port's answers are from theirs.
    real, kind, deviation = {}, "port", None
    try:
        import faiss  # noqa: F401
        import rank_bm25  # noqa: F401

        fx = faiss.IndexFlatIP(args.dim)
        fx.add(rows_sample)
        real = {"flat": fx, "bm25": rank_bm25.BM25Okapi(docs)}
        Df, If = fx.search(xq[:1], args.k)
        Do, Io = oracle.flat_ip_search_ref(rows_sample, xq[:1], args.k)
        so, sr = bm.get_scores(queries[0]), real["bm25"].get_scores(queries[0])
        deviation = {"dense_ids_equal": bool(np.array_equal(If, Io)),
                     "dense_score_max_rel": float(np.max(np.abs(Df - Do) / np.maximum(np.abs(Do), 1e-30))),
                     "bm25_score_max_abs": float(np.max(np.abs(so - sr)))}
        kind = "reference"
    except Exception:  # noqa: BLE001 - not installed (or unusable): the port is the baseline
        real = {}
    # one untimed pass (BLAS thread pool, page faults)
    oracle.flat_ip_search_ref(rows_sample, xq[:1], args.k)
    t_dense = t_bm = 0.0
    for i in range(steps):
        t0 = time.perf_counter()
        if real:
            D, I = real["flat"].search(xq[i: i + 1], args.k)
        else:
            D, I = oracle.flat_ip_search_ref(rows_sample, xq[i: i + 1], args.k)
        t1 = time.perf_counter()
        sc = (real["bm25"] if real else bm).get_scores(queries[i])
        top = oracle.bm25_topn_ref(sc, args.k)
        S = np.array([[sc[j] for j in top]])
        t2 = time.perf_counter()
        oracle.hybrid_fuse_ref(D, I, S, np.array([top]), args.k)
        t_dense += t1 - t0
        t_bm += t2 - t1
    t_dense /= max(steps, 1)
    t_bm /= max(steps, 1)
    per_query = t_dense * (args.rows / n_dense) + t_bm * (args.rows / n_docs)
    return {
        "value": 1.0 / per_query, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "port_vs_packages": deviation,
        "threads": {"dense_scan_openblas": os.cpu_count(), "bm25_python_loops": 1},
        "sample": (("faiss-cpu IndexFlatIP.search + rank_bm25.BM25Okapi.get_scores (the packages themselves, installed on this "
                    "machine) + Python stable sort" if real else
                    "oracle port (numpy/OpenBLAS fp32 scan on all cores + partial sort; rank_bm25-style per-document dict loops + "
                    "Python stable sort)") + f": dense {n_dense} x {args.dim} rows {t_dense * 1e3:.1f} ms/query, BM25 {n_docs} docs "
                   f"{t_bm * 1e3:.1f} ms/query, both scaled linearly to {args.rows} chunks (extrapolated); {steps} queries"),
        "t_dense_ms_sample": t_dense * 1e3, "t_bm25_ms_sample": t_bm * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(max(0, min(args.warmup, 1))):
        pass
    steps = max(1, min(args.steps, 20))  # bounded: each step is one hybrid query on the CPU sample
    cb = cpu_baseline(args, steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.nq / cb["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "threads", "port_vs_packages")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "timed_steps": steps,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"hybrid top-{args.k}: flat IP over {args.rows} x {args.dim} fp32 + BM25Okapi over {args.rows} "
                        f"chunks (vocab {args.vocab}, Zipf) + fusion; {args.nq} query/step",
            "rows": args.rows, "dim": args.dim, "vocab": args.vocab, "k": args.k, "queries_per_step": args.nq,
            "sharding": f"rows/{world}", "bm25_overlapped_with_scan": (False if args.no_overlap else args.overlap), "l2": "inputs (rows + postings) far larger than the 126 MB L2; no flush needed"}


# --------------------------------------------------------------------------------------------------
# the B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rag_challenge_2_b200 as pkg
    from rag_challenge_2_b200 import hybrid, synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    for kv in args.tunable:
        key, v = kv.split("=")
        pkg._lib.tunable_set(key, int(v))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- build this rank's shard in HBM ------------------------------------------------------------
    bounds = hybrid.shard_bounds(args.rows, world, 128)   # balanced to 128 rows: the largest shard sets the pace
    lo, hi = bounds[rank], bounds[rank + 1]
    rows = synth.dense_rows_torch(lo, hi, args.dim, dev)
    flat = pkg.FlatIPIndex.from_device_tensor(rows)
    post = synth.token_postings_torch(lo, hi, args.vocab, dev)
    df = post["df"].clone()
    tot = torch.tensor([post["total_len"]], dtype=torch.int64, device=dev)
    if world > 1:  # corpus-global statistics, as an unsharded BM25Okapi would compute them
        dist.all_reduce(df)
        dist.all_reduce(tot)
    idf, _ = synth.bm25_idf_from_df(df.cpu().numpy(), args.rows)
    avgdl = float(tot.item()) / args.rows
    bm25 = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"],
                                   torch.from_numpy(idf).to(dev), avgdl, device=local_rank, weight_bits=args.bm25_weight_bits)
    n_post, wbits = bm25.n_postings()
    del post
    torch.cuda.empty_cache()
    exchange = args.exchange
    if world > 1 and exchange == "peer":
        try:
            index = hybrid.HybridIndex(flat, bm25, lo, overlap=(None if args.no_overlap else args.overlap), exchange="peer")
        except Exception as e:  # noqa: BLE001 - symmetric memory needs P2P / IPC support on the box
            if rank == 0:
                print(f"[bench] peer exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
            exchange = "nccl"
            index = hybrid.HybridIndex(flat, bm25, lo, overlap=(None if args.no_overlap else args.overlap))
    else:
        exchange = "nccl" if world > 1 else "none"
        index = hybrid.HybridIndex(flat, bm25, lo, overlap=(None if args.no_overlap else args.overlap))

    if args.no_scan_priority:
        index.searcher.scan_stream = None

    # ---- queries (rank 0 decides, everyone gets the same) --------------------------------------------
    pool = max(args.pool, args.nq)
    q_pool = synth.dense_queries_torch(pool, args.dim, dev, rows if rank == 0 else None)
    if world > 1:
        dist.broadcast(q_pool, 0)
    tok_pool = synth.token_queries_np(pool, vocab=args.vocab, seed=4243)

    def batch(i):
        idx = [(i * args.nq + j) % pool for j in range(args.nq)]
        toks = [tok_pool[j] for j in idx]
        offs = np.zeros(args.nq + 1, np.int32)
        np.cumsum([len(t) for t in toks], out=offs[1:])
        return idx, toks, offs

    n_batches = max(1, pool // args.nq)
    staged = []
    for i in range(n_batches):
        idx, toks, offs = batch(i)
        staged.append((q_pool[idx].contiguous(), torch.from_numpy(np.concatenate(toks).astype(np.int32)).to(dev),
                       torch.from_numpy(offs).to(dev), toks))

    # time the dominant kernel (the flat scan) with events around its launch, on the launching stream
    scan_events = []
    inner_dense = index.searcher.local_dense

    pdl = (not args.no_overlap) and args.overlap == "pdl"

    def timed_dense(q, k):
        # pdl: no stream operation may sit between the scan and the BM25 launch (it would serialise them), so
        # the closing event is recorded after BM25: the interval is the scan with BM25 running beside it
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = inner_dense(q, k)
        if not pdl:
            b.record()
        scan_events.append((a, b))
        return out

    timed_dense.last_launches = inner_dense.last_launches
    index.searcher.local_dense = timed_dense
    bm25_events = []
    inner_bm25 = index.searcher.local_bm25

    def timed_bm25(t, o, k, **kw):  # events on the stream BM25 is enqueued on
        if pdl:
            out = inner_bm25(t, o, k, **kw)
            scan_events[-1][1].record()
            return out
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = inner_bm25(t, o, k, **kw)
        b.record()
        bm25_events.append((a, b))
        return out

    if pdl or not args.no_bm25_events:
        index.searcher.local_bm25 = timed_bm25

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: inputs resident ------------------------------------------------------------------------
    for i in range(args.warmup):
        q, t, o, _ = staged[i % n_batches]
        index.search_hybrid_dev(q, t, o, args.k)
    sync_all()
    scan_events.clear()
    bm25_events.clear()
    launches = 0
    collectives = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        q, t, o, _ = staged[i % n_batches]
        F, O = index.search_hybrid_dev(q, t, o, args.k)
        launches += index.searcher.launches
        collectives += index.searcher.collectives
    e1.record()
    sync_all()
    t_wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    scan_ms = sum(a.elapsed_time(b) for a, b in scan_events) / max(1, len(scan_events))
    bm25_ms = sum(a.elapsed_time(b) for a, b in bm25_events) / max(1, len(bm25_events))
    top1 = O[:, 0].tolist()

    # ---- e2e: host buffers in, host buffers out, through the public API ---------------------------------
    host_batches = [(staged[i][0].cpu().numpy(), staged[i][3]) for i in range(n_batches)]
    for i in range(min(args.warmup, 3)):
        index.search_hybrid(host_batches[i % n_batches][0], host_batches[i % n_batches][1], args.k)
    sync_all()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(args.steps):
        x, toks = host_batches[i % n_batches]
        Fh, Oh = index.search_hybrid(x, toks, args.k)
    g1.record()
    sync_all()
    ms_e2e = g0.elapsed_time(g1)
    h2d, d2h = index.last_h2d_bytes, index.last_d2h_bytes
    sampler.stop()

    # ---- explanatory extras (not the metric): dense-only single query, and dense batches on the tensor-core
    # path (K2; 256 queries: the CTA-pair kernel); local shard only, CUDA events, a few iterations each
    extras = {}
    index.searcher.local_dense = inner_dense
    index.searcher.local_bm25 = inner_bm25
    for name, nqx, iters in (("dense_single_query", 1, 20), ("dense_batch64_tensor_core", 64, 5),
                             ("dense_batch128_tensor_core", 128, 5), ("dense_batch256_tensor_core_cta_pairs", 256, 5)):
        qx = q_pool[:nqx].contiguous() if pool >= nqx else q_pool.repeat((nqx + pool - 1) // pool, 1)[:nqx].contiguous()
        for _ in range(2):
            flat.search_dev(qx, args.k)
        torch.cuda.synchronize()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record()
        for _ in range(iters):
            flat.search_dev(qx, args.k)
        x1.record()
        torch.cuda.synchronize()
        msx = x0.elapsed_time(x1) / iters
        extras[name] = {"ms": msx, "local_rows": hi - lo, "queries_per_s_local_shard": nqx / msx * 1e3,
                        "GBps_equiv": 4.0 * (hi - lo) * args.dim / msx / 1e6}

    times = torch.tensor([ms_total, ms_e2e, scan_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, scan_ms = times.tolist()

    if rank == 0:
        value = args.nq * args.steps / (ms_total / 1e3)
        e2e_value = args.nq * args.steps / (ms_e2e / 1e3)
        peak, peak_src = measured_peak_gbs()
        alg_bytes = 4.0 * (hi - lo) * args.dim
        achieved = alg_bytes / (scan_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "collectives": int(collectives), "exchange": exchange,
            "roofline": {"bound": "hbm", "kernel": "flat_scan_tma_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(hi - lo, args.dim),
                         "peak_source": peak_src,
                         "peak_note": "the measured peak is a copy (half reads, half writes); a read-only scan can exceed it — see frac_of_8TBps_nominal",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": scan_ms, "kernel_ms_note": ("scan with the BM25 kernel running beside it (events around the pair)"
                                                                   if pdl else "events around the scan launch"),
                         "frac_of_8TBps_nominal": achieved / 8000.0,
                         "share_of_step": scan_ms * max(1, -(-args.nq // 4)) / (ms_total / args.steps)},
            "clocks": sampler.summary(t_wall0, t_wall1),
            "bm25": {"postings_local": n_post, "bytes_per_posting": 4 + wbits // 8,
                     "kernel_ms_in_step": (None if pdl or not bm25_events else bm25_ms)},
            "top1_last_step": top1,
            "extras": extras,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_s = min(args.cpu_rows, hi - lo)
            cb = cpu_baseline(args, 10, rows_sample=rows[:n_s].cpu().numpy())
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "threads", "port_vs_packages")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()
