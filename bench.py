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
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the CPU-baseline dense sample (0: the whole corpus if host RAM allows)")
    ap.add_argument("--cpu-docs", type=int, default=250_000, help="documents of the CPU-baseline BM25 sample inside the B200 arm's line")
    ap.add_argument("--ref-docs", type=int, default=1_000_000, help="documents of the BM25 sample of --impl reference (extrapolation x10 at 10M)")
    ap.add_argument("--bm25-weight-bits", type=int, default=32, choices=[32, 64],
                    help="stored BM25 weight width: 32 = 8 B/posting (scores within 1e-6 of rank_bm25), 64 = bit-exact")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (extras) after the headline")
    ap.add_argument("--no-overlap", action="store_true", help="run BM25 after the scan instead of beside it")
    ap.add_argument("--overlap", default="pdl", choices=["pdl", "streams"],
                    help="how BM25 runs beside the scan: programmatic dependent launch (one stream) or two streams")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="do not chain consecutive steps by programmatic dependent launch (each step then starts after the previous one's last kernel has finished)")
    ap.add_argument("--fake-bm25", action="store_true",
                    help="experiment (wrong results): BM25 launches nothing and returns one fixed list — what is left is scan + fusion")
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
# CPU baseline: the oracle (port of faiss / rank_bm25 semantics) timed on the host cores
# --------------------------------------------------------------------------------------------------
def host_mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1048576.0
    except OSError:
        pass
    return 0.0


def host_dense_rows(n, d, seed=1234, block=1 << 20):
    """n unit rows on the HOST without minutes of random-number generation: one random block of `block` rows, then
    column rotations of it (every row distinct, still unit norm).  A flat scan's cost does not depend on the values."""
    import numpy as np

    from rag_challenge_2_b200 import synth

    base = synth.dense_rows_np(min(block, n), d, seed=seed)
    if n <= block:
        return base
    out = np.empty((n, d), dtype=np.float32)
    for i, lo in enumerate(range(0, n, block)):
        hi = min(n, lo + block)
        out[lo:hi] = np.roll(base[: hi - lo], 37 * i, axis=1)
    return out


def cpu_baseline(args, n_queries, rows_host=None, bm25_docs=None, budget_s=60.0, verbose=False):
    """Times the CPU restatement of the reference path — faiss-style flat scan (numpy/OpenBLAS sgemv over ALL host
    cores + partial sort; also with ONE thread, which is what faiss itself uses below 20 queries) and rank_bm25's
    get_scores (one Python dict lookup per document and query token) + the reference's full Python sort — on the
    largest sample the host holds and the time budget allows, and scales each leg linearly to the corpus size (both are
    O(N) per query; the factors are reported, the dense leg runs at full size when 41 GB of host RAM are free)."""
    import numpy as np

    import oracle
    from oracle.bm25 import bm25_ref_from_arrays, token_corpus_arrays_np
    from rag_challenge_2_b200 import synth

    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # noqa: BLE001
        threadpool_limits = None

    N = args.rows
    n_docs = min(N, bm25_docs if bm25_docs is not None else args.cpu_docs)
    avail = host_mem_available_gb()
    if rows_host is None:
        bm_gb = 7.0 * n_docs / 1e6 + 4.0                     # the dict-of-dicts corpus + slack
        fit = int(max(0.0, avail - bm_gb) * 0.8 * 2**30 / (4.0 * args.dim))
        n_dense = min(N, args.cpu_rows if args.cpu_rows > 0 else N, max(fit, min(N, 1 << 20)))
        rows_host = host_dense_rows(n_dense, args.dim)
    n_dense = int(rows_host.shape[0])
    rng = np.random.default_rng(99)
    xq = rng.standard_normal((max(n_queries, 1), args.dim)).astype(np.float32)
    xq /= np.linalg.norm(xq, axis=1, keepdims=True)
    t_build0 = time.perf_counter()
    ranks, lens = token_corpus_arrays_np(n_docs, vocab=args.vocab, seed=4242)
    bm = bm25_ref_from_arrays(ranks, lens, args.vocab)
    del ranks, lens
    t_build = time.perf_counter() - t_build0
    queries = [[f"t{r}" for r in q] for q in synth.token_queries_np(max(n_queries, 1), vocab=args.vocab, seed=4243)]
    # The real third-party packages, if this machine has them (this image does not: SURVEY 8c): the reference arm
    # then times faiss / rank_bm25 themselves and reports how far the port's answers are from theirs.
    real, kind, deviation = {}, "port", None
    try:
        import faiss  # noqa: F401
        import rank_bm25  # noqa: F401

        fx = faiss.IndexFlatIP(args.dim)
        fx.add(rows_host)
        rb = rank_bm25.BM25Okapi.__new__(rank_bm25.BM25Okapi)
        rb.__dict__.update({k: getattr(bm, k) for k in ("k1", "b", "epsilon", "corpus_size", "doc_len", "doc_freqs", "idf",
                                                        "avgdl", "average_idf")})
        real = {"flat": fx, "bm25": rb}
        Df, If = fx.search(xq[:1], args.k)
        Do, Io = oracle.flat_ip_search_ref(rows_host, xq[:1], args.k)
        so, sr = bm.get_scores(queries[0]), rb.get_scores(queries[0])
        deviation = {"dense_ids_equal": bool(np.array_equal(If, Io)),
                     "dense_score_max_rel": float(np.max(np.abs(Df - Do) / np.maximum(np.abs(Do), 1e-30))),
                     "bm25_score_max_abs": float(np.max(np.abs(so - sr)))}
        kind = "reference"
    except Exception:  # noqa: BLE001 - not installed (or unusable): the port is the baseline
        real = {}
    dense = (lambda x: real["flat"].search(x, args.k)) if real else (lambda x: oracle.flat_ip_search_ref(rows_host, x, args.k))
    # "all the host threads it can use": torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently make the
    # all-cores leg a one-thread leg at N > 1 — the BLAS pool is set explicitly, and what it really has is reported
    cores = os.cpu_count()
    blas_threads = None
    if threadpool_limits is not None:
        try:
            from threadpoolctl import threadpool_info

            threadpool_limits(limits=cores)   # stays in force for the process (the one-thread leg below narrows it again)
            blas_threads = max([int(i.get("num_threads", 0)) for i in threadpool_info()] or [0]) or None
        except Exception:  # noqa: BLE001
            blas_threads = None
    dense(xq[:1])   # one untimed pass (BLAS thread pool, page faults)
    t_dense = t_bm = t_fuse = 0.0
    done = 0
    t_start = time.perf_counter()
    for i in range(max(n_queries, 1)):
        t0 = time.perf_counter()
        D, I = dense(xq[i: i + 1])
        t1 = time.perf_counter()
        sc = (real["bm25"] if real else bm).get_scores(queries[i])
        top = oracle.bm25_topn_ref(sc, args.k)                      # src/retrieval.py:264-265: full Python sort
        S = np.array([[sc[j] for j in top]])
        t2 = time.perf_counter()
        oracle.hybrid_fuse_ref(D, I, S, np.array([top]), args.k)
        t3 = time.perf_counter()
        t_dense += t1 - t0
        t_bm += t2 - t1
        t_fuse += t3 - t2
        done += 1
        if time.perf_counter() - t_start > budget_s and done >= 2:
            break
    t_dense, t_bm, t_fuse = t_dense / done, t_bm / done, t_fuse / done
    # one thread (faiss's own nq < 20 path is single-threaded per query): a 2M-row slice, scaled
    t_dense_1t, n_1t = None, min(n_dense, 2_000_000)
    if threadpool_limits is not None and not real:
        with threadpool_limits(limits=1):
            oracle.flat_ip_search_ref(rows_host[:n_1t], xq[:1], args.k)
            t0 = time.perf_counter()
            for i in range(2):
                oracle.flat_ip_search_ref(rows_host[:n_1t], xq[i % len(xq): i % len(xq) + 1], args.k)
            t_dense_1t = (time.perf_counter() - t0) / 2
    f_dense, f_bm = N / n_dense, N / n_docs
    dense_at_n, bm_at_n = t_dense * f_dense, t_bm * f_bm
    per_query = dense_at_n + bm_at_n + t_fuse
    return {
        "value": 1.0 / per_query, "unit": UNIT, "cores": blas_threads or cores, "kind": kind, "port_vs_packages": deviation,
        "threads": {"dense_scan_openblas": blas_threads or cores, "bm25_python_loops": 1, "host_cpus": cores},
        "sample": (("faiss-cpu IndexFlatIP.search + rank_bm25.BM25Okapi.get_scores (the packages themselves, installed on this "
                    "machine) + Python stable sort" if real else
                    "oracle port (numpy/OpenBLAS fp32 scan on all cores + partial sort; rank_bm25's per-document dict loops + "
                    "the reference's full Python sort)") +
                   f": dense MEASURED on {n_dense} x {args.dim} rows ({t_dense * 1e3:.1f} ms/query, x{f_dense:.2f} to {N}), "
                   f"BM25 MEASURED on {n_docs} documents ({t_bm * 1e3:.1f} ms/query, x{f_bm:.2f} to {N}); {done} queries"),
        "extrapolation": {"dense_factor": f_dense, "bm25_factor": f_bm, "law": "linear in the corpus size (flat scan; one dict lookup "
                          "per document and token; n log n sort treated as linear)"},
        "parts": {"dense_ms_per_query_at_N": dense_at_n * 1e3, "bm25_ms_per_query_at_N": bm_at_n * 1e3, "fusion_ms": t_fuse * 1e3,
                  "dense_1thread_ms_per_query_at_N": (None if t_dense_1t is None else t_dense_1t * (N / n_1t) * 1e3),
                  "dense_1thread_sample_rows": n_1t, "dense_GBps_all_cores": 4.0 * n_dense * args.dim / t_dense / 1e9,
                  "bm25_index_build_s": t_build},
        "measured_ms_per_query_on_sample": (t_dense + t_bm + t_fuse) * 1e3, "queries_timed": done,
        "host_mem_available_gb": avail,
    }


CB_KEYS = ("value", "unit", "cores", "kind", "sample", "threads", "port_vs_packages", "extrapolation", "parts",
           "measured_ms_per_query_on_sample", "queries_timed")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps K: up to K queries are timed, bounded so that the whole run ends within a few minutes
    cb = cpu_baseline(args, max(1, min(args.steps, 20)), bm25_docs=args.ref_docs, budget_s=90.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.nq / cb["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the config of the arm this line is compared with (same workload at N GPUs); the CPU arm itself does not shard
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_sharding": "none: one host process over the whole corpus",
        "cpu_baseline": {k: cb[k] for k in CB_KEYS},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "timed_steps": cb["queries_timed"],
        "ms_per_step_note": ("ms_per_step is the per-query time AT THE FULL CORPUS SIZE: the dense leg is measured at "
                             f"x{cb['extrapolation']['dense_factor']:.2f} and the BM25 leg at x{cb['extrapolation']['bm25_factor']:.2f} "
                             "of it (rank_bm25's list of per-document dicts for 10M documents needs ~70 GB and ~15 min to build); "
                             "the time actually spent per query on the sample is cpu_baseline.measured_ms_per_query_on_sample"),
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"hybrid top-{args.k}: flat IP over {args.rows} x {args.dim} fp32 + BM25Okapi over {args.rows} "
                        f"chunks (vocab {args.vocab}, Zipf) + fusion; {args.nq} query/step",
            "rows": args.rows, "dim": args.dim, "vocab": args.vocab, "k": args.k, "queries_per_step": args.nq,
            "sharding": f"rows/{world}", "bm25_overlapped_with_scan": (False if args.no_overlap else args.overlap), "l2": "inputs (rows + postings) far larger than the 126 MB L2; no flush needed"}


# --------------------------------------------------------------------------------------------------
# extras: the other BASELINE configs, each with its own roofline (explanatory; the metric is the headline)
# --------------------------------------------------------------------------------------------------
def cuda_time_ms(fn, iters, warm=1):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def measured_tensor_peaks(dev):
    """Denominators for the tensor-core extras: tf32 measured here (cuBLAS fp32 matmul with TF32 allowed, 8192^3, best of
    5 — library call as a yardstick only), bf16 from MEASURED_PEAKS.json (sustained) or the same probe."""
    import torch

    out = {}
    a = torch.randn((8192, 8192), device=dev)
    b = torch.randn((8192, 8192), device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        best = min(cuda_time_ms(lambda: torch.matmul(a, b), 3) for _ in range(5))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out["tf32_tflops"] = 2 * 8192.0**3 / best / 1e9
    out["tf32_source"] = "measured in this run: torch.matmul fp32 with TF32 allowed, 8192^3, best of 5"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        out["bf16_tflops"] = float(mp.get("bf16_tflops_sustained") or mp["bf16_tflops"])
        out["bf16_source"] = "MEASURED_PEAKS.json (sustained)"
    except Exception:  # noqa: BLE001
        a16, b16 = a.bfloat16(), b.bfloat16()
        best = min(cuda_time_ms(lambda: torch.matmul(a16, b16), 3) for _ in range(5))
        out["bf16_tflops"] = 2 * 8192.0**3 / best / 1e9
        out["bf16_source"] = "measured in this run: torch.matmul bf16 8192^3, best of 5"
    return out


def config_extras(args, flat, bm25, rows, df, dev, hbm_peak):
    import numpy as np
    import torch

    import rag_challenge_2_b200 as pkg
    from rag_challenge_2_b200 import cosine, retrieval, synth
    from rag_challenge_2_b200.bm25 import BM25Okapi

    ex = {}
    n, d = int(rows.shape[0]), args.dim

    def guarded(name, fn):
        try:
            ex[name] = fn()
        except Exception as e:  # noqa: BLE001 - an extra must never cost the headline line
            ex[name] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.synchronize()

    peaks = {}
    guarded("tensor_peaks", lambda: peaks.update(measured_tensor_peaks(dev)) or dict(peaks))

    # ---- config 3: dense batches of 1024 and 4096 queries on the tensor cores (tf32 candidates; bf16 copy of the rows) ----
    def dense_batch(nq, bf16):
        def run():
            q = synth.dense_queries_torch(nq, d, dev, rows, seed=500 + nq)
            flat.enable_bf16_candidates(bf16)
            try:
                ms = cuda_time_ms(lambda: flat.search_dev(q, args.k), 3 if nq <= 1024 else 2)
            finally:
                flat.enable_bf16_candidates(False)
            tf = 2.0 * n * d * nq / ms / 1e9
            peak = peaks.get("bf16_tflops" if bf16 else "tf32_tflops")
            return {"ms": ms, "rows": n, "queries_per_s": nq / ms * 1e3,
                    "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                                 "frac": (tf / peak if peak else None), "flops_per_launch_sequence": 2.0 * n * d * nq,
                                 "peak_source": peaks.get("bf16_source" if bf16 else "tf32_source"),
                                 "note": "whole search (pre-pass, candidate pass on tcgen05, select, fp32 rescoring, redo check); "
                                         "useful flops 2*n*d*nq"},
                    "hbm_GBps_equiv": (4.0 if not bf16 else 2.0) * n * d * -(-nq // 256) / ms / 1e6}
        return run

    for nq in (1024, 4096):
        guarded(f"config3_dense_batch{nq}_tf32", dense_batch(nq, False))
        guarded(f"config3_dense_batch{nq}_bf16_candidates", dense_batch(nq, True))

    # ---- config 4 alone: BM25 top-k, 1 and 64 queries per launch (bytes = sum of df * bytes per posting) ----
    def bm25_only(nq):
        def run():
            pool = synth.token_queries_np(64, vocab=args.vocab, seed=4243)
            _, wbits = bm25.n_postings()
            bpp = 4 + wbits // 8
            batches = []
            for i in range(0, 64, nq):
                toks = pool[i: i + nq]
                offs = np.zeros(len(toks) + 1, np.int32)
                np.cumsum([len(t) for t in toks], out=offs[1:])
                batches.append((torch.from_numpy(np.concatenate(toks).astype(np.int32)).to(dev), torch.from_numpy(offs).to(dev),
                                sum(int(df[t]) for q in toks for t in q)))

            def sweep():
                for t, o, _ in batches:
                    bm25.topk_dev(t, o, args.k)
            ms = cuda_time_ms(sweep, 3) / len(batches)
            nbytes = float(np.mean([b[2] for b in batches])) * bpp
            gbs = nbytes / ms / 1e6
            return {"ms": ms, "queries_per_s": nq / ms * 1e3,
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                 "algorithmic_bytes_per_launch": nbytes, "bytes_per_posting": bpp}}
        return run

    guarded("bm25_only_1_query", bm25_only(1))
    guarded("bm25_only_64_queries", bm25_only(64))

    # the same 64-query batch on an index with the library's DEFAULT float64 weights (12 B / posting, scores bit-identical
    # to rank_bm25; the headline corpus uses the 8-byte opt-in): built here from the same seeded postings, then freed
    def bm25_only_f64():
        nonlocal bm25
        if bm25.n_postings()[1] == 64:
            return {"note": "the bench index already stores float64 weights: see bm25_only_64_queries"}
        post = synth.token_postings_torch(0, args.rows, args.vocab, dev)
        idf64, _ = synth.bm25_idf_from_df(np.asarray(df), args.rows)
        keep = bm25
        bm25 = BM25Okapi.from_postings(post["indptr"], post["doc_ids"], post["tf"], post["doc_len"],
                                       torch.from_numpy(idf64).to(dev), float(post["total_len"]) / args.rows,
                                       device=dev.index or 0, weight_bits=64)
        del post
        try:
            out = bm25_only(64)()
        finally:
            bm25.close()
            bm25 = keep
            torch.cuda.empty_cache()
        return out

    guarded("bm25_only_64_queries_float64_weights", bm25_only_f64)

    # ---- config 2: six time-routed year databases, a 127-question batch, host strings in / record dicts out ----
    def config2():
        company, years = "\u91d1\u76d8\u79d1\u6280", [2020, 2021, 2022, 2023, 2024, 2025]
        sizes = np.linspace(200_000, 2_000_000, len(years)).astype(np.int64)
        reports = []
        for i, (year, m) in enumerate(zip(years, sizes)):
            lo = (700 + 130 * i) * synth.ROW_BLOCK                      # seeded blocks disjoint from the main corpus
            r = synth.dense_rows_torch(lo, lo + int(m), d, dev)
            reports.append(synth.synthetic_year_report(f"J{year}", company, int(m), pkg.FlatIPIndex.from_device_tensor(r),
                                                       materialise=True))   # real chunk / page lists, as loaded from JSON
        rt = retrieval.VectorRetriever.from_reports(reports, qwen=retrieval.HashEmbedder(d), parallel_workers=8)
        questions = [f"{years[i % 6]}\u5e74{company}\u7684\u7b2c{i}\u4e2a\u8d22\u52a1\u6307\u6807\u662f\u591a\u5c11\uff1f" for i in range(127)]
        kw = dict(top_n=args.k, use_hyde=False, use_multi_query=False)
        out = {"dbs": 6, "rows_per_db": [int(m) for m in sizes], "questions": 127, "n_gpus": 1,
               "chunk_records": "materialised lists (6.6M chunk dictionaries on the host)",
               "record_helper": "C (_b2r_records)" if retrieval._b2r_records is not None else "python loops"}
        for name, sel in (("all_years", None), ("three_years", [2023, 2024, 2025])):
            rt.retrieve_batch(company, questions, selected_years=sel, **kw)
            best, got = 1e30, None
            for _ in range(5):
                got = None   # the previous batch's records are released before the clock starts, not inside the call
                t0 = time.perf_counter()
                got = rt.retrieve_batch(company, questions, selected_years=sel, **kw)
                best = min(best, time.perf_counter() - t0)
            emb = torch.from_numpy(np.concatenate([rt.qwen.embed(q).reshape(1, -1) for q in questions]).astype(np.float32)).to(dev)
            idxs = [r["vector_db"] for r in reports if sel is None or int(r["name"][1:]) in sel]
            dev_ms = cuda_time_ms(lambda: [ix.search_dev(emb, args.k) for ix in idxs], 3)
            scanned = sum(int(ix.ntotal) for ix in idxs)
            # the retriever's own timing keys (reference :1081-1101): what the embedding stand-in took in the last batch —
            # a seeded numpy generator per question plus the list -> float32 conversion; a deployment waits for a service here
            emb_ms = 1e3 * sum(g["timing"]["embedding_generation"] for g in got)
            out[name] = {"e2e_questions_per_s": 127 / best, "e2e_ms": best * 1e3, "device_ms": dev_ms,
                         "host_over_device": best * 1e3 / dev_ms, "records_returned": sum(len(g["results"]) for g in got),
                         "records_built": sum(len(g["initial_retrieval_results"]) for g in got),
                         "embedding_stand_in_ms": emb_ms,
                         "roofline": {"bound": "hbm", "achieved": 4.0 * scanned * d / dev_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": 4.0 * scanned * d / dev_ms / 1e6 / hbm_peak,
                                      "algorithmic_bytes_per_launch_sequence": 4.0 * scanned * d,
                                      "note": "127 queries share every row read (one tensor-core batch per database)"}}
        for r in reports:
            r["vector_db"].close()
        return out

    guarded("config2_year_dbs_127_questions", config2)

    # ---- config 5: N x N cosine statistics, N = 200k x 1024 ----
    def config5():
        m = 200_000
        g = torch.Generator(device=dev).manual_seed(777)
        x = torch.randn((m, d), generator=g, device=dev) * (0.5 + 1.5 * torch.rand((m, 1), generator=g, device=dev))
        x = x.contiguous()
        cosine.cosine_stats(x[:8192].contiguous())
        times = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = cosine.cosine_stats(x)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        dt = min(times)
        useful = 2.0 * m * m * d / 2            # symmetry: each pair once
        issued = 3.0 * useful                   # 3xTF32: three tf32 MMAs per product
        peak = peaks.get("tf32_tflops")
        return {"n": m, "seconds": dt, "stats": st,
                "roofline": {"bound": "tensor", "achieved": issued / dt / 1e12, "peak": peak, "unit": "TFLOP/s",
                             "frac": (issued / dt / 1e12 / peak if peak else None), "useful_fp32_equiv_tflops": useful / dt / 1e12,
                             "flops_per_launch": issued, "peak_source": peaks.get("tf32_source"),
                             "note": "tf32 MMA flops issued (3 per fp32-accurate product, upper triangle only) / wall time of the call"}}

    guarded("config5_cosine_200k", config5)
    return ex


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
    df_host = df.cpu().numpy()
    idf, _ = synth.bm25_idf_from_df(df_host, args.rows)
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

    pdl = (not args.no_overlap) and args.overlap == "pdl"
    pipelined = pdl and not args.no_pipeline
    inner_dense = index.searcher.local_dense
    inner_bm25 = index.searcher.local_bm25
    if args.fake_bm25:
        S_fix, J_fix = inner_bm25(staged[0][1], staged[0][2], args.k)
        torch.cuda.synchronize()
        inner_bm25 = index.searcher.local_bm25 = (lambda t, o, k, **kw: (S_fix, J_fix))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: inputs resident; a stream of K steps.  With --overlap pdl the steps are chained by programmatic
    # dependent launch (the scan of step s + 1 starts while the fusion / exchange kernel of step s finishes), so no
    # event may sit between kernels here: the kernel timing below uses its own K steps.
    index.searcher.pipeline = pipelined
    for i in range(args.warmup):
        q, t, o, _ = staged[i % n_batches]
        index.search_hybrid_dev(q, t, o, args.k)
    sync_all()
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
    top1 = O[:, 0].tolist()
    if world > 1 and index.searcher.peer_exchange is not None:
        index.searcher.peer_exchange.check()

    # ---- the dominant kernel (the flat scan), K more steps with CUDA events around its launch on the launching
    # stream; steps NOT chained here (an event between two kernels serialises them)
    index.searcher.pipeline = False
    scan_events = []

    def timed_dense(q, k, **kw):
        # pdl: no stream operation may sit between the scan and the BM25 launch (it would serialise them), so
        # the closing event is recorded after BM25: the interval is the scan with BM25 running beside it
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = inner_dense(q, k, **kw)
        if not pdl:
            b.record()
        scan_events.append((a, b))
        return out

    timed_dense.last_launches = inner_dense.last_launches
    index.searcher.local_dense = timed_dense
    bm25_events = []

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
    for i in range(2):
        q, t, o, _ = staged[i % n_batches]
        index.search_hybrid_dev(q, t, o, args.k)
    sync_all()
    scan_events.clear()
    bm25_events.clear()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for i in range(args.steps):
        q, t, o, _ = staged[i % n_batches]
        index.search_hybrid_dev(q, t, o, args.k)
    u1.record()
    sync_all()
    ms_unpipelined = u0.elapsed_time(u1)
    scan_ms = sum(a.elapsed_time(b) for a, b in scan_events) / max(1, len(scan_events))
    bm25_ms = sum(a.elapsed_time(b) for a, b in bm25_events) / max(1, len(bm25_events))
    index.searcher.local_dense = inner_dense
    index.searcher.local_bm25 = inner_bm25

    # ---- e2e: host buffers in, host buffers out, through the public API ---------------------------------
    # Every step copies its inputs host -> device from pinned staging and delivers its result to host memory inside the
    # timed region.  Headline: the next question is SUBMITTED while the previous one is still on the GPU (two in flight:
    # search_hybrid_submit / _collect — what a caller answering questions from a pool does); also reported: one blocking
    # call at a time.
    host_batches = [(staged[i][0].cpu().numpy(), staged[i][3]) for i in range(n_batches)]
    can_submit = index.bm25 is not None and (world == 1 or index.searcher.peer_exchange is not None)
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
    ms_e2e_sync = g0.elapsed_time(g1)
    ms_e2e = ms_e2e_sync
    if can_submit:
        for _ in range(2):  # untimed: the two-in-flight pattern once through both staging slots
            a_t = index.search_hybrid_submit(host_batches[0][0], host_batches[0][1], args.k)
            b_t = index.search_hybrid_submit(host_batches[1 % n_batches][0], host_batches[1 % n_batches][1], args.k)
            index.search_hybrid_collect(a_t)
            index.search_hybrid_collect(b_t)
        sync_all()
        g0.record()
        pending = index.search_hybrid_submit(host_batches[0][0], host_batches[0][1], args.k)
        for i in range(1, args.steps):
            x, toks = host_batches[i % n_batches]
            nxt = index.search_hybrid_submit(x, toks, args.k)
            Fh, Oh = index.search_hybrid_collect(pending)
            pending = nxt
        Fh, Oh = index.search_hybrid_collect(pending)
        g1.record()
        sync_all()
        ms_e2e = g0.elapsed_time(g1)  # both events are recorded on an idle stream with nothing outstanding: wall time
    h2d, d2h = index.last_h2d_bytes, index.last_d2h_bytes
    sampler.stop()

    # ---- explanatory extras (not the metric): dense-only single query, and dense batches on the tensor-core
    # path (K2; 256 queries: the CTA-pair kernel); local shard only, CUDA events, a few iterations each
    extras = {}
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

    times = torch.tensor([ms_total, ms_e2e, scan_ms, ms_unpipelined, ms_e2e_sync], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, scan_ms, ms_unpipelined, ms_e2e_sync = times.tolist()

    if rank == 0:
        value = args.nq * args.steps / (ms_total / 1e3)
        # Both call patterns of the public API were timed over the same K steps (max over ranks each): one blocking call
        # per question, and two questions in flight (submit the next, collect the previous).  The line's e2e is the
        # pattern that served this configuration better — on several GPUs keeping two in flight hides the host's part
        # of a 0.74 ms step; on one GPU the two are within a percent and the blocking call is the steadier one.
        ms_e2e_two = ms_e2e
        two_better = can_submit and ms_e2e_two <= ms_e2e_sync
        ms_e2e = ms_e2e_two if two_better else ms_e2e_sync
        e2e_value = args.nq * args.steps / (ms_e2e / 1e3)
        peak, peak_src = measured_peak_gbs()
        alg_bytes = 4.0 * (hi - lo) * args.dim
        achieved = alg_bytes / (scan_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps,
                    "in_flight": (2 if two_better else 1),
                    "how": ("HybridIndex.search_hybrid_submit / _collect: numpy in, numpy out, the next question submitted while the "
                            "previous one is on the GPU; every step's H2D copy and result delivery inside the timed region"
                            if two_better else
                            "HybridIndex.search_hybrid: numpy in, numpy out, one blocking call per step; every step's H2D copy and "
                            "result delivery inside the timed region"),
                    "two_in_flight": ({"value": args.nq * args.steps / (ms_e2e_two / 1e3), "ms_per_step": ms_e2e_two / args.steps}
                                      if can_submit else None),
                    "one_call_at_a_time": {"value": args.nq * args.steps / (ms_e2e_sync / 1e3), "ms_per_step": ms_e2e_sync / args.steps}},
            "gpu_launches": int(launches), "collectives": int(collectives), "exchange": exchange,
            "roofline": {"bound": "hbm", "kernel": "flat_scan_tma_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic_bytes(hi - lo, args.dim),
                         "peak_source": peak_src,
                         "peak_note": "the measured peak is a copy (half reads, half writes); a read-only scan can exceed it — see frac_of_8TBps_nominal",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": scan_ms, "kernel_ms_note": (("scan with the BM25 kernel running beside it (events around the pair)"
                                                                    if pdl else "events around the scan launch") +
                                                                   f"; averaged over {args.steps} steps timed right after the metric's steps, "
                                                                   "un-chained (an event between kernels would serialise the chained steps)"),
                         "frac_of_8TBps_nominal": achieved / 8000.0,
                         "share_of_step": scan_ms * max(1, -(-args.nq // 4)) / (ms_unpipelined / args.steps),
                         "share_of_step_note": "kernel_ms / ms_per_step of the un-chained steps the events were taken in"},
            "steps_chained": ("programmatic dependent launch: the scan of step s+1 starts while the fusion / exchange kernel of step s "
                              "finishes" if pipelined else None),
            "ms_per_step_unchained": ms_unpipelined / args.steps,
            "clocks": sampler.summary(t_wall0, t_wall1),
            "bm25": {"postings_local": n_post, "bytes_per_posting": 4 + wbits // 8,
                     "kernel_ms_in_step": (None if pdl or not bm25_events else bm25_ms)},
            "top1_last_step": top1,
            "extras": extras,
        }
        if world == 1 and not args.no_extras:
            line["extras"].update(config_extras(args, flat, bm25, rows, df_host, dev, peak))
        if world == 1 and not args.no_cpu_baseline:
            # the dense leg at the FULL corpus size when the host can hold it (the rows come back from HBM, untimed)
            n_s = hi - lo if args.cpu_rows <= 0 else min(args.cpu_rows, hi - lo)
            fit = int(max(0.0, host_mem_available_gb() - 12.0) * 0.8 * 2**30 / (4.0 * args.dim))
            n_s = min(n_s, max(fit, min(hi - lo, 1 << 20)))
            rows_host = rows[:n_s].cpu().numpy()
            cb = cpu_baseline(args, 10, rows_host=rows_host, budget_s=30.0)
            del rows_host
            line["cpu_baseline"] = {k: cb[k] for k in CB_KEYS}
            ex = line["extras"]
            line["vs_cpu_by_leg"] = {   # each leg of the step against the same leg on the host cores (both at the full corpus size)
                "dense_all_cores": cb["parts"]["dense_ms_per_query_at_N"] / ex["dense_single_query"]["ms"],
                "dense_1_thread": (None if cb["parts"]["dense_1thread_ms_per_query_at_N"] is None
                                   else cb["parts"]["dense_1thread_ms_per_query_at_N"] / ex["dense_single_query"]["ms"]),
                "bm25": (cb["parts"]["bm25_ms_per_query_at_N"] / ex["bm25_only_1_query"]["ms"]
                         if "ms" in ex.get("bm25_only_1_query", {}) else None),
                "note": "CPU ms per query at N (see cpu_baseline.extrapolation) / GPU ms per query of the same leg, device-timed"}
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
