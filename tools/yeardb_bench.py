"""BASELINE config 2: six time-routed year databases (J2020..J2025), a 127-question batch, one DB per GPU.

    python tools/yeardb_bench.py [--rows-min 200000] [--rows-max 2000000] [--questions 127] [--top-n 30]

Every visible GPU gets the year-DBs i % G (in-process placement, searched concurrently from a thread pool,
one batched tensor-core search per DB for all the questions routed to it).  Reports, per routing mix
(selected_years None / [2024] / [2023, 2024, 2025]):
  batch   VectorRetriever.retrieve_batch — host strings in, record dicts out (embedding, H2D, search, D2H,
          aggregation and fusion in Python)
  loop    the reference's shape: one retrieve_by_company_name call per question
  device  the batched searches alone (CUDA events per GPU, max over GPUs)
and checks batch == loop.  Prints one JSON line and appends it to gpurun_out/yeardb_bench.json.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import retrieval, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows-min", type=int, default=200_000)
ap.add_argument("--rows-max", type=int, default=2_000_000)
ap.add_argument("--questions", type=int, default=127)
ap.add_argument("--top-n", type=int, default=30)
ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--lazy-records", action="store_true", help="chunk / page records generated on demand (r01 behaviour)")
ap.add_argument("--python-records", action="store_true", help="build the records with the Python loops (no C helper)")
ap.add_argument("--phases", action="store_true", help="also report where the host time of one batch goes")
args = ap.parse_args()

COMPANY = "金盘科技"
YEARS = [2020, 2021, 2022, 2023, 2024, 2025]
G = torch.cuda.device_count()
sizes = np.linspace(args.rows_min, args.rows_max, len(YEARS)).astype(np.int64)
reports = []
block = synth.ROW_BLOCK
for i, (year, n) in enumerate(zip(YEARS, sizes)):
    dev = torch.device("cuda", i % G)
    lo = i * ((args.rows_max + block - 1) // block + 1) * block     # disjoint seeded blocks per year
    rows = synth.dense_rows_torch(lo, lo + int(n), args.dim, dev)
    index = pkg.FlatIPIndex.from_device_tensor(rows)
    reports.append(synth.synthetic_year_report(f"J{year}", COMPANY, int(n), index, materialise=not args.lazy_records))
if args.python_records:
    retrieval._b2r_records = None
retriever = retrieval.VectorRetriever.from_reports(reports, qwen=retrieval.HashEmbedder(args.dim), parallel_workers=8)

questions = [f"{YEARS[i % 6]}年{COMPANY}的第{i}个财务指标是多少？" for i in range(args.questions)]
mixes = {"all_years": None, "single_year": [2024], "three_years": [2023, 2024, 2025]}
kw = dict(top_n=args.top_n, use_hyde=False, use_multi_query=False)


def strip(res):
    return [(r["source_sha1"], r["page"], r["text"], r["vector_similarity"]) for r in res["results"]]


out = {"config": "6 year-DBs J2020-J2025, one per GPU (i % G)", "n_gpus": G, "rows_per_db": [int(n) for n in sizes],
       "questions": args.questions, "top_n": args.top_n, "dim": args.dim,
       "chunk_records": "on demand" if args.lazy_records else "materialised lists",
       "record_helper": "C (_b2r_records)" if retrieval._b2r_records is not None else "python loops", "mixes": {}}
for name, years in mixes.items():
    n_dbs = len(years) if years else len(YEARS)
    rows_scanned = int(sum(int(n) for y, n in zip(YEARS, sizes) if years is None or y in years))
    retriever.retrieve_batch(COMPANY, questions, selected_years=years, **kw)  # warm-up (workspaces, row norms)
    t_batch, got = [], None
    for _ in range(args.reps):
        got = None   # the previous batch's records are released before the clock starts, not inside the call
        t0 = time.perf_counter()
        got = retriever.retrieve_batch(COMPANY, questions, selected_years=years, **kw)
        t_batch.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    want = [retriever.retrieve_by_company_name(COMPANY, q, selected_years=years, **kw) for q in questions]
    t_loop = time.perf_counter() - t0
    same = all(strip(a) == strip(b) for a, b in zip(got, want))
    # device-only: the batched searches, events per GPU
    emb = np.concatenate([retriever.qwen.embed(q).reshape(1, -1) for q in questions]).astype(np.float32)
    per_dev = {}
    for rep in reports:
        year = int(rep["name"][1:])
        if years is not None and year not in years:
            continue
        per_dev.setdefault(rep["vector_db"].device, []).append(rep["vector_db"])
    evs = []
    for d, idxs in per_dev.items():
        with torch.cuda.device(d):
            q = torch.from_numpy(emb).to(f"cuda:{d}")
            for idx in idxs:
                idx.search_dev(q, args.top_n)
            torch.cuda.synchronize(d)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for idx in idxs:
                idx.search_dev(q, args.top_n)
            e1.record()
            evs.append((d, e0, e1))
    dev_ms = 0.0
    for d, e0, e1 in evs:
        torch.cuda.synchronize(d)
        dev_ms = max(dev_ms, e0.elapsed_time(e1))
    phases = None
    if args.phases:   # wall time inside the retriever's own phases for one more batch (wrappers add ~1 us per call)
        spent = {}

        def timed(fn, key):
            def run(*a, **k):
                t = time.perf_counter()
                try:
                    return fn(*a, **k)
                finally:
                    spent[key] = spent.get(key, 0.0) + time.perf_counter() - t
            return run

        originals = {}
        for key in ("_prepare", "_embed_queries", "_search_reports", "_finalize", "_direct_records"):
            originals[key] = getattr(retriever, key)
            setattr(retriever, key, timed(originals[key], key))
        got = None
        t0 = time.perf_counter()
        got = retriever.retrieve_batch(COMPANY, questions, selected_years=years, **kw)
        spent["total"] = time.perf_counter() - t0
        for key, fn in originals.items():
            setattr(retriever, key, fn)
        phases = {key: round(v * 1e3, 2) for key, v in spent.items()}
    best = min(t_batch)
    out["mixes"][name] = {
        "host_phases_ms": phases,
        "dbs_searched": n_dbs, "rows_scanned_per_question": rows_scanned,
        "batch_s": round(best, 4), "batch_questions_per_s": round(args.questions / best, 1),
        "loop_s": round(t_loop, 4), "loop_questions_per_s": round(args.questions / t_loop, 1),
        "device_search_ms_max_over_gpus": round(dev_ms, 3),
        "device_questions_per_s": round(args.questions / dev_ms * 1e3, 1),
        "device_GBps_equiv_all_gpus": round(rows_scanned * args.dim * 4 / dev_ms / 1e6, 1),
        "batch_equals_loop": bool(same),
    }
print(json.dumps(out, ensure_ascii=False), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/yeardb_bench.json", "a") as f:
    f.write(json.dumps(out, ensure_ascii=False) + "\n")
