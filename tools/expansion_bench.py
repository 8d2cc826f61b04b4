"""Next-row f3: graph expansions (ssg / triangulation / hybrid_expansion) — lockstep batched walks vs the
reference's one-walk-at-a-time shape, on a resident synthetic report.

    python tools/expansion_bench.py [ROWS] [QUESTIONS]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_challenge_2_b200 as pkg  # noqa: E402
from rag_challenge_2_b200 import retrieval, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rows = synth.dense_rows_torch(0, n, 1024, torch.device("cuda", 0))
report = synth.synthetic_year_report("J2024", "金盘科技", n, pkg.FlatIPIndex.from_device_tensor(rows))
r = retrieval.VectorRetriever.from_reports([report], qwen=retrieval.HashEmbedder(1024))
questions = [f"2024年的第{i}个指标" for i in range(nq)]
out = {"rows": n, "questions": nq, "methods": {}}
for method in ("ssg", "triangulation", "hybrid_expansion"):
    res = {}
    for mode, flag in (("lockstep", True), ("sequential", False)):
        r.batched_expansion = flag
        kw = dict(top_n=30, use_hyde=False, use_multi_query=False, retrieval_method=method, max_hops=4, neighbor_k=30)
        r.retrieve_by_company_name("金盘科技", questions[0], **kw)
        t0 = time.perf_counter()
        got = [r.retrieve_by_company_name("金盘科技", q, **kw) for q in questions]
        res[mode] = (time.perf_counter() - t0) / nq
        res[mode + "_records"] = [[(x["page"], x["vector_similarity"], sorted(x["retrieval_sources"])) for x in g["results"]] for g in got]
    same = res.pop("lockstep_records") == res.pop("sequential_records")
    out["methods"][method] = {"lockstep_ms_per_question": round(res["lockstep"] * 1e3, 2),
                              "sequential_ms_per_question": round(res["sequential"] * 1e3, 2),
                              "speedup": round(res["sequential"] / res["lockstep"], 2), "same_records": same}
print(json.dumps(out, ensure_ascii=False), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/expansion_bench.json", "a") as f:
    f.write(json.dumps(out, ensure_ascii=False) + "\n")
