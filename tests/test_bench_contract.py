"""CPU: the JSON line of bench.py's reference arm (the oracle port timed on the host cores) carries every key the
measurement contract names, and the synthetic year-database helpers behave like the lists the retriever expects."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--rows", "100000", "--cpu-rows", "20000", "--ref-docs", "10000"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("exact top-30 queries/s") and d["value"] > 0 and d["vs_baseline"] is None
    assert d["scaling"] == "strong" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "MEASURED on 10000 documents" in cb["sample"]
    # the extrapolation from the sample to the corpus is stated per leg, and the per-leg times add up to the value
    assert cb["extrapolation"]["dense_factor"] == 5.0 and cb["extrapolation"]["bm25_factor"] == 10.0
    parts = cb["parts"]
    total_ms = parts["dense_ms_per_query_at_N"] + parts["bm25_ms_per_query_at_N"] + parts["fusion_ms"]
    assert abs(total_ms - 1e3 / d["value"]) < 1e-6 * total_ms and parts["dense_1thread_ms_per_query_at_N"] > 0
    assert cb["measured_ms_per_query_on_sample"] < total_ms and cb["queries_timed"] == d["timed_steps"] == 2
    assert abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]


def test_reference_arm_prints_the_config_of_the_arm_it_is_compared_with():
    """At N GPUs the driver compares the two arms' `config` objects: the CPU arm names the same workload (sharding
    rows/N of the GPU arm), says separately that it does not shard itself, and uses every host CPU for its dense leg
    even under torchrun's OMP_NUM_THREADS=1."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="4", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "4", "--steps", "1",
                          "--warmup", "0", "--rows", "50000", "--cpu-rows", "10000", "--ref-docs", "5000"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["n_gpus"] == 4 and d["config"]["sharding"] == "rows/4" and d["cpu_sharding"].startswith("none")
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    ns = argparse.Namespace(k=d["config"]["k"], rows=d["config"]["rows"], dim=d["config"]["dim"], vocab=d["config"]["vocab"],
                            nq=d["config"]["queries_per_step"], no_overlap=False, overlap=d["config"]["bm25_overlapped_with_scan"])
    assert d["config"] == bench.workload_config(ns, 4)
    threads = d["cpu_baseline"]["threads"]
    assert threads["dense_scan_openblas"] == threads["host_cpus"] or threads["host_cpus"] == 1


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_synthetic_year_report_records():
    sys.path.insert(0, ROOT)
    from rag_challenge_2_b200 import synth

    rep = synth.synthetic_year_report("J2024", "ACME", 1003, vector_db=None, chunks_per_page=8)
    doc = rep["document"]
    chunks, pages = doc["content"]["chunks"], doc["content"]["pages"]
    assert len(chunks) == 1003 and len(pages) == 126 and doc["metainfo"]["pages_amount"] == 126
    assert chunks[0]["page"] == 1 and chunks[1002]["page"] == 126 and chunks[-1]["id"] == 1002
    assert chunks[17] == {"page": 3, "length_tokens": 300, "text": "J2024 chunk 17", "id": 17, "type": "content"}
    with pytest.raises(IndexError):
        chunks[1003]
    assert [p["page"] for p in pages][:3] == [1, 2, 3] and sum(1 for _ in chunks) == 1003
    key, lookup = rep["_page_lookup"]
    assert key == id(pages) and lookup[126]["text"] == "J2024 page 126" and lookup[126] is lookup[126]
    with pytest.raises(KeyError):
        lookup[127]
    import numpy as np

    assert 1 in lookup and 126 in lookup and np.int64(5) in lookup       # membership without materialising the record
    assert 0 not in lookup and 127 not in lookup and "3" not in lookup and 3.5 not in lookup
