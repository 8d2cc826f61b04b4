"""Deterministic miniature `databases/` tree (chunked_reports + vector_dbs + bm25_dbs) in the
reference's on-disk layout: six year-databases J2020..J2025 of 金盘科技 plus one
report of another company whose year comes from subset.csv.  Regenerated identically by
make_golden.py and by the tests; nothing but the golden outputs is committed."""
import json
import pickle
import struct
import sys
import types
from pathlib import Path

import numpy as np

from . import stubs

TOPICS = ["revenue", "gross", "margin", "cash", "flow", "dividend", "transformer", "export", "overseas", "capacity",
          "research", "development", "inventory", "receivable", "subsidy", "tax", "profit", "equity", "debt", "order",
          "backlog", "factory", "mexico", "poland", "digital", "energy", "storage", "wind", "solar", "grid"]
FILL = ["the", "company", "reported", "that", "during", "period", "increase", "decrease", "of", "percent", "million",
        "yuan", "compared", "with", "previous", "year", "mainly", "due", "to", "and", "in", "for", "total", "net"]
COMPANY = "金盘科技"
OTHER = "Tradition"
OTHER_SHA1 = "2779336b845a41544348abb7b3e6e5bd2ff893a2"


def _text(rng, year, n_words):
    words = []
    for _ in range(n_words):
        words.append(str(rng.choice(TOPICS)) if rng.random() < 0.35 else str(rng.choice(FILL)))
    return f"{year} " + " ".join(words)


def make_report(sha1_name, company, year_label, n_pages, seed, with_year_field=False):
    rng = np.random.default_rng(seed)
    pages, chunks = [], []
    cid = 0
    for p in range(1, n_pages + 1):
        n_chunks = int(rng.integers(2, 5))
        texts = [_text(rng, year_label, int(rng.integers(12, 40))) for _ in range(n_chunks)]
        pages.append({"page": p, "text": " ".join(texts)})
        for t in texts:
            chunks.append({"page": p, "length_tokens": len(t.split()), "text": t, "id": cid, "type": "content"})
            cid += 1
    # a couple of exact duplicate chunks: exercises equal scores / tie ordering
    for _ in range(2):
        src = chunks[int(rng.integers(0, len(chunks)))]
        chunks.append({"page": src["page"], "length_tokens": src["length_tokens"], "text": src["text"], "id": cid,
                       "type": "content"})
        cid += 1
    meta = {"sha1_name": sha1_name, "company_name": company, "pages_amount": n_pages, "text_blocks_amount": len(chunks),
            "tables_amount": 0, "pictures_amount": 0, "equations_amount": 0, "footnotes_amount": 0}
    if with_year_field:
        meta["year"] = str(year_label)
    return {"metainfo": meta, "content": {"pages": pages, "chunks": chunks}}


def write_ixfi(path, rows):
    rows = np.ascontiguousarray(rows, dtype="<f4")
    with open(path, "wb") as f:
        f.write(struct.pack("<4siqqqBi", b"IxFI", rows.shape[1], rows.shape[0], 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", rows.size))
        f.write(rows.tobytes())


def _pickle_as_rank_bm25(obj_state, path):
    """Write a pickle whose class reference is rank_bm25.BM25Okapi, as src/ingestion.py:46-47 does."""
    had = sys.modules.get("rank_bm25")
    fake = types.ModuleType("rank_bm25")

    class BM25Okapi(object):
        pass

    BM25Okapi.__module__ = "rank_bm25"
    BM25Okapi.__qualname__ = "BM25Okapi"
    fake.BM25Okapi = BM25Okapi
    sys.modules["rank_bm25"] = fake
    try:
        o = BM25Okapi()
        o.__dict__.update(obj_state)
        with open(path, "wb") as f:
            pickle.dump(o, f)
    finally:
        if had is not None:
            sys.modules["rank_bm25"] = had
        else:
            del sys.modules["rank_bm25"]


def build(root):
    """Create <root>/databases/{chunked_reports,vector_dbs,bm25_dbs} and <root>/subset.csv."""
    root = Path(root)
    docs_dir = root / "databases" / "chunked_reports"
    vec_dir = root / "databases" / "vector_dbs"
    bm_dir = root / "databases" / "bm25_dbs"
    for d in (docs_dir, vec_dir, bm_dir):
        d.mkdir(parents=True, exist_ok=True)
    reports = []
    for i, year in enumerate(range(2020, 2026)):
        reports.append(make_report(f"J{year}", COMPANY, year, n_pages=10 + 2 * i, seed=100 + i,
                                   with_year_field=(year == 2022)))
    reports.append(make_report(OTHER_SHA1, OTHER, 2021, n_pages=6, seed=200))
    for rep in reports:
        name = rep["metainfo"]["sha1_name"]
        with open(docs_dir / f"{name}.json", "w", encoding="utf-8") as f:
            json.dump(rep, f, ensure_ascii=False)
        texts = [c["text"] for c in rep["content"]["chunks"]]
        rows = np.stack([stubs.embed_text(t) for t in texts])
        write_ixfi(vec_dir / f"{name}.faiss", rows)
        bm = stubs.OracleBM25([t.split() for t in texts])
        state = {"corpus_size": bm.corpus_size, "avgdl": bm.avgdl, "doc_freqs": bm.doc_freqs, "idf": bm.idf,
                 "doc_len": bm.doc_len, "tokenizer": None, "k1": 1.5, "b": 0.75, "epsilon": 0.25,
                 "average_idf": bm.average_idf}
        if name != "J2021":  # one missing BM25 index: the retriever must warn and skip it (:249-251)
            _pickle_as_rank_bm25(state, bm_dir / f"{name}.pkl")
    # a JSON without an index: VectorRetriever skips it (:498-500), BM25Retriever finds no pickle (:249-251)
    with open(docs_dir / "orphan.json", "w") as f:
        json.dump(make_report("orphan", COMPANY, 2019, 2, 300), f)
    # second documents dir = the same reports plus files VectorRetriever._load_dbs must skip (invalid
    # JSON, wrong schema); the reference's BM25Retriever has no such guard, so it gets the clean dir
    dirty_dir = root / "databases" / "chunked_reports_dirty"
    dirty_dir.mkdir(parents=True, exist_ok=True)
    for src in docs_dir.glob("*.json"):
        (dirty_dir / src.name).write_bytes(src.read_bytes())
    with open(dirty_dir / "broken.json", "w") as f:
        f.write("{ not json")
    write_ixfi(vec_dir / "broken.faiss", np.zeros((1, stubs.DIM), np.float32))
    with open(dirty_dir / "noschema.json", "w") as f:
        json.dump({"foo": 1}, f)
    write_ixfi(vec_dir / "noschema.faiss", np.zeros((1, stubs.DIM), np.float32))
    with open(root / "subset.csv", "w", encoding="utf-8") as f:
        f.write("sha1,cur,company_name,year\n")
        f.write(f"{OTHER_SHA1},CHF,{OTHER},2021\n")
        f.write("J2020,CNY,金盘科技,2020\n")
        f.write("J2023,CNY,金盘科技,notayear\n")
    return {"documents_dir": dirty_dir, "bm25_documents_dir": docs_dir, "vector_db_dir": vec_dir, "bm25_db_dir": bm_dir, "subset_path": root / "subset.csv"}


QUESTIONS = [
    "金盘科技2024年的毛利率是多少 gross margin revenue",
    "overseas factory capacity in mexico and poland 2023",
    "cash flow and dividend compared with previous year",
    "transformer export order backlog energy storage",
]
