"""BASELINE config 1 — "data/test_set chunked reports: faiss IndexFlatIP + BM25Okapi top-30 for the test-set
questions with fixed synthetic 1024-d embeddings (runs on CPU)" (SURVEY 8d, config 1).

The reference ships no chunked reports / indexes for data/test_set (data/test_set/README.md:17-19), only the
five PDFs' page texts quoted in its answer dumps (data/test_set/answers_qwen_max_v2*.json ->
answers[].references[].chunk_text, 64 distinct pages of 5 reports), questions.json and subset.csv.  The
fixture keeps the TOKEN STRUCTURE of that material and none of its text: every whitespace token becomes
`w<id>` (ids in order of first appearance), pages keep their report / page number, the five questions become
id lists the same way.  BM25 sees exactly the term statistics of the real pages; the dense side uses the
deterministic bag-of-words embedder of stubs.py on the `w<id>` strings.

    python -m tests.golden.config1 --fixture     read /root/reference/data/test_set -> config1_corpus.json
    python -m tests.golden.config1 --golden      run the REAL reference retrievers -> config1_outputs.json

Chunking mirrors src/text_splitter.py:70-85 without tiktoken: per page, windows of 300 tokens, overlap 50.
"""
import argparse
import csv
import glob
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.golden import corpus, stubs  # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "config1_corpus.json")
GOLDEN = os.path.join(ROOT, "tests", "golden", "config1_outputs.json")
CHUNK, OVERLAP, TOP_N = 300, 50, 30


# --------------------------------------------------------------------------------------------------
# fixture: token structure of the reference's test-set pages (authoring container only)
# --------------------------------------------------------------------------------------------------
def make_fixture():
    base = os.path.join(REF, "data", "test_set")
    pages = {}
    for path in sorted(glob.glob(os.path.join(base, "answers_*.json"))):
        with open(path, encoding="utf-8") as f:
            d = json.load(f)
        for a in d.get("answers") or d.get("questions") or []:
            for r in a.get("references") or []:
                if r.get("chunk_text"):
                    pages[(r["pdf_sha1"], int(r["page_index"]))] = r["chunk_text"]
    with open(os.path.join(base, "subset.csv"), encoding="utf-8") as f:
        companies = {row["sha1"]: row["company_name"] for row in csv.DictReader(f)}
    with open(os.path.join(base, "questions.json"), encoding="utf-8") as f:
        questions = json.load(f)
    vocab = {}

    def ids(text):
        return [vocab.setdefault(tok, len(vocab)) for tok in text.split()]

    reports = {}
    for (sha1, page), text in sorted(pages.items()):
        reports.setdefault(sha1, []).append({"page": page, "tokens": ids(text)})
    out = {"reports": [{"sha1": s, "company": companies[s], "pages": p} for s, p in sorted(reports.items())],
           "questions": []}
    for q in questions:
        # longest company name contained in the question, as _extract_companies_from_subset (src/questions_processing.py:1021-1040)
        company = max((c for c in companies.values() if c.lower() in q["text"].lower()), key=len)
        out["questions"].append({"company": company, "kind": q["kind"], "tokens": ids(q["text"])})
    out["vocab_size"] = len(vocab)
    with open(FIXTURE, "w", encoding="utf-8") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", FIXTURE, os.path.getsize(FIXTURE), "bytes;", len(pages), "pages,", len(vocab), "distinct tokens")


def load_fixture():
    with open(FIXTURE, encoding="utf-8") as f:
        return json.load(f)


def text_of(tokens):
    return " ".join(f"w{t}" for t in tokens)


# --------------------------------------------------------------------------------------------------
# databases tree in the reference's layout
# --------------------------------------------------------------------------------------------------
def build(root):
    fx = load_fixture()
    root = Path(root)
    docs_dir, vec_dir, bm_dir = (root / "databases" / n for n in ("chunked_reports", "vector_dbs", "bm25_dbs"))
    for d in (docs_dir, vec_dir, bm_dir):
        d.mkdir(parents=True, exist_ok=True)
    for rep in fx["reports"]:
        pages, chunks = [], []
        for pg in rep["pages"]:
            toks = pg["tokens"]
            pages.append({"page": pg["page"], "text": text_of(toks)})
            start = 0
            while True:
                window = toks[start: start + CHUNK]
                chunks.append({"page": pg["page"], "length_tokens": len(window), "text": text_of(window),
                               "id": len(chunks), "type": "content"})
                if start + CHUNK >= len(toks):
                    break
                start += CHUNK - OVERLAP
        meta = {"sha1_name": rep["sha1"], "company_name": rep["company"], "pages_amount": len(pages),
                "text_blocks_amount": len(chunks), "tables_amount": 0, "pictures_amount": 0, "equations_amount": 0,
                "footnotes_amount": 0}
        with open(docs_dir / f"{rep['sha1']}.json", "w", encoding="utf-8") as f:
            json.dump({"metainfo": meta, "content": {"pages": pages, "chunks": chunks}}, f)
        texts = [c["text"] for c in chunks]
        corpus.write_ixfi(vec_dir / f"{rep['sha1']}.faiss", np.stack([stubs.embed_text(t) for t in texts]))
        bm = stubs.OracleBM25([t.split() for t in texts])
        corpus._pickle_as_rank_bm25(
            {"corpus_size": bm.corpus_size, "avgdl": bm.avgdl, "doc_freqs": bm.doc_freqs, "idf": bm.idf,
             "doc_len": bm.doc_len, "tokenizer": None, "k1": 1.5, "b": 0.75, "epsilon": 0.25,
             "average_idf": bm.average_idf}, bm_dir / f"{rep['sha1']}.pkl")
    with open(root / "subset.csv", "w", encoding="utf-8") as f:  # the test set's subset.csv has no year column
        f.write("sha1,cur,company_name\n")
        for rep in fx["reports"]:
            f.write(f"{rep['sha1']},USD,{rep['company']}\n")
    return {"documents_dir": docs_dir, "vector_db_dir": vec_dir, "bm25_db_dir": bm_dir, "subset_path": root / "subset.csv",
            "questions": [{"company": q["company"], "text": text_of(q["tokens"])} for q in fx["questions"]]}


def run_all(R, paths, qwen_kw=None, reranker_kw=None):
    """The config-1 calls, through retrieval module R (the reference's, or this repo's)."""
    from tests.golden import cases

    qwen_kw, reranker_kw = qwen_kw or {}, reranker_kw or {}
    common = dict(vector_db_dir=paths["vector_db_dir"], documents_dir=paths["documents_dir"],
                  subset_path=paths["subset_path"], parallel_workers=1)
    out = {"vector_chunks": [], "vector_pages": [], "rerank": [], "bm25_chunks": [], "bm25_pages": []}
    v = R.VectorRetriever(**common, **qwen_kw)
    h = R.HybridRetriever(**common, **qwen_kw, **reranker_kw)
    b = R.BM25Retriever(paths["bm25_db_dir"], paths["documents_dir"], subset_path=paths["subset_path"])
    for q in paths["questions"]:
        kw = dict(company_name=q["company"], query=q["text"], use_hyde=False, use_multi_query=False)
        out["vector_chunks"].append(cases.normalise(v.retrieve_by_company_name(top_n=TOP_N, **kw)))
        out["vector_pages"].append(cases.normalise(v.retrieve_by_company_name(top_n=TOP_N, return_parent_pages=True, **kw)))
        out["rerank"].append(cases.normalise(h.retrieve_by_company_name(llm_reranking_sample_size=TOP_N, top_n=10, **kw)))
        out["bm25_chunks"].append(cases.normalise(b.retrieve_by_company_name(q["company"], q["text"], top_n=TOP_N)))
        out["bm25_pages"].append(cases.normalise(b.retrieve_by_company_name(q["company"], q["text"], top_n=TOP_N,
                                                                             return_parent_pages=True)))
    return out


def compact(obj):
    """Texts are long `w<id>` strings repeated in every record: keep a digest (the corpus itself is in the fixture)."""
    import hashlib

    if isinstance(obj, dict):
        return {k: ("sha1:" + hashlib.sha1(v.encode("utf-8")).hexdigest()[:16] + f":{len(v)}"
                    if k in ("text", "query", "original_query") and isinstance(v, str) and not v.startswith("sha1:")
                    else compact(v)) for k, v in obj.items()}
    if isinstance(obj, list):
        return [compact(v) for v in obj]
    return obj


def make_golden():
    from tests.golden import make_golden as mg

    mg.install_stubs()
    sys.path.insert(0, REF)
    import src.retrieval as ref

    assert ref.__file__.startswith(REF), ref.__file__
    with tempfile.TemporaryDirectory() as tmp, redirect_stdout(io.StringIO()):
        out = compact(run_all(ref, build(tmp)))
    with open(GOLDEN, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, sort_keys=True, separators=(",", ":"))
    print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--fixture", action="store_true")
    ap.add_argument("--golden", action="store_true")
    a = ap.parse_args()
    if a.fixture:
        make_fixture()
    if a.golden:
        make_golden()
