"""Retriever classes with the reference's interface (src/retrieval.py), backed by the GPU indexes.

Same class names, constructor / retrieve_* signatures, returned record dictionaries, progress
callbacks, timing keys and error behaviour as the reference, so `src/questions_processing.py`,
`src/pipeline.py` and `app_jinpan_qa.py` keep working when `src.retrieval` re-exports this module
(see INTEGRATION.md).  Differences that are invisible to callers:

  * index objects are `FlatIPIndex` / `BM25Okapi` from this package (GPU resident) instead of
    faiss / rank_bm25 objects;
  * reports are loaded once per process and shared by every retriever built on the same
    directories (the reference re-reads every JSON and .faiss file per question,
    src/questions_processing.py:610-647);
  * the (query x report) fan-out of src/retrieval.py:1498-1504 is issued as ONE batched search
    per report instead of one thread-pool task per pair — the hits and their aggregation order
    are deterministic (report order, then query order) instead of thread-completion order;
  * the embedding / LLM / reranker services are injectable (`qwen=`, `llm=`, `reranker=`); when
    the reference's `src.api_requests` / `src.reranking` are importable they are the defaults.

Reference line numbers in the docstrings refer to src/retrieval.py of the reference checkout.
"""
import csv
import gc
import hashlib
import json
import logging
import operator
import re
import threading
import time
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from . import flat_index
from .bm25 import BM25Okapi, load_bm25_pickle  # noqa: F401  (BM25Okapi re-exported)

try:  # CPython helper built by build.sh (csrc/host_records.c): the record loops of _RawHits.records in C
    from . import _b2r_records
except ImportError:  # not built: the Python loops below produce the same records (tests/test_host_records.py)
    _b2r_records = None

_log = logging.getLogger(__name__)

_BY_SIMILARITY = operator.itemgetter("vector_similarity")
_YEAR_IN_QUESTION = re.compile(r"(20\d{2})年")
_YEAR_IN_NAME = re.compile(r"[J]?(20\d{2})")


# --------------------------------------------------------------------------------------------------
# routing (reference :24-165)
# --------------------------------------------------------------------------------------------------
def extract_years_from_question(question: str, expand_window: bool = True) -> List[int]:
    """Years written as `20XX年` in the question; with expand_window the covered range is widened
    by one year on both sides (reference :24-64; worked examples :39-42)."""
    found = [int(y) for y in _YEAR_IN_QUESTION.findall(question)]
    if not found:
        return []
    if expand_window:
        years = range(min(found) - 1, max(found) + 2)
    else:
        years = found
    return sorted(set(years))


def _report_year(metainfo: dict):
    """`metainfo.year` if present, else the first 20XX in sha1_name ("J2025" -> 2025); None when
    neither yields an integer (reference :107-123, :140-157)."""
    year = metainfo.get("year")
    if year is None:
        m = _YEAR_IN_NAME.search(metainfo.get("sha1_name", ""))
        if m:
            year = int(m.group(1))
    if year is None:
        return None
    try:
        return int(year)
    except (ValueError, TypeError):
        return None


def route_reports_by_time(company_name: str, question: str, all_reports: List[Dict],
                          fallback_strategy: str = "all", selected_years: List[int] = None) -> List[Dict]:
    """Company filter, then optional year filter, then the fallback strategy (reference :67-165).
    `question` is accepted for signature compatibility; the reference no longer reads it."""
    company_reports = [r for r in all_reports
                       if r.get("document", {}).get("metainfo", {}).get("company_name") == company_name]
    if not company_reports:
        return []
    if selected_years and len(selected_years) > 0:
        picked = []
        for report in company_reports:
            year = _report_year(report.get("document", {}).get("metainfo", {}))
            if year is not None and year in selected_years:
                picked.append(report)
        if picked:
            return picked
        # no report for the requested years: fall through to the fallback (reference :128-129)
    if fallback_strategy == "latest":
        latest_year, latest = None, []
        for report in company_reports:
            year = _report_year(report.get("document", {}).get("metainfo", {}))
            if year is None:
                continue
            if latest_year is None or year > latest_year:
                latest_year, latest = year, [report]
            elif year == latest_year:
                latest.append(report)
        if latest:
            return latest
    return company_reports


def _load_year_lookup(subset_path) -> dict:
    """sha1 -> year from subset.csv (columns `sha1`, `year`); unreadable file -> {} (reference :464-482)."""
    lookup = {}
    try:
        with open(subset_path, "r", encoding="utf-8") as f:
            for row in csv.DictReader(f):
                sha1 = (row.get("sha1") or "").strip()
                year = (row.get("year") or "").strip()
                if sha1 and year:
                    try:
                        lookup[sha1] = int(year)
                    except ValueError:
                        pass
    except Exception as e:  # noqa: BLE001 - the reference swallows this and continues
        _log.warning("cannot load year information from %s: %s", subset_path, e)
    return lookup


# --------------------------------------------------------------------------------------------------
# offline stand-ins for the network services (the seams are reference :888, :1085-1095, :323)
# --------------------------------------------------------------------------------------------------
class HashEmbedder:
    """Deterministic synthetic embedder with the reference embedder's surface
    (`get_embeddings(texts) -> [{"embedding": [...], "text": ...}]`, src/api_requests.py:118):
    a unit vector seeded by sha1(text).  Stands in for Dashscope text-embedding-v4 (offline)."""

    def __init__(self, dim: int = 1024):
        self.dim = dim

    def embed(self, text: str) -> np.ndarray:
        seed = int(hashlib.sha1(text.encode("utf-8")).hexdigest()[:16], 16)
        v = np.random.default_rng(seed).standard_normal(self.dim).astype(np.float32)
        return v / np.linalg.norm(v)

    def get_embeddings(self, texts, model=None):
        if isinstance(texts, str):
            texts = [texts]
        return [{"embedding": self.embed(t).tolist(), "text": t} for t in texts]

    def send_message(self, *args, **kwargs):
        raise RuntimeError("no LLM service configured (offline); pass llm= to enable query expansion")


class IdentityReranker:
    """Reranker seam (src/reranking.py:286-379) without a network LLM: relevance_score = 1.0, so
    combined_score = round(relevance * vector_similarity, 4) keeps the retrieval order."""

    def __init__(self):
        self._n = 0

    def rerank_documents(self, query, documents, documents_batch_size=4, llm_weight=0.7, progress_callback=None):
        self._n = len(documents)
        out = []
        for doc in documents:
            d = dict(doc)
            d["relevance_score"] = 1.0
            d["reasoning"] = "identity reranker (no LLM configured)"
            d["combined_score"] = round(d["relevance_score"] * doc.get("vector_similarity", doc.get("distance", 0.0)), 4)
            out.append(d)
        out.sort(key=lambda x: x["combined_score"], reverse=True)
        return out

    def get_stats(self):
        return {"total_requests": 0, "successful_requests": 0, "failed_requests": 0, "batch_fallbacks": 0,
                "missing_rankings": 0, "avg_llm_latency": 0.0, "last_error": None, "success_rate": 0.0,
                "request_rate_limit": 0.0, "max_concurrent_requests": 0, "documents": self._n}


def _default_qwen(dim):
    try:
        from src.api_requests import BaseQwenProcessor  # the reference's client, when on sys.path

        return BaseQwenProcessor()
    except Exception:  # noqa: BLE001
        return HashEmbedder(dim)


def _default_reranker():
    try:
        from src.reranking import LLMReranker

        return LLMReranker()
    except Exception:  # noqa: BLE001
        return IdentityReranker()


try:  # glossary context for the multi-query prompts (reference :951-953); prompt text only
    from src.financial_glossary import find_financial_concepts, format_concepts_for_prompt
except Exception:  # noqa: BLE001
    def find_financial_concepts(question, limit=5):
        return []

    def format_concepts_for_prompt(concepts):
        return ""


# --------------------------------------------------------------------------------------------------
# resident report cache (next-row f1: load once, keep shards in HBM)
# --------------------------------------------------------------------------------------------------
_CACHE_LOCK = threading.Lock()
_REPORT_CACHE: Dict[tuple, dict] = {}


def clear_resident_cache():
    with _CACHE_LOCK:
        _REPORT_CACHE.clear()


def _load_report_cached(document_path: Path, index_path: Path, device: int):
    """(document dict, FlatIPIndex) for one report, keyed by path + mtime + size."""
    dst, ist = document_path.stat(), index_path.stat()
    key = (str(document_path.resolve()), dst.st_mtime_ns, dst.st_size, str(index_path.resolve()), ist.st_mtime_ns,
           ist.st_size, device)
    with _CACHE_LOCK:
        hit = _REPORT_CACHE.get(key)
    if hit is not None:
        return hit["document"], hit["vector_db"]
    with open(document_path, "r", encoding="utf-8") as f:
        document = json.load(f)
    if not (isinstance(document, dict) and "metainfo" in document and "content" in document):
        raise ValueError("does not match the expected schema")
    vector_db = flat_index.read_index(index_path, device=device)
    with _CACHE_LOCK:
        # a rewritten report supersedes its older entries (same paths and device, other mtime / size): drop them, or
        # every re-ingestion would leave one more resident copy of the rows in HBM
        for old in [k for k in _REPORT_CACHE if k[0] == key[0] and k[3] == key[3] and k[6] == key[6] and k != key]:
            del _REPORT_CACHE[old]
        _REPORT_CACHE[key] = {"document": document, "vector_db": vector_db}
    return document, vector_db


def _resolve_devices(device):
    if isinstance(device, str):
        if device != "auto":
            raise ValueError("device must be a GPU ordinal, a list of ordinals or 'auto'")
        from . import _lib

        return list(range(max(1, _lib.device_count())))
    if isinstance(device, (list, tuple)):
        if not device:
            raise ValueError("device list is empty")
        return [int(d) for d in device]
    return [int(device)]


def _page_text_lookup(pages):
    """page number -> first page dict with that number (reference does a linear next(...) scan per
    hit, :1145; same result)."""
    lookup = {}
    for page in pages:
        lookup.setdefault(page["page"], page)
    return lookup


# --------------------------------------------------------------------------------------------------
# BM25Retriever (reference :167-295)
# --------------------------------------------------------------------------------------------------
class BM25Retriever:
    def __init__(self, bm25_db_dir: Path, documents_dir: Path, subset_path: Path = None, device: int = 0):
        self.bm25_db_dir = Path(bm25_db_dir)
        self.documents_dir = Path(documents_dir)
        self.subset_path = subset_path
        self.device = device
        self.year_lookup = _load_year_lookup(subset_path) if subset_path else {}
        self._index_cache = {}
        self._lock = threading.Lock()

    def _load_year_lookup(self) -> dict:
        return _load_year_lookup(self.subset_path)

    def _bm25_index(self, path: Path):
        st = path.stat()
        key = (str(path), st.st_mtime_ns, st.st_size)
        with self._lock:
            idx = self._index_cache.get(key)
        if idx is None:
            idx = load_bm25_pickle(path, device=self.device)
            with self._lock:
                self._index_cache[key] = idx
        return idx

    def _documents(self):
        """Every report JSON of the documents directory (reference :199-206 re-reads and re-parses all of them on
        every call); parsed documents and their page lookups stay cached until the file changes.  A file that does
        not parse raises, as it does in the reference."""
        out = []
        for path in self.documents_dir.glob("*.json"):
            st = path.stat()
            key = ("doc", str(path), st.st_mtime_ns, st.st_size)
            with self._lock:
                entry = self._index_cache.get(key)
            if entry is None:
                with open(path, "r", encoding="utf-8") as f:
                    doc = json.load(f)
                entry = {"path": path, "document": doc, "sha1": doc["metainfo"]["sha1_name"],
                         "pages": _page_text_lookup(doc["content"]["pages"])}
                with self._lock:
                    self._index_cache[key] = entry
            out.append(entry)
        return out

    def retrieve_by_company_name(self, company_name: str, query: str, top_n: int = 3,
                                 return_parent_pages: bool = False, selected_years: List[int] = None) -> List[Dict]:
        return self.retrieve_batch(company_name, [query], top_n=top_n, return_parent_pages=return_parent_pages,
                                   selected_years=selected_years)[0]

    def retrieve_batch(self, company_name: str, queries: List[str], top_n: int = 3, return_parent_pages: bool = False,
                       selected_years=None) -> List[List[Dict]]:
        """[retrieve_by_company_name(q) for q in queries] with ONE fused scoring + top-k launch per report for all the
        questions routed to it (reference :194-295 scores one query at a time).  `selected_years` is one list for
        the batch, or a list with one entry (list or None) per question."""
        # tens of thousands of acyclic record dictionaries are born in a batch: pausing the cyclic collector for the
        # call saves its young-generation passes (and full passes over a process holding millions of chunk dicts)
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._retrieve_batch(company_name, queries, top_n, return_parent_pages, selected_years)
        finally:
            if gc_was_enabled:
                gc.enable()

    def _retrieve_batch(self, company_name, queries, top_n, return_parent_pages, selected_years):
        per_question_years = (isinstance(selected_years, (list, tuple)) and len(selected_years) == len(queries)
                              and all(y is None or isinstance(y, (list, tuple)) for y in selected_years))
        all_documents = self._documents()
        all_reports = [{"document": d["document"], "name": d["sha1"]} for d in all_documents]
        routed = []    # per question: the documents to search, in directory order
        for i, query in enumerate(queries):
            years = selected_years[i] if per_question_years else selected_years
            matching = route_reports_by_time(company_name=company_name, question=query, all_reports=all_reports,
                                             fallback_strategy="all", selected_years=years)
            wanted = {rep["name"] for rep in matching}
            docs = [d for d in all_documents if d["sha1"] in wanted]
            if not docs:
                raise ValueError(f"No report found with '{company_name}' company name.")
            routed.append(docs)

        tokenized = [query.split() for query in queries]
        results = [[] for _ in queries]
        for info in all_documents:
            members = [qi for qi, docs in enumerate(routed) if any(d is info for d in docs)]
            if not members:
                continue
            document, sha1 = info["document"], info["sha1"]
            bm25_path = self.bm25_db_dir / f"{sha1}.pkl"
            if not bm25_path.exists():
                for _ in members:
                    print(f"[WARNING] BM25 index not found for {sha1}, skipping")
                continue
            bm25_index = self._bm25_index(bm25_path)
            chunks, pages = document["content"]["chunks"], info["pages"]
            actual_top_n = min(top_n, bm25_index.corpus_size)
            if actual_top_n <= 0:
                continue
            # get_scores + full stable sort (:262-265) fused on the GPU, all routed questions in one launch
            scores, indices = bm25_index.topk([tokenized[qi] for qi in members], actual_top_n)
            scores, indices = np.asarray(scores).tolist(), np.asarray(indices).tolist()
            for row, qi in enumerate(members):
                out = results[qi]
                seen_pages = set()
                for score, index in zip(scores[row], indices[row]):
                    if index < 0:
                        continue
                    score = round(score, 4)
                    chunk = chunks[index]
                    if return_parent_pages:
                        parent = pages[chunk["page"]]
                        if parent["page"] in seen_pages:
                            continue
                        seen_pages.add(parent["page"])
                        out.append({"vector_similarity": score, "page": parent["page"], "text": parent["text"],
                                    "source_sha1": sha1})
                    else:
                        out.append({"vector_similarity": score, "page": chunk["page"], "text": chunk["text"],
                                    "source_sha1": sha1})
        for out in results:
            out.sort(key=lambda x: x["vector_similarity"], reverse=True)
        return [out[:top_n] for out in results]


# --------------------------------------------------------------------------------------------------
# prompts for the optional LLM query expansion.  The wording is ours (prompt text is outside the
# hot path); the protocol is the reference's: HyDE asks for a hypothetical markdown table
# (:899-934); every multi-query method answers with queries wrapped in <...>, or <SKIP>
# (:970-1032).
# --------------------------------------------------------------------------------------------------
_HYDE_SYSTEM = (
    "You analyse corporate annual reports. For the user's question, write a hypothetical markdown table "
    "(header row, separator row, a few sample rows, a unit line such as '单位：万元' when amounts are involved) "
    "of the kind a Chinese financial statement would contain to answer it. Realistic structure matters, "
    "accurate figures do not: the table is only used to retrieve similar tables."
)
_MQ_SYSTEM = ("You help an enterprise RAG system that searches annual reports. Stay financially rigorous and "
              "never change the company name.")
_MQ_METHODS = [
    (1, "synonym",
     "补充财务术语解释：识别问题中的财务专业名词，参考给出的术语表，在原问题后面追加每个术语的定义、近义词和计算方法。"
     "输出1-2个新查询，每个用尖括号<>包裹；若问题不含任何财务术语，输出<SKIP>。"),
    (2, "subquestion",
     "拆分子问题：按财务指标、时间段或业务板块把问题拆成更细的子问题，每个子问题单独用<>包裹；"
     "若无合适拆分，输出<SKIP>。"),
    (3, "variant",
     "情景化改写：仅当原问题比较开放或信息不足时，从不同角度生成1-2个变体提问，每个用<>包裹并保持公司主体不变；"
     "若问题已经明确，输出<SKIP>。"),
]


# --------------------------------------------------------------------------------------------------
# VectorRetriever (reference :431-1858)
# --------------------------------------------------------------------------------------------------
def _raise_reference_unbound_local(name):
    """A hybrid_expansion task of the reference builds its result AFTER its `except` block and reads
    `ssg_total_expanded` / `tri_total_expanded` there (src/retrieval.py:1461,1466); a task that failed before those
    counters were first assigned (:1289, :1356 — e.g. a hit whose parent page is missing from the report JSON) ends
    in UnboundLocalError, which `future.result()` (:1508) re-raises out of retrieve_by_company_name.  Kept as is
    (behaviour visible to callers is preserved; tests/golden/differential.py --corrupt compares it).
    The interpreter raises it, so the message is the running Python's own."""
    if name == "ssg_total_expanded":
        def task():
            if name is None:
                ssg_total_expanded = 0  # noqa: F841
            return ssg_total_expanded  # noqa: F821
    else:
        def task():
            if name is None:
                tri_total_expanded = 0  # noqa: F841
            return tri_total_expanded  # noqa: F821
    task()


class _RawHits:
    """Result rows of one (query, report) basic search, not yet turned into hit tuples (reference :1142-1157):
    row `row` of the float32 [nq, k] scores / int64 [nq, k] ids a batched search returned."""
    __slots__ = ("sha1", "chunks", "pages", "D", "I", "row", "parent_pages", "report_name", "ids_ok")

    def __init__(self, sha1, chunks, pages, D, I, row, parent_pages, report_name, ids_ok=None):
        self.sha1, self.chunks, self.pages = sha1, chunks, pages
        self.D, self.I, self.row = D, I, row
        self.parent_pages, self.report_name, self.ids_ok = parent_pages, report_name, ids_ok

    def __bool__(self):
        return self.I.shape[1] > 0

    # .tolist() yields the same Python floats / ints as float(d) / int(i) per element
    @property
    def scores(self):
        return self.D[self.row].tolist()

    @property
    def ids(self):
        return self.I[self.row].tolist()

    def hits(self, query_text, log):
        """The reference's hit tuples, in rank order; a failing hit ends the list, the hits before it are kept
        (the per-task try / except of the reference, :1441-1442)."""
        out = []
        try:
            for distance, index in zip(self.scores, self.ids):
                out.append(VectorRetriever._hit(self.sha1, self.chunks, self.pages, index, round(distance * 1.0, 4),
                                                "basic", self.parent_pages))
        except Exception as e:  # noqa: BLE001
            log(f"[ERROR] Vector search failed for query '{query_text[:60]}' in report {self.report_name}: {e}")
        return out

    def records(self, query_text, log, helper=True):
        """Final records of a question that has ONE query text and only basic hits: every aggregation key then
        holds one query and one method, so both bonuses are exactly 1.0 and a record is its best hit
        (fuse_similarity / _final_record with count == 1).  Chunk mode needs distinct, valid ids (anything else
        goes through hits()); page mode merges the chunks of a page in first-hit order, as the aggregation does.
        The loops run in the C helper when it is built (`helper=False`: the Python statement of the same)."""
        out = []
        try:
            # the helper reads faiss's result types (float32 scores, int64 ids); any other index object keeps the
            # Python loops, whose floats are the index's own
            if helper and _b2r_records is not None and self.D.dtype == np.float32 and self.I.dtype == np.int64:
                build = _b2r_records.append_page_records if self.parent_pages else _b2r_records.append_chunk_records
                build(out, self.D, self.I, self.row, self.chunks, self.pages, self.sha1, query_text)
            else:
                self._records_py(out, query_text)
        except Exception as e:  # noqa: BLE001
            log(f"[ERROR] Vector search failed for query '{query_text[:60]}' in report {self.report_name}: {e}")
        return out

    def _records_py(self, out, query_text):
        sha1, chunks, pages = self.sha1, self.chunks, self.pages
        if not self.parent_pages:
            for distance, index in zip(self.scores, self.ids):
                chunk = chunks[index]
                if chunk["page"] not in pages:  # a missing parent page fails the hit, as in the reference (:1145)
                    raise KeyError(chunk["page"])
                r = round(distance * 1.0, 4)
                out.append({"vector_similarity": r, "max_original_similarity": r, "page": chunk["page"],
                            "text": chunk["text"], "hit_count": 1, "retrieval_sources": ["basic"],
                            "source_sha1": sha1, "query_sources": [query_text]})
        else:
            by_page = {}
            for distance, index in zip(self.scores, self.ids):
                parent = pages[chunks[index]["page"]]
                r = round(distance * 1.0, 4)
                rec = by_page.get(parent["page"])
                if rec is None:
                    rec = by_page[parent["page"]] = {
                        "vector_similarity": r, "max_original_similarity": r, "page": parent["page"],
                        "text": parent["text"], "hit_count": 1, "retrieval_sources": ["basic"],
                        "source_sha1": sha1, "query_sources": [query_text]}
                    out.append(rec)
                elif r > rec["vector_similarity"]:
                    rec["vector_similarity"] = rec["max_original_similarity"] = r

    def direct_ok(self):
        if self.parent_pages:
            return True
        if self.ids_ok is not None:
            return bool(self.ids_ok)
        ids = self.ids
        return len(set(ids)) == len(ids) and (not ids or min(ids) >= 0)


class VectorRetriever:
    def __init__(self, vector_db_dir: Path, documents_dir: Path, use_hyde: bool = True, use_multi_query: bool = True,
                 subset_path: Path = None, parallel_workers: int = 4,
                 multi_query_methods: Optional[Dict[str, bool]] = None, retrieval_method: str = "basic",
                 max_hops: int = 4, neighbor_k: int = 30, qwen=None, llm=None, device: int = 0):
        self.vector_db_dir = Path(vector_db_dir)
        self.documents_dir = Path(documents_dir)
        self.subset_path = subset_path
        # device: one GPU ordinal, a list of ordinals, or "auto" (every visible GPU).  Report i (sorted by
        # file stem, so J2020..J2025 are stable) lives on devices[i % len(devices)]: one year-DB per GPU.
        self.devices = _resolve_devices(device)
        self.device = self.devices[0]
        self.year_lookup = _load_year_lookup(subset_path) if subset_path else {}
        self.all_dbs = self._load_dbs()
        dim = self.all_dbs[0]["vector_db"].d if self.all_dbs else 1024
        self.qwen = qwen if qwen is not None else _default_qwen(dim)
        self.llm = llm  # LLM used for HyDE / multi-query; None -> self.qwen (reference builds a fresh client, :888)
        self.use_hyde = use_hyde
        self.use_multi_query = use_multi_query
        self.parallel_workers = max(1, parallel_workers)
        self.multi_query_methods = multi_query_methods or {"synonym": True, "subquestion": True, "variant": True}
        self.retrieval_method = retrieval_method
        self.max_hops = max_hops
        self.neighbor_k = neighbor_k

    @classmethod
    def from_reports(cls, all_dbs, qwen=None, llm=None, use_hyde=False, use_multi_query=False, parallel_workers=8,
                     multi_query_methods=None, retrieval_method="basic", max_hops=4, neighbor_k=30):
        """A retriever over reports that are already resident: all_dbs = [{"name", "vector_db", "document"}]
        (the structure `_load_dbs` builds and callers mutate, reference :529-534).  Used for corpora generated
        or sharded straight into HBM."""
        self = cls.__new__(cls)
        self.vector_db_dir = self.documents_dir = self.subset_path = None
        self.devices = sorted({getattr(r["vector_db"], "device", 0) for r in all_dbs}) or [0]
        self.device = self.devices[0]
        self.year_lookup = {}
        self.all_dbs = list(all_dbs)
        dim = self.all_dbs[0]["vector_db"].d if self.all_dbs else 1024
        self.qwen = qwen if qwen is not None else _default_qwen(dim)
        self.llm = llm
        self.use_hyde, self.use_multi_query = use_hyde, use_multi_query
        self.parallel_workers = max(1, parallel_workers)
        self.multi_query_methods = multi_query_methods or {"synonym": True, "subquestion": True, "variant": True}
        self.retrieval_method, self.max_hops, self.neighbor_k = retrieval_method, max_hops, neighbor_k
        return self

    def _load_year_lookup(self) -> dict:
        return _load_year_lookup(self.subset_path)

    def _load_dbs(self):
        """Pair chunked_reports/<stem>.json with vector_dbs/<stem>.faiss (reference :488-541); a report
        whose JSON or index cannot be read is logged and skipped."""
        all_dbs = []
        index_files = {p.stem: p for p in self.vector_db_dir.glob("*.faiss")}
        document_paths = list(self.documents_dir.glob("*.json"))
        placement = {p.stem: self.devices[i % len(self.devices)]
                     for i, p in enumerate(sorted(document_paths, key=lambda p: p.stem))}
        for document_path in document_paths:
            stem = document_path.stem
            if stem not in index_files:
                _log.warning(f"No matching vector DB found for document {document_path.name}")
                continue
            try:
                document, vector_db = _load_report_cached(document_path, index_files[stem], placement[stem])
            except Exception as e:  # noqa: BLE001
                _log.error(f"Error loading report {document_path.name}: {e}")
                continue
            # the cached document is shared: inject the year on a shallow copy of metainfo (:518-521)
            sha1_name = document.get("metainfo", {}).get("sha1_name", stem)
            if sha1_name in self.year_lookup:
                document = dict(document)
                document["metainfo"] = dict(document["metainfo"], year=self.year_lookup[sha1_name])
            all_dbs.append({"name": stem, "vector_db": vector_db, "document": document})
        return all_dbs

    @staticmethod
    def get_strings_cosine_similarity(str1, str2, qwen=None):
        """Cosine of two embedded strings, 4 dp (reference :543-550; that version indexes the result
        as a dict and needs the network — here the embedder is injectable)."""
        qwen = qwen if qwen is not None else _default_qwen(1024)
        emb1 = np.asarray(qwen.get_embeddings([str1])[0]["embedding"])
        emb2 = np.asarray(qwen.get_embeddings([str2])[0]["embedding"])
        score = np.dot(emb1, emb2) / (np.linalg.norm(emb1) * np.linalg.norm(emb2))
        return round(float(score), 4)

    # ---- small helpers (reference :554-583) -------------------------------------------------------
    def _safe_print(self, *args, **kwargs):
        try:
            print(*args, **kwargs)
        except (BrokenPipeError, OSError):
            pass

    def _get_vector_by_id(self, vector_db, doc_id):
        try:
            return vector_db.reconstruct(int(doc_id))
        except Exception as e:  # noqa: BLE001
            self._safe_print(f"[WARNING] Failed to reconstruct vector for ID {doc_id}: {e}")
            return None

    def _normalize_vector(self, vec):
        norm = np.linalg.norm(vec)
        return vec if norm == 0 else vec / norm

    # ---- graph-style expansions (reference :585-806) ----------------------------------------------
    def _neighbours(self, vector_db, vec, neighbor_k, visited):
        """ids of the neighbor_k+1 nearest rows of `vec` that are valid and not yet visited."""
        _, indices = vector_db.search(x=np.asarray(vec, dtype=np.float32).reshape(1, -1), k=neighbor_k + 1)
        return [int(i) for i in indices[0] if int(i) != -1 and int(i) not in visited]

    def _ssg_search(self, vector_db, anchor_id, anchor_vec, max_hops=4, neighbor_k=30):
        """Greedy chunk-to-chunk walk (reference :585-688): from the current chunk, look at its
        neighbours, move to the one most similar to the current chunk, stop as soon as that
        similarity does not exceed the previous hop's (the first comparison is against 1.0)."""
        visited = {int(anchor_id)}
        results = [(1.0, int(anchor_id))]
        details = {"anchor": {"idx": int(anchor_id), "score": None}, "hops": [], "path": [int(anchor_id)],
                   "total_hops": 0, "total_discovered": 1}
        current_idx, current_vec, previous = int(anchor_id), anchor_vec, 1.0
        for hop in range(1, max_hops + 1):
            candidates = []
            for idx in self._neighbours(vector_db, current_vec, neighbor_k, visited):
                vec = self._get_vector_by_id(vector_db, idx)
                if vec is None:
                    continue
                candidates.append({"idx": idx, "score": float(np.dot(current_vec.flatten(), vec.flatten())),
                                   "selected": False})
            if not candidates:
                break
            candidates.sort(key=lambda c: c["score"], reverse=True)
            best = candidates[0]
            if best["score"] <= previous:
                break
            best["selected"] = True
            details["hops"].append({"hop_number": hop, "current_chunk": current_idx, "candidates": candidates[:10],
                                    "selected_idx": best["idx"], "selected_score": best["score"]})
            visited.add(best["idx"])
            details["path"].append(best["idx"])
            current_idx = best["idx"]
            current_vec = self._get_vector_by_id(vector_db, current_idx)
            if current_vec is None:
                break
            results.append((best["score"], current_idx))
            previous = best["score"]
            details["total_discovered"] += 1
        details["total_hops"] = len(details["hops"])
        return {"results": results, "traversal_details": details}

    def _triangulation_search(self, vector_db, query_vec, anchor_id, anchor_vec, max_hops=4, neighbor_k=30):
        """Greedy walk scored by the query / current / candidate triangle (reference :690-806):
        score = 1 / (1 + |centroid - query|), centroid = (query + current + candidate) / 3."""
        visited = {int(anchor_id)}
        q = query_vec.flatten()
        anchor_sim = float(np.dot(q, anchor_vec.flatten()))
        results = [(anchor_sim, int(anchor_id))]
        details = {"anchor": {"idx": int(anchor_id), "score": anchor_sim}, "hops": [], "path": [int(anchor_id)],
                   "total_hops": 0}
        current_idx, current_vec = int(anchor_id), anchor_vec
        for hop in range(1, max_hops + 1):
            candidates = []
            cur = current_vec.flatten()
            for idx in self._neighbours(vector_db, current_vec, neighbor_k, visited):
                vec = self._get_vector_by_id(vector_db, idx)
                if vec is None:
                    continue
                cand = vec.flatten()
                centroid = (q + cur + cand) / 3.0
                dist = float(np.linalg.norm(centroid - q))
                candidates.append({"idx": idx, "query_to_candidate": float(np.dot(q, cand)),
                                   "centroid_score": 1.0 / (1.0 + dist), "centroid_distance": dist, "selected": False})
            if not candidates:
                break
            candidates.sort(key=lambda c: c["centroid_score"], reverse=True)
            best = candidates[0]
            best["selected"] = True
            details["hops"].append({"hop_number": hop, "current_chunk": current_idx, "candidates": candidates[:10],
                                    "selected_idx": best["idx"], "centroid_score": best["centroid_score"],
                                    "selection_reason": "质心距离最优"})
            visited.add(best["idx"])
            details["path"].append(best["idx"])
            current_idx = best["idx"]
            current_vec = self._get_vector_by_id(vector_db, current_idx)
            if current_vec is None:
                break
            results.append((best["centroid_score"], current_idx))
        details["total_hops"] = len(details["hops"])
        return {"results": results, "traversal_details": details}

    def _walks_lockstep(self, vector_db, walks, max_hops=4, neighbor_k=30):
        """Many greedy walks over one report, advanced one hop at a time (next-row f3).

        walks: [{"kind": "ssg" | "triangulation", "anchor_id", "anchor_vec", "query_vec" (triangulation)}].
        Returns, per walk, exactly what _ssg_search / _triangulation_search return.  The walks are
        independent, so hop h of all live walks is ONE batched search (k = neighbor_k + 1) plus ONE
        batched row gather, instead of one search and up to neighbor_k + 1 reconstruct round trips per
        walk and hop (reference :622-631, :730-738 — hybrid_expansion runs 30 walks per query and report)."""
        states = []
        for w in walks:
            aid, avec = int(w["anchor_id"]), w["anchor_vec"]
            if w["kind"] == "ssg":
                states.append({"kind": "ssg", "visited": {aid}, "results": [(1.0, aid)], "cur_idx": aid, "cur_vec": avec,
                               "previous": 1.0, "alive": True,
                               "details": {"anchor": {"idx": aid, "score": None}, "hops": [], "path": [aid],
                                           "total_hops": 0, "total_discovered": 1}})
            else:
                q = w["query_vec"].flatten()
                anchor_sim = float(np.dot(q, avec.flatten()))
                states.append({"kind": "triangulation", "q": q, "visited": {aid}, "results": [(anchor_sim, aid)],
                               "cur_idx": aid, "cur_vec": avec, "alive": True,
                               "details": {"anchor": {"idx": aid, "score": anchor_sim}, "hops": [], "path": [aid],
                                           "total_hops": 0}})
        for hop in range(1, max_hops + 1):
            live = [st for st in states if st["alive"]]
            if not live:
                break
            X = np.stack([np.asarray(st["cur_vec"], dtype=np.float32).reshape(-1) for st in live])
            _, I = vector_db.search(x=X, k=neighbor_k + 1)
            neigh = [[int(i) for i in I[r] if int(i) != -1 and int(i) not in st["visited"]] for r, st in enumerate(live)]
            need = sorted({i for ids in neigh for i in ids})
            vecs = {}
            if need:
                rows = vector_db.reconstruct_batch(np.asarray(need, dtype=np.int64))
                vecs = {i: rows[j] for j, i in enumerate(need)}
            for st, ids in zip(live, neigh):
                cur = st["cur_vec"].flatten()
                candidates = []
                if st["kind"] == "ssg":
                    for idx in ids:
                        candidates.append({"idx": idx, "score": float(np.dot(cur, vecs[idx].flatten())), "selected": False})
                    if not candidates:
                        st["alive"] = False
                        continue
                    candidates.sort(key=lambda c: c["score"], reverse=True)
                    best = candidates[0]
                    if best["score"] <= st["previous"]:
                        st["alive"] = False
                        continue
                    best["selected"] = True
                    st["details"]["hops"].append({"hop_number": hop, "current_chunk": st["cur_idx"],
                                                  "candidates": candidates[:10], "selected_idx": best["idx"],
                                                  "selected_score": best["score"]})
                    st["visited"].add(best["idx"])
                    st["details"]["path"].append(best["idx"])
                    st["cur_idx"], st["cur_vec"] = best["idx"], vecs[best["idx"]]
                    st["results"].append((best["score"], best["idx"]))
                    st["previous"] = best["score"]
                    st["details"]["total_discovered"] += 1
                else:
                    q = st["q"]
                    for idx in ids:
                        cand = vecs[idx].flatten()
                        centroid = (q + cur + cand) / 3.0
                        dist = float(np.linalg.norm(centroid - q))
                        candidates.append({"idx": idx, "query_to_candidate": float(np.dot(q, cand)),
                                           "centroid_score": 1.0 / (1.0 + dist), "centroid_distance": dist,
                                           "selected": False})
                    if not candidates:
                        st["alive"] = False
                        continue
                    candidates.sort(key=lambda c: c["centroid_score"], reverse=True)
                    best = candidates[0]
                    best["selected"] = True
                    st["details"]["hops"].append({"hop_number": hop, "current_chunk": st["cur_idx"],
                                                  "candidates": candidates[:10], "selected_idx": best["idx"],
                                                  "centroid_score": best["centroid_score"],
                                                  "selection_reason": "质心距离最优"})
                    st["visited"].add(best["idx"])
                    st["details"]["path"].append(best["idx"])
                    st["cur_idx"], st["cur_vec"] = best["idx"], vecs[best["idx"]]
                    st["results"].append((best["centroid_score"], best["idx"]))
        out = []
        for st in states:
            st["details"]["total_hops"] = len(st["details"]["hops"])
            out.append({"results": st["results"], "traversal_details": st["details"]})
        return out

    # ---- query expansion (reference :892-1047) ------------------------------------------------------
    def _expand_queries(self, query, use_hyde, use_multi_query, multi_query_config, timing, expansion_texts,
                        progress_callback):
        llm = self.llm if self.llm is not None else self.qwen
        queries = [query]
        if use_hyde:
            if progress_callback:
                progress_callback("🔮 HYDE 查询扩展中...", 32)
            t0 = time.time()
            try:
                fake = llm.send_message(
                    model="qwen-turbo", system_content=_HYDE_SYSTEM,
                    human_content=("Write a markdown table that could appear in the company's financial report and "
                                   f"would answer this question: {query}\nInclude a unit line, fitting column headers, "
                                   "sample rows and any formula notes."),
                    is_structured=False)
                fake = "".join(fake) if isinstance(fake, list) else str(fake)
                queries.append(fake)
                expansion_texts["hyde_text"] = fake
            except Exception as e:  # noqa: BLE001
                self._safe_print(f"[ERROR] HYDE expansion failed: {e}")
            timing["hyde_expansion"] = time.time() - t0
        if use_multi_query:
            if progress_callback:
                progress_callback("🔄 Multi-Query 查询扩展中...", 38)
            t0 = time.time()
            concepts = find_financial_concepts(query, limit=5)
            concept_terms = [c["term"] for c in concepts]
            glossary = format_concepts_for_prompt(concepts)
            expansion_texts["glossary_context"] = glossary
            expansion_texts["multi_query_methods"] = multi_query_config
            for method_id, key, instruction in _MQ_METHODS:
                if not multi_query_config.get(key, False):
                    continue
                try:
                    response = llm.send_message(
                        model="qwen-turbo", system_content=_MQ_SYSTEM,
                        human_content=f"{instruction}\n\nFinancial glossary:\n{glossary}\n\nOriginal question: {query}",
                        is_structured=False)
                    for q in re.findall(r"<(.*?)>", response, flags=re.DOTALL):
                        q = q.strip()
                        if not q or q.upper() == "SKIP":
                            continue
                        queries.append(q)
                        expansion_texts["multi_query_texts"].append({"method_id": method_id, "query": q,
                                                                     "concepts": concept_terms})
                except Exception as e:  # noqa: BLE001
                    self._safe_print(f"Expansion method {method_id} failed: {e}")
            timing["multi_query_expansion"] = time.time() - t0
        deduped, seen = [], set()
        for q in queries:
            q = q.strip()
            if q and q not in seen:
                deduped.append(q)
                seen.add(q)
        return deduped

    def _embed_queries(self, queries):
        """query text -> float32 [1, d] (reference :1081-1098); failures are reported and skipped."""
        out = {}
        for q in queries:
            try:
                res = self.qwen.get_embeddings([q])
                if not res or not isinstance(res, list) or not res[0] or "embedding" not in res[0]:
                    self._safe_print(f"[ERROR] embedding result is empty or invalid for query: {q[:80]}")
                    continue
                out[q] = np.array(res[0]["embedding"], dtype=np.float32).reshape(1, -1)
            except Exception as e:  # noqa: BLE001
                self._safe_print(f"[ERROR] Failed to get embedding for query snippet '{q[:50]}': {e}")
        return out

    # ---- per-report search (reference :1119-1473) ----------------------------------------------------
    @staticmethod
    def _hit(sha1, chunks, pages, index, similarity, source, return_parent_pages):
        chunk = chunks[index]
        parent = pages[chunk["page"]]  # KeyError (-> task failure) if the page is missing, like next() in the reference
        if return_parent_pages:
            return ((sha1, "page", parent["page"]), parent["page"], parent["text"], similarity, sha1, source)
        return ((sha1, "chunk", index), chunk["page"], chunk["text"], similarity, sha1, source)

    def _query_similarity(self, vector_db, embedding_array, idx, fallback_score):
        vec = self._get_vector_by_id(vector_db, idx)
        if vec is not None:
            return round(float(np.dot(embedding_array.flatten(), vec.flatten())) * 1.0, 4)
        return round(float(fallback_score) * 1.0, 4)

    def _search_report(self, report, query_items, top_n, return_parent_pages, retrieval_method, max_hops, neighbor_k):
        """All queries against one report.  Returns a list of (query_text, result) in query order,
        where result is the list of hit tuples, or a dict {"hits", "traversal_details",
        "algorithm_specific_results"} exactly as the reference's per-task return value (:1444-1473)."""
        document, vector_db = report["document"], report["vector_db"]
        chunks = document["content"]["chunks"]
        # page lookup built once per resident report (rebuilt if the caller swapped the pages list)
        cached = report.get("_page_lookup")
        if cached is None or cached[0] != id(document["content"]["pages"]):
            cached = report["_page_lookup"] = (id(document["content"]["pages"]),
                                               _page_text_lookup(document["content"]["pages"]))
        pages = cached[1]
        sha1 = document["metainfo"]["sha1_name"]
        actual_top_n = min(top_n, len(chunks))
        if actual_top_n == 0:
            return [(q, []) for q, _ in query_items]
        texts = [q for q, _ in query_items]
        X = np.concatenate([e for _, e in query_items], axis=0)
        out = []
        basic_like = retrieval_method not in ("ssg", "triangulation", "hybrid_expansion")
        if basic_like:
            if retrieval_method != "basic":
                self._safe_print(f"[WARNING] Unknown retrieval_method '{retrieval_method}', falling back to basic")
            try:
                D, I = vector_db.search(x=X, k=actual_top_n)  # ONE batched launch for the whole fan-out
            except Exception as e:  # noqa: BLE001
                self._safe_print(f"[ERROR] Vector search failed in report {report.get('name')}: {e}")
                return [(q, []) for q in texts]
            # The rows stay raw here (23k hits for a 127-question batch over six databases): _finalize turns
            # them into records directly when it can, into hit tuples otherwise (_RawHits).
            D, I = np.ascontiguousarray(D), np.ascontiguousarray(I)   # dtypes as the index returned them
            ids_ok = None
            if I.shape[1] > 1:  # distinct, non-negative ids per row, for the whole batch at once
                S = np.sort(I, axis=1)
                ids_ok = ((S[:, 0] >= 0) & (S[:, 1:] != S[:, :-1]).all(axis=1)).tolist()
            name = report.get("name")
            return [(q, _RawHits(sha1, chunks, pages, D, I, qi, return_parent_pages, name,
                                 None if ids_ok is None else ids_ok[qi]))
                    for qi, q in enumerate(texts)]

        if getattr(self, "batched_expansion", True):
            try:
                return self._search_report_walks_lockstep(report, query_items, return_parent_pages, retrieval_method,
                                                          max_hops, neighbor_k)
            except Exception as e:  # noqa: BLE001 - e.g. an index object without reconstruct_batch
                self._safe_print(f"[WARNING] batched expansion failed in report {report.get('name')} ({e}); "
                                 "walking sequentially")
        return self._search_report_walks_sequential(report, query_items, return_parent_pages, retrieval_method,
                                                    max_hops, neighbor_k)

    def _search_report_walks_lockstep(self, report, query_items, return_parent_pages, retrieval_method, max_hops,
                                      neighbor_k):
        """ssg / triangulation / hybrid_expansion for all queries of one report with batched GPU calls:
        one anchor (or top-50) search for all queries, one row gather for all anchors, then every walk of
        every query in lockstep (_walks_lockstep).  Produces the per-query values of the sequential code."""
        document, vector_db = report["document"], report["vector_db"]
        chunks = document["content"]["chunks"]
        pages = report["_page_lookup"][1]
        sha1 = document["metainfo"]["sha1_name"]
        X = np.concatenate([e for _, e in query_items], axis=0)
        nq = len(query_items)
        out = [None] * nq

        def exp_hits(qi, exp, source):
            """(hit, walk score) of every step of a walk, lazily: a failing record (e.g. a chunk whose page is missing)
            stops the (query, report) task where the reference's try / except does, keeping the hits before it."""
            emb = query_items[qi][1]
            for score, idx in exp.get("results", []):
                if idx == -1:
                    continue
                sim = self._query_similarity_vec(emb, idx, score, row_of)
                yield self._hit(sha1, chunks, pages, idx, sim, source, return_parent_pages), score

        row_of = {}   # id -> stored row, filled by the batched gathers below

        def gather(ids):
            need = sorted({int(i) for i in ids if int(i) not in row_of})
            if need:
                rows = vector_db.reconstruct_batch(np.asarray(need, dtype=np.int64))
                for j, i in enumerate(need):
                    row_of[i] = rows[j]

        if retrieval_method in ("ssg", "triangulation"):
            D, I = vector_db.search(x=X, k=1)
            live = [qi for qi in range(nq) if len(D[qi]) > 0 and int(I[qi][0]) != -1]
            for qi in range(nq):
                if qi not in live:
                    out[qi] = (query_items[qi][0], [])
            gather(int(I[qi][0]) for qi in live)
            walks = [{"kind": retrieval_method, "anchor_id": int(I[qi][0]), "anchor_vec": row_of[int(I[qi][0])],
                      "query_vec": query_items[qi][1]} for qi in live]
            exps = self._walks_lockstep(vector_db, walks, max_hops=max_hops, neighbor_k=neighbor_k)
            gather(idx for exp in exps for _, idx in exp["results"])
            for qi, exp in zip(live, exps):
                q = query_items[qi][0]
                hits = []
                details = exp["traversal_details"]
                details["anchor"]["score"] = float(D[qi][0])
                try:
                    for h, _ in exp_hits(qi, exp, retrieval_method):
                        hits.append(h)
                except Exception as e:  # noqa: BLE001
                    self._safe_print(f"[ERROR] Vector search failed for query '{q[:60]}' in report {report.get('name')}: {e}")
                out[qi] = (q, {"hits": hits, "traversal_details": details})
            return out

        # hybrid_expansion: basic top-50, SSG from the top 10, triangulation from the top 20 (:1258-1418)
        D, I = vector_db.search(x=X, k=min(50, len(chunks)))
        per_q = []
        walks, owner = [], []
        for qi, (q, emb) in enumerate(query_items):
            hits, basic_results, basic_keys = [], [], set()
            failed = None
            try:
                for distance, index in zip(D[qi], I[qi]):
                    h = self._hit(sha1, chunks, pages, int(index), round(float(distance) * 1.0, 4), "basic",
                                  return_parent_pages)
                    basic_results.append((h[0], h[1], h[2], h[3], sha1, int(index)))
                    basic_keys.add(h[0])
                    hits.append(h)
            except Exception as e:  # noqa: BLE001
                failed = e
                self._safe_print(f"[ERROR] Vector search failed for query '{q[:60]}' in report {report.get('name')}: {e}")
            per_q.append({"hits": hits, "basic": basic_results, "keys": basic_keys, "failed": failed,
                          "unbound": "ssg_total_expanded" if failed is not None else None})
        gather(b[5] for st in per_q if st["failed"] is None for b in st["basic"][:20])
        for qi, st in enumerate(per_q):
            if st["failed"] is not None:
                continue
            for limit, source in ((10, "ssg"), (20, "triangulation")):
                for key, page_id, text, sim0, _sha, idx in st["basic"][:limit]:
                    walks.append({"kind": source, "anchor_id": idx, "anchor_vec": row_of[idx],
                                  "query_vec": query_items[qi][1]})
                    owner.append((qi, source, page_id))
        exps = self._walks_lockstep(vector_db, walks, max_hops=max_hops, neighbor_k=neighbor_k) if walks else []
        gather(idx for exp in exps for _, idx in exp["results"])
        stats = [{"ssg_new": [], "tri_new": [], "ssg_total": 0, "ssg_new_only": 0, "tri_total": 0, "tri_new_only": 0}
                 for _ in range(nq)]
        for (qi, source, page_id), exp in zip(owner, exps):
            st, acc = per_q[qi], stats[qi]
            if st["failed"] is not None:
                continue
            try:
                for h, score in exp_hits(qi, exp, source):
                    st["hits"].append(h)
                    is_new = h[0] not in st["keys"]
                    rec = {"key": h[0], "page": h[1], "text": h[2], "vector_similarity": h[3], "source_sha1": sha1,
                           "anchor_page": page_id, "score": score}
                    if source == "ssg":
                        acc["ssg_total"] += 1
                        if is_new:
                            acc["ssg_new"].append(rec)
                            acc["ssg_new_only"] += 1
                    else:
                        acc["tri_total"] += 1
                        if is_new:
                            acc["tri_new"].append(rec)
                            acc["tri_new_only"] += 1
            except Exception as e:  # noqa: BLE001 - the reference abandons the rest of this (query, report) task
                st["failed"] = e
                if source == "ssg":  # the triangulation counters were never assigned (:1356)
                    st["unbound"] = "tri_total_expanded"
                q = query_items[qi][0]
                self._safe_print(f"[ERROR] Vector search failed for query '{q[:60]}' in report {report.get('name')}: {e}")
        for qi, (q, _) in enumerate(query_items):
            st, acc = per_q[qi], stats[qi]
            out[qi] = (q, {"hits": st["hits"], "algorithm_specific_results": {
                "ssg_new_chunks": acc["ssg_new"], "triangulation_new_chunks": acc["tri_new"],
                "basic_count": len(st["basic"]),
                "ssg_stats": {"total_expanded": acc["ssg_total"], "new_only": acc["ssg_new_only"],
                              "in_basic_top50": acc["ssg_total"] - acc["ssg_new_only"]},
                "triangulation_stats": {"total_expanded": acc["tri_total"], "new_only": acc["tri_new_only"],
                                        "in_basic_top50": acc["tri_total"] - acc["tri_new_only"]}}})
            if st["unbound"]:
                out[qi][1]["_reference_unbound_local"] = st["unbound"]
        return out

    def _query_similarity_vec(self, embedding_array, idx, fallback_score, row_of):
        vec = row_of.get(int(idx))
        if vec is not None:
            return round(float(np.dot(embedding_array.flatten(), vec.flatten())) * 1.0, 4)
        return round(float(fallback_score) * 1.0, 4)

    def _search_report_walks_sequential(self, report, query_items, return_parent_pages, retrieval_method, max_hops,
                                        neighbor_k):
        """One walk at a time, as the reference does (kept as the fallback of the lockstep path)."""
        document, vector_db = report["document"], report["vector_db"]
        chunks = document["content"]["chunks"]
        pages = report["_page_lookup"][1]
        sha1 = document["metainfo"]["sha1_name"]
        out = []
        for qi, (q, emb) in enumerate(query_items):
            hits, trav = [], []
            ssg_new, tri_new, basic_results = [], [], []
            ssg_total = ssg_new_only = tri_total = tri_new_only = 0
            try:
                if retrieval_method in ("ssg", "triangulation"):
                    D, I = vector_db.search(x=emb, k=1)
                    if not (len(D[0]) > 0 and I[0][0] != -1):
                        out.append((q, []))
                        continue
                    anchor_idx, anchor_score = int(I[0][0]), float(D[0][0])
                    anchor_vec = self._get_vector_by_id(vector_db, anchor_idx)
                    if anchor_vec is None:
                        out.append((q, []))
                        continue
                    if retrieval_method == "ssg":
                        exp = self._ssg_search(vector_db, anchor_idx, anchor_vec, max_hops=max_hops, neighbor_k=neighbor_k)
                    else:
                        exp = self._triangulation_search(vector_db, emb, anchor_idx, anchor_vec, max_hops=max_hops,
                                                         neighbor_k=neighbor_k)
                    details = exp.get("traversal_details")
                    if details:
                        details["anchor"]["score"] = anchor_score
                        trav.append(details)
                    for score, idx in exp.get("results", []):
                        if idx == -1:
                            continue
                        sim = self._query_similarity(vector_db, emb, idx, score)
                        hits.append(self._hit(sha1, chunks, pages, idx, sim, retrieval_method, return_parent_pages))
                else:  # hybrid_expansion: basic top-50, SSG from the top 10, triangulation from the top 20 (:1258-1418)
                    unbound = "ssg_total_expanded"   # what the reference has not assigned yet if the task fails now
                    D, I = vector_db.search(x=emb, k=min(50, len(chunks)))
                    basic_keys = set()
                    for distance, index in zip(D[0], I[0]):
                        h = self._hit(sha1, chunks, pages, int(index), round(float(distance) * 1.0, 4), "basic",
                                      return_parent_pages)
                        basic_results.append((h[0], h[1], h[2], h[3], sha1, int(index)))
                        basic_keys.add(h[0])
                        hits.append(h)
                    for limit, source in ((10, "ssg"), (20, "triangulation")):
                        unbound = "tri_total_expanded" if source == "ssg" else None
                        for key, page_id, text, sim0, _sha, idx in basic_results[:limit]:
                            vec = self._get_vector_by_id(vector_db, idx)
                            if vec is None:
                                continue
                            if source == "ssg":
                                exp = self._ssg_search(vector_db, idx, vec, max_hops=max_hops, neighbor_k=neighbor_k)
                            else:
                                exp = self._triangulation_search(vector_db, emb, idx, vec, max_hops=max_hops,
                                                                 neighbor_k=neighbor_k)
                            for score, eidx in exp.get("results", []):
                                if eidx == -1:
                                    continue
                                sim = self._query_similarity(vector_db, emb, eidx, score)
                                h = self._hit(sha1, chunks, pages, eidx, sim, source, return_parent_pages)
                                hits.append(h)
                                is_new = h[0] not in basic_keys
                                rec = {"key": h[0], "page": h[1], "text": h[2], "vector_similarity": sim,
                                       "source_sha1": sha1, "anchor_page": page_id, "score": score}
                                if source == "ssg":
                                    ssg_total += 1
                                    if is_new:
                                        ssg_new.append(rec)
                                        ssg_new_only += 1
                                else:
                                    tri_total += 1
                                    if is_new:
                                        tri_new.append(rec)
                                        tri_new_only += 1
                    unbound = None
            except Exception as e:  # noqa: BLE001
                self._safe_print(f"[ERROR] Vector search failed for query '{q[:60]}' in report {report.get('name')}: {e}")
            else:
                unbound = None
            if trav or retrieval_method == "hybrid_expansion":
                result = {"hits": hits}
                if retrieval_method == "hybrid_expansion" and unbound:
                    result["_reference_unbound_local"] = unbound
                if trav:
                    result["traversal_details"] = trav[0] if len(trav) == 1 else trav
                if retrieval_method == "hybrid_expansion":
                    result["algorithm_specific_results"] = {
                        "ssg_new_chunks": ssg_new, "triangulation_new_chunks": tri_new,
                        "basic_count": len(basic_results),
                        "ssg_stats": {"total_expanded": ssg_total, "new_only": ssg_new_only,
                                      "in_basic_top50": ssg_total - ssg_new_only},
                        "triangulation_stats": {"total_expanded": tri_total, "new_only": tri_new_only,
                                                "in_basic_top50": tri_total - tri_new_only}}
                out.append((q, result))
            else:
                out.append((q, hits))
        return out

    # ---- the public retrieval call (reference :808-1819) -------------------------------------------------
    def retrieve_by_company_name(self, company_name: str, query: str, llm_reranking_sample_size: int = None,
                                 top_n: int = 3, return_parent_pages: bool = False, use_hyde: bool = None,
                                 use_multi_query: bool = None, progress_callback=None,
                                 selected_years: List[int] = None, multi_query_config: Optional[Dict[str, bool]] = None,
                                 retrieval_method: str = "basic", max_hops: int = 2, neighbor_k: int = 10):
        plan = self._prepare(company_name, query, use_hyde, use_multi_query, progress_callback, selected_years,
                             multi_query_config, retrieval_method, max_hops, neighbor_k)
        if progress_callback:
            progress_callback("🔎 向量检索中...", 45)
        t0 = time.time()
        per_report = self._search_reports(plan["reports"], [plan["query_items"]] * len(plan["reports"]), top_n,
                                          return_parent_pages, plan["retrieval_method"], plan["max_hops"],
                                          plan["neighbor_k"])
        plan["timing"]["vector_search"] = time.time() - t0
        return self._finalize(plan, per_report, top_n, progress_callback)

    def retrieve_batch(self, company_name: str, queries: List[str], top_n: int = 3, return_parent_pages: bool = False,
                       use_hyde: bool = None, use_multi_query: bool = None, selected_years=None,
                       multi_query_config: Optional[Dict[str, bool]] = None, retrieval_method: str = "basic",
                       max_hops: int = 2, neighbor_k: int = 10) -> List[Dict]:
        """A whole question batch (the reference loops over questions, src/questions_processing.py:1104-1138,
        and re-creates the retriever per question).  Returns [retrieve_by_company_name(q) for q in queries],
        but every report is searched ONCE with the embeddings of all the questions routed to it (one
        tensor-core batch per year-DB, the year-DBs on their own GPUs in parallel).  `selected_years` is one
        list for the batch, or a list with one entry (list or None) per question."""
        gc_was_enabled = gc.isenabled()   # see BM25Retriever.retrieve_batch
        gc.disable()
        try:
            return self._retrieve_batch(company_name, queries, top_n, return_parent_pages, use_hyde, use_multi_query,
                                        selected_years, multi_query_config, retrieval_method, max_hops, neighbor_k)
        finally:
            if gc_was_enabled:
                gc.enable()

    def _retrieve_batch(self, company_name, queries, top_n, return_parent_pages, use_hyde, use_multi_query,
                        selected_years, multi_query_config, retrieval_method, max_hops, neighbor_k):
        per_question_years = (isinstance(selected_years, (list, tuple)) and len(selected_years) == len(queries)
                              and all(y is None or isinstance(y, (list, tuple)) for y in selected_years))
        plans = []
        for i, query in enumerate(queries):
            years = selected_years[i] if per_question_years else selected_years
            plans.append(self._prepare(company_name, query, use_hyde, use_multi_query, None, years, multi_query_config,
                                       retrieval_method, max_hops, neighbor_k))
        if any(p["retrieval_method"] in ("ssg", "triangulation", "hybrid_expansion") for p in plans):
            # graph walks are sequential per (question, report): no cross-question batching
            out = []
            for p in plans:
                t0 = time.time()
                per_report = self._search_reports(p["reports"], [p["query_items"]] * len(p["reports"]), top_n,
                                                  return_parent_pages, p["retrieval_method"], p["max_hops"],
                                                  p["neighbor_k"])
                p["timing"]["vector_search"] = time.time() - t0
                out.append(self._finalize(p, per_report, top_n, None))
            return out
        # group the (question, expansion) embeddings by report
        reports, items, owners = [], [], []   # owners[r] = [(plan index, n_items)]
        slot = {}
        for pi, p in enumerate(plans):
            for report in p["reports"]:
                r = slot.get(id(report))
                if r is None:
                    r = slot[id(report)] = len(reports)
                    reports.append(report)
                    items.append([])
                    owners.append([])
                items[r].extend(p["query_items"])
                owners[r].append((pi, len(p["query_items"])))
        t0 = time.time()
        searched = self._search_reports(reports, items, top_n, return_parent_pages, "basic", max_hops, neighbor_k)
        elapsed = time.time() - t0
        per_plan = [[] for _ in plans]
        for r, res in enumerate(searched):
            pos = 0
            for pi, cnt in owners[r]:
                per_plan[pi].append(None if res is None else res[pos: pos + cnt])
                pos += cnt
        out = []
        for p, per_report in zip(plans, per_plan):
            p["timing"]["vector_search"] = elapsed / max(1, len(plans))
            out.append(self._finalize(p, per_report, top_n, None))
        return out

    def _prepare(self, company_name, query, use_hyde, use_multi_query, progress_callback, selected_years,
                 multi_query_config, retrieval_method, max_hops, neighbor_k):
        """Routing, expansion and embedding of one question (reference :813-1101)."""
        timing_info = {"hyde_expansion": 0.0, "multi_query_expansion": 0.0, "embedding_generation": 0.0,
                       "vector_search": 0.0}
        expansion_texts = {"hyde_text": None, "multi_query_texts": [], "glossary_context": None,
                           "multi_query_methods": {}}
        if progress_callback:
            progress_callback("📚 定位相关文档中...", 28)
        matching_reports = route_reports_by_time(company_name=company_name, question=query, all_reports=self.all_dbs,
                                                 fallback_strategy="all", selected_years=selected_years)
        if not matching_reports:
            _log.error(f"No report found with '{company_name}' company name.")
            raise ValueError(f"No report found with '{company_name}' company name.")

        use_hyde = self.use_hyde if use_hyde is None else use_hyde
        use_multi_query = self.use_multi_query if use_multi_query is None else use_multi_query
        multi_query_config = multi_query_config or self.multi_query_methods or {}
        expansion_texts["multi_query_methods"] = multi_query_config
        if use_multi_query and not any(multi_query_config.values()):
            use_multi_query = False
        # default-value sniffing kept as in the reference (:868-885): an argument left at "basic" / 4 / 30
        # defers to the instance attribute (the UI mutates those after construction)
        if retrieval_method == "basic" and getattr(self, "retrieval_method", "basic") != "basic":
            retrieval_method = self.retrieval_method
        if max_hops == 4 and getattr(self, "max_hops", 4) != 4:
            max_hops = self.max_hops
        if neighbor_k == 30 and getattr(self, "neighbor_k", 30) != 30:
            neighbor_k = self.neighbor_k

        queries = self._expand_queries(query, use_hyde, use_multi_query, multi_query_config, timing_info,
                                       expansion_texts, progress_callback)
        t0 = time.time()
        query_embeddings = self._embed_queries(queries)
        timing_info["embedding_generation"] = time.time() - t0
        if not query_embeddings:
            raise ValueError("Failed to generate embeddings for all queries.")
        return {"reports": matching_reports, "query_items": list(query_embeddings.items()), "timing": timing_info,
                "expansion_texts": expansion_texts, "retrieval_method": retrieval_method, "max_hops": max_hops,
                "neighbor_k": neighbor_k}

    def _search_reports(self, reports, items_per_report, top_n, return_parent_pages, retrieval_method, max_hops,
                        neighbor_k):
        """_search_report for every report; reports on different GPUs run concurrently on a thread pool
        (ctypes releases the GIL during the search — the reference's ThreadPoolExecutor fan-out, :1498-1504,
        at report granularity).  Returns per report the (query_text, result) list, or None if it failed."""
        def one(args):
            report, query_items = args
            try:
                return self._search_report(report, query_items, top_n, return_parent_pages, retrieval_method,
                                           max_hops, neighbor_k)
            except Exception as e:  # noqa: BLE001 - a failed report must not fail the retrieval (:1441-1442)
                self._safe_print(f"[ERROR] Vector search failed in report {report.get('name')}: {e}")
                return None

        work = list(zip(reports, items_per_report))
        gpus = {getattr(r["vector_db"], "device", 0) for r in reports}
        workers = min(self.parallel_workers, len(work), max(1, len(gpus)))
        if workers <= 1:
            return [one(w) for w in work]
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=workers) as pool:
            return list(pool.map(one, work))  # report order, not completion order

    def _finalize(self, plan, per_report, top_n, progress_callback):
        """Aggregation over reports and queries, fusion, sort, cut (reference :1552-1819)."""
        retrieval_method, max_hops, neighbor_k = plan["retrieval_method"], plan["max_hops"], plan["neighbor_k"]
        aggregated = {}
        traversal_details = []
        algo = {"ssg_new_chunks": [], "triangulation_new_chunks": [], "basic_count": 0,
                "ssg_stats": {"total_expanded": 0, "new_only": 0, "in_basic_top50": 0},
                "triangulation_stats": {"total_expanded": 0, "new_only": 0, "in_basic_top50": 0}}
        direct = self._direct_records(per_report)
        for per_query in (() if direct is not None else per_report):
            if per_query is None:
                continue
            for query_text, result in per_query:
                if not result:
                    continue
                if isinstance(result, _RawHits):
                    result = result.hits(query_text, self._safe_print)
                    if not result:
                        continue
                if isinstance(result, dict):
                    if result.get("_reference_unbound_local"):
                        _raise_reference_unbound_local(result["_reference_unbound_local"])
                    hits = result.get("hits", [])
                    if result.get("traversal_details"):
                        traversal_details.append(result["traversal_details"])
                    spec = result.get("algorithm_specific_results")
                    if spec:
                        algo["ssg_new_chunks"].extend(spec.get("ssg_new_chunks", []))
                        algo["triangulation_new_chunks"].extend(spec.get("triangulation_new_chunks", []))
                        algo["basic_count"] += spec.get("basic_count", 0)
                        for name in ("ssg_stats", "triangulation_stats"):
                            for field in ("total_expanded", "new_only", "in_basic_top50"):
                                algo[name][field] += spec.get(name, {}).get(field, 0)
                else:
                    hits = result
                self._aggregate(aggregated, hits, query_text)

        final_results = direct if direct is not None else [self._final_record(info) for info in aggregated.values()]
        final_results = sorted(final_results, key=_BY_SIMILARITY, reverse=True)
        all_initial_results = final_results.copy()
        final_results = final_results[:top_n]
        if progress_callback:
            progress_callback("✅ 检索完成，准备重排序...", 55)

        retrieval_details = None
        if traversal_details:
            retrieval_details = {"method": retrieval_method,
                                 "traversal_info": traversal_details[0] if len(traversal_details) == 1 else traversal_details,
                                 "max_hops": max_hops, "neighbor_k": neighbor_k}
        algorithm_contribution = None
        if retrieval_method == "hybrid_expansion":
            algorithm_contribution = {
                "basic_retrieval_count": algo["basic_count"],
                "ssg_new_chunks_count": len(algo["ssg_new_chunks"]),
                "triangulation_new_chunks_count": len(algo["triangulation_new_chunks"]),
                "ssg_new_chunks": algo["ssg_new_chunks"],
                "triangulation_new_chunks": algo["triangulation_new_chunks"],
                "ssg_stats": algo["ssg_stats"], "triangulation_stats": algo["triangulation_stats"]}
        return {"results": final_results, "timing": plan["timing"], "expansion_texts": plan["expansion_texts"],
                "retrieval_details": retrieval_details, "initial_retrieval_results": all_initial_results,
                "algorithm_contribution": algorithm_contribution}

    def _direct_records(self, per_report):
        """Records of a question whose hits need no cross-hit aggregation — one query text, basic searches only,
        every report under its own sha1 (keys of different reports cannot collide) — built straight from the result
        rows in aggregation order (reports, then ranks).  None when the general path is needed."""
        raws = []
        for per_query in per_report:
            if per_query is None:
                continue
            if len(per_query) != 1 or not isinstance(per_query[0][1], _RawHits) or not per_query[0][1].direct_ok():
                return None
            raws.append(per_query[0])
        if not raws or len({q for q, _ in raws}) != 1 or len({r.sha1 for _, r in raws}) != len(raws):
            return None
        out = []
        for query_text, raw in raws:
            out.extend(raw.records(query_text, self._safe_print))
        return out

    # ---- aggregation and score fusion (reference :1552-1599, :1639-1716) -----------------------------------
    @staticmethod
    def _aggregate(aggregated, hits, query_text):
        for hit in hits:
            if len(hit) == 5:
                key, page_id, text, similarity, sha1 = hit
                source = "basic"
            elif len(hit) == 6:
                key, page_id, text, similarity, sha1, source = hit
            else:
                continue
            info = aggregated.get(key)
            if info is None:
                aggregated[key] = {"page": page_id, "text": text, "similarities": [similarity],
                                   "query_sources": [query_text], "retrieval_sources": [source], "count": 1,
                                   "source_sha1": sha1}
                continue
            info["similarities"].append(similarity)
            if source not in info["retrieval_sources"]:
                info["retrieval_sources"].append(source)
            if query_text not in info["query_sources"]:  # one count per distinct query (:1597-1599)
                info["count"] += 1
                info["query_sources"].append(query_text)

    @staticmethod
    def fuse_similarity(info):
        """final = max(similarities) * (1 + 0.2*(count-1)) * (1 + 0.1*(methods-1) if methods >= 2 else 1)
        (reference :1639-1661; worked examples in docs/retrieval_algorithms/...20251203.md:139-151)."""
        sims = info["similarities"]
        base = sims[0] if len(sims) == 1 else max(sims)
        query_bonus = 1.0 + 0.2 * (info["count"] - 1)
        methods = set(info.get("retrieval_sources") or ["basic"])
        method_bonus = 1.0 + 0.1 * (len(methods) - 1) if len(methods) >= 2 else 1.0
        return base * query_bonus * method_bonus, base, methods

    @classmethod
    def _final_record(cls, info):
        sims, sources = info["similarities"], info.get("retrieval_sources")
        if len(sims) == 1 and info["count"] == 1 and sources and len(sources) == 1:
            # the common record (one query, one method hit it once): both bonuses are exactly 1.0
            r = round(sims[0], 4)
            return {"vector_similarity": r, "max_original_similarity": r, "page": info["page"], "text": info["text"],
                    "hit_count": 1, "retrieval_sources": [sources[0]], "source_sha1": info["source_sha1"],
                    "query_sources": info.get("query_sources", [])}
        final, base, methods = cls.fuse_similarity(info)
        return {"vector_similarity": round(final, 4), "max_original_similarity": round(base, 4),
                "page": info["page"], "text": info["text"], "hit_count": info["count"],
                "retrieval_sources": list(methods), "source_sha1": info["source_sha1"],
                "query_sources": info.get("query_sources", [])}

    def retrieve_all(self, company_name: str) -> List[Dict]:
        """Every page of every report of the company, constant score 0.5 (reference :1821-1858)."""
        matching = [r for r in self.all_dbs
                    if r.get("document", {}).get("metainfo")
                    and r["document"]["metainfo"].get("company_name") == company_name]
        if not matching:
            _log.error(f"No report found with '{company_name}' company name.")
            raise ValueError(f"No report found with '{company_name}' company name.")
        pages_out = []
        for report in matching:
            document = report["document"]
            sha1 = document["metainfo"]["sha1_name"]
            for page in sorted(document["content"]["pages"], key=lambda p: p["page"]):
                pages_out.append({"vector_similarity": 0.5, "page": page["page"], "text": page["text"],
                                  "source_sha1": sha1})
        return pages_out


# --------------------------------------------------------------------------------------------------
# HybridRetriever (reference :297-429): vector retrieval + (LLM) rerank
# --------------------------------------------------------------------------------------------------
class HybridRetriever:
    def __init__(self, vector_db_dir: Path, documents_dir: Path, use_hyde: bool = True, use_multi_query: bool = True,
                 subset_path: Path = None, parallel_workers: int = 4,
                 multi_query_methods: Optional[Dict[str, bool]] = None, retrieval_method: str = "basic",
                 max_hops: int = 4, neighbor_k: int = 30, qwen=None, llm=None, reranker=None, device: int = 0):
        self.vector_retriever = VectorRetriever(
            vector_db_dir, documents_dir, use_hyde=use_hyde, use_multi_query=use_multi_query, subset_path=subset_path,
            parallel_workers=parallel_workers, multi_query_methods=multi_query_methods,
            retrieval_method=retrieval_method, max_hops=max_hops, neighbor_k=neighbor_k, qwen=qwen, llm=llm,
            device=device)
        self.reranker = reranker if reranker is not None else _default_reranker()

    def retrieve_by_company_name(self, company_name: str, query: str, llm_reranking_sample_size: int = 28,
                                 documents_batch_size: int = 2, top_n: int = 6, llm_weight: float = 0.7,
                                 return_parent_pages: bool = False, use_hyde: bool = None, use_multi_query: bool = None,
                                 progress_callback=None, selected_years: List[int] = None,
                                 multi_query_config: Optional[Dict[str, bool]] = None, retrieval_method: str = "basic",
                                 max_hops: int = 4, neighbor_k: int = 30):
        vec = self.vector_retriever.retrieve_by_company_name(
            company_name=company_name, query=query, top_n=llm_reranking_sample_size,
            return_parent_pages=return_parent_pages, use_hyde=use_hyde, use_multi_query=use_multi_query,
            progress_callback=progress_callback, selected_years=selected_years, multi_query_config=multi_query_config,
            retrieval_method=retrieval_method, max_hops=max_hops, neighbor_k=neighbor_k)
        return self._rerank(query, vec, documents_batch_size, top_n, llm_weight, progress_callback)

    def retrieve_batch(self, company_name: str, queries: List[str], llm_reranking_sample_size: int = 28,
                       documents_batch_size: int = 2, top_n: int = 6, llm_weight: float = 0.7,
                       return_parent_pages: bool = False, use_hyde: bool = None, use_multi_query: bool = None,
                       selected_years=None, multi_query_config: Optional[Dict[str, bool]] = None,
                       retrieval_method: str = "basic", max_hops: int = 4, neighbor_k: int = 30) -> List[Dict]:
        """[retrieve_by_company_name(q) for q in queries]: the vector stage of the whole question list runs as
        VectorRetriever.retrieve_batch (one batched search per report), the reranker then scores each question's sample."""
        vecs = self.vector_retriever.retrieve_batch(
            company_name, queries, top_n=llm_reranking_sample_size, return_parent_pages=return_parent_pages,
            use_hyde=use_hyde, use_multi_query=use_multi_query, selected_years=selected_years,
            multi_query_config=multi_query_config, retrieval_method=retrieval_method, max_hops=max_hops,
            neighbor_k=neighbor_k)
        return [self._rerank(q, vec, documents_batch_size, top_n, llm_weight, None) for q, vec in zip(queries, vecs)]

    def _rerank(self, query, vec, documents_batch_size, top_n, llm_weight, progress_callback):
        """Reference :387-429 — rerank the vector stage's sample, keep the first top_n."""
        timing_info = {"hyde_expansion": 0.0, "multi_query_expansion": 0.0, "vector_search": 0.0, "llm_reranking": 0.0}
        expansion_texts, algorithm_contribution = {}, None
        if isinstance(vec, dict) and "timing" in vec:
            timing_info.update(vec["timing"])
            vector_results = vec["results"]
            expansion_texts = vec.get("expansion_texts", {})
            algorithm_contribution = vec.get("algorithm_contribution")
        else:
            vector_results = vec
        if progress_callback:
            progress_callback("🎯 LLM 重排序中（这可能需要一些时间）...", 58)
        t0 = time.time()
        reranked = self.reranker.rerank_documents(query=query, documents=vector_results,
                                                  documents_batch_size=documents_batch_size, llm_weight=llm_weight)
        timing_info["llm_reranking"] = time.time() - t0
        return {"results": reranked[:top_n], "initial_retrieval_results": vector_results, "timing": timing_info,
                "expansion_texts": expansion_texts, "reranker_stats": self.reranker.get_stats(),
                "algorithm_contribution": algorithm_contribution}
