"""Index builders with the reference's names and call shapes (src/ingestion.py:15-103): `BM25Ingestor` and
`VectorDBIngestor`, the classes `Pipeline.create_bm25_db / create_vector_dbs` drive (src/pipeline.py:186-202).
They read `chunked_reports/*.json`, build one index per report and write it under the report's `sha1_name`:

  BM25Ingestor      BM25Okapi([chunk.split() ...]) -> pickle `<sha1>.pkl`            (src/ingestion.py:19-47)
  VectorDBIngestor  embeddings in batches of 100 -> float32 array -> IndexFlatIP(d).add -> `<sha1>.faiss`
                                                                                        (src/ingestion.py:58-102)

The index objects are this package's (postings and rows resident on the GPU); the files are the formats the
retrievers load (`load_bm25_pickle`, `flat_index.read_index`).  The embedding client is the same injectable seam
as in retrieval.py (offline: pass `qwen=`).
"""
import json
import pickle
import time
from pathlib import Path
from typing import List, Union

import numpy as np

from . import flat_index
from .bm25 import BM25Okapi


def _report_chunks(reports_dir):
    """(sha1_name, [chunk text, ...]) for every `*.json` under reports_dir, in directory order — the loop both
    builders of the reference run (src/ingestion.py:32-44, :93-100)."""
    for path in list(Path(reports_dir).glob("*.json")):
        with open(path, "r", encoding="utf-8") as f:
            report = json.load(f)
        yield report["metainfo"]["sha1_name"], [c["text"] for c in report["content"]["chunks"]]


class BM25Ingestor:
    """`BM25Ingestor().process_reports(chunked_reports_dir, bm25_dbs_dir)`"""

    def __init__(self, device: int = 0):
        self.device = int(device)

    def create_bm25_index(self, chunks: List[str]) -> BM25Okapi:
        """One document per chunk, whitespace tokens (src/ingestion.py:19-22); postings land on `device`."""
        return BM25Okapi([text.split() for text in chunks], device=self.device)

    def process_reports(self, all_reports_dir: Path, output_dir: Path):
        """`<output_dir>/<sha1_name>.pkl` per report (src/ingestion.py:24-49)."""
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        n = 0
        for sha1_name, texts in _report_chunks(all_reports_dir):
            index = self.create_bm25_index(texts)
            with open(out / f"{sha1_name}.pkl", "wb") as f:
                pickle.dump(index, f)
            index.close()   # the pickle carries the statistics; the device copy is rebuilt on load
            n += 1
        print(f"Processed {n} reports")


class VectorDBIngestor:
    """`VectorDBIngestor().process_reports(chunked_reports_dir, vector_dbs_dir)`"""

    BATCH = 100  # texts per embedding request (src/ingestion.py:71-72)

    def __init__(self, qwen=None, device: int = 0, retry_wait: float = 20.0):
        if qwen is None:
            from .retrieval import _default_qwen

            qwen = _default_qwen(1024)
        self.qwen, self.device, self.retry_wait = qwen, int(device), float(retry_wait)

    def _request(self, text, model):
        if isinstance(text, str):
            if not text.strip():
                raise ValueError("Input text cannot be an empty string.")
            text = [text]
        vectors = []
        for lo in range(0, len(text), self.BATCH):
            reply = self.qwen.get_embeddings(texts=text[lo: lo + self.BATCH], model=model)
            vectors += [item["embedding"] for item in reply]
        return vectors

    def _get_embeddings(self, text: Union[str, List[str]], model: str = "text-embedding-v4") -> List[List[float]]:
        """src/ingestion.py:57-76, including its `@retry(wait=wait_fixed(20), stop=stop_after_attempt(2))`: with
        tenacity installed the same policy is applied (a second failure raises tenacity.RetryError, as in the
        reference); without it the request is repeated once and the second exception propagates."""
        try:
            import tenacity
        except ImportError:
            try:
                return self._request(text, model)
            except Exception:  # noqa: BLE001
                time.sleep(self.retry_wait)
                return self._request(text, model)
        policy = tenacity.retry(wait=tenacity.wait_fixed(self.retry_wait), stop=tenacity.stop_after_attempt(2))
        return policy(self._request)(text, model)

    def _create_vector_db(self, embeddings: List[List[float]]):
        """Rows stored verbatim, inner-product metric (src/ingestion.py:78-84)."""
        rows = np.array(embeddings, dtype=np.float32)
        index = flat_index.IndexFlatIP(len(embeddings[0]), device=self.device)
        index.add(rows)
        return index

    def _process_report(self, report: dict):
        return self._create_vector_db(self._get_embeddings([c["text"] for c in report["content"]["chunks"]]))

    def process_reports(self, all_reports_dir: Path, output_dir: Path):
        """`<output_dir>/<sha1_name>.faiss` per report (src/ingestion.py:92-102)."""
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        n = 0
        for sha1_name, texts in _report_chunks(all_reports_dir):
            index = self._create_vector_db(self._get_embeddings(texts))
            flat_index.write_index(index, str(out / f"{sha1_name}.faiss"))
            index.close()
            n += 1
        print(f"Processed {n} reports")
