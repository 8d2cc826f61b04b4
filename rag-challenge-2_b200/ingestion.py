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


class BM25Ingestor:
    def __init__(self, device: int = 0):
        self.device = int(device)

    def create_bm25_index(self, chunks: List[str]) -> BM25Okapi:
        """src/ingestion.py:19-22 — whitespace tokenisation, one document per chunk."""
        tokenized_chunks = [chunk.split() for chunk in chunks]
        return BM25Okapi(tokenized_chunks, device=self.device)

    def process_reports(self, all_reports_dir: Path, output_dir: Path):
        """src/ingestion.py:24-49 — one `<sha1_name>.pkl` per report JSON."""
        all_reports_dir, output_dir = Path(all_reports_dir), Path(output_dir)
        output_dir.mkdir(parents=True, exist_ok=True)
        all_report_paths = list(all_reports_dir.glob("*.json"))
        for report_path in all_report_paths:
            with open(report_path, "r", encoding="utf-8") as f:
                report_data = json.load(f)
            text_chunks = [chunk["text"] for chunk in report_data["content"]["chunks"]]
            bm25_index = self.create_bm25_index(text_chunks)
            sha1_name = report_data["metainfo"]["sha1_name"]
            with open(output_dir / f"{sha1_name}.pkl", "wb") as f:
                pickle.dump(bm25_index, f)
            bm25_index.close()
        print(f"Processed {len(all_report_paths)} reports")


class VectorDBIngestor:
    BATCH = 100  # texts per embedding request (src/ingestion.py:71-72)

    def __init__(self, qwen=None, device: int = 0, retry_wait: float = 20.0):
        if qwen is None:
            from .retrieval import _default_qwen

            qwen = _default_qwen(1024)
        self.qwen = qwen
        self.device = int(device)
        self.retry_wait = float(retry_wait)

    def _get_embeddings(self, text: Union[str, List[str]], model: str = "text-embedding-v4") -> List[List[float]]:
        """src/ingestion.py:57-76, including its `@retry(wait=wait_fixed(20), stop=stop_after_attempt(2))`: with
        tenacity installed the same decorator is applied (a second failure raises tenacity.RetryError, as in the
        reference); without it the call is retried once and the second exception propagates."""
        try:
            from tenacity import retry, stop_after_attempt, wait_fixed
        except ImportError:
            try:
                return self._get_embeddings_once(text, model)
            except Exception:  # noqa: BLE001
                time.sleep(self.retry_wait)
                return self._get_embeddings_once(text, model)
        return retry(wait=wait_fixed(self.retry_wait), stop=stop_after_attempt(2))(self._get_embeddings_once)(text, model)

    def _get_embeddings_once(self, text, model):
        if isinstance(text, str) and not text.strip():
            raise ValueError("Input text cannot be an empty string.")
        text_chunks = [text] if isinstance(text, str) else text
        embeddings = []
        for i in range(0, len(text_chunks), self.BATCH):
            response = self.qwen.get_embeddings(texts=text_chunks[i: i + self.BATCH], model=model)
            embeddings.extend([item["embedding"] for item in response])
        return embeddings

    def _create_vector_db(self, embeddings: List[List[float]]):
        """src/ingestion.py:78-84 — rows stored verbatim (no normalisation), inner-product metric."""
        embeddings_array = np.array(embeddings, dtype=np.float32)
        dimension = len(embeddings[0])
        index = flat_index.IndexFlatIP(dimension, device=self.device)
        index.add(embeddings_array)
        return index

    def _process_report(self, report: dict):
        text_chunks = [chunk["text"] for chunk in report["content"]["chunks"]]
        return self._create_vector_db(self._get_embeddings(text_chunks))

    def process_reports(self, all_reports_dir: Path, output_dir: Path):
        """src/ingestion.py:92-102 — one `<sha1_name>.faiss` per report JSON."""
        all_reports_dir, output_dir = Path(all_reports_dir), Path(output_dir)
        all_report_paths = list(all_reports_dir.glob("*.json"))
        output_dir.mkdir(parents=True, exist_ok=True)
        for report_path in all_report_paths:
            with open(report_path, "r", encoding="utf-8") as file:
                report_data = json.load(file)
            index = self._process_report(report_data)
            sha1_name = report_data["metainfo"]["sha1_name"]
            flat_index.write_index(index, str(output_dir / f"{sha1_name}.faiss"))
            index.close()
        print(f"Processed {len(all_report_paths)} reports")
