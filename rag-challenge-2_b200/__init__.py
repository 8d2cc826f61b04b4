"""rag-challenge-2_b200 — B200-native retrieval hot path of HEIcby/RAG-Challenge-2.

Exact inner-product top-k (faiss.IndexFlatIP's role), BM25Okapi scoring, top-k merge and the N x N
cosine analysis run in hand-written sm_100a CUDA kernels behind a C ABI (include/b2r.h, libb2r.so);
the Python classes here mirror the reference's retriever interface (src/retrieval.py).
"""
from . import _lib  # noqa: F401
from .flat_index import FlatIPIndex, IndexFlatIP, read_index, write_index, read_index_rows  # noqa: F401

__version__ = "0.1.0"
