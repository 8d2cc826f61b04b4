"""oracle/ — CPU restatement of the reference's retrieval arithmetic.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package, and only as the checker or the timed CPU baseline — never as the shipped path.  The
product (rag-challenge-2_b200/) never imports it and has no CPU fallback.

PARITY UNPINNED beyond the packages' published vectors: the reference's arithmetic lives in
faiss-cpu==1.9.0.post1 and rank-bm25==0.2.2 (reference requirements.txt:8,13).  Neither is vendored in
/root/reference nor installable here (no network), and the reference ships no tests, golden vectors or
prebuilt indexes.  The functions below restate the published algorithms of those two packages and are
pinned by the only numeric outputs the packages themselves publish — rank_bm25's README example and the
faiss wiki's "Getting started" ids / distances (tests/test_oracle_kat.py) — plus hand-derived known-answer
tests, independent sklearn / scipy formulations and the worked examples in the reference's docs; they
have never been compared with the packages running.  The reference's OWN Python logic around that arithmetic
(routing, hit aggregation, score fusion, record building: src/retrieval.py) IS pinned: the
fixtures in tests/golden/ were produced by importing the real src/retrieval.py with faiss /
rank_bm25 replaced by this oracle (tests/golden/make_golden.py).
"""
from .flat_ip import flat_ip_search_ref, flat_ip_scores_ref, assert_topk_equivalent  # noqa: F401
from .bm25 import BM25OkapiRef, bm25_topn_ref  # noqa: F401
from .fusion import merge_lists_ref, hybrid_fuse_ref  # noqa: F401
from .cosine import cosine_matrix_ref, cosine_stats_ref  # noqa: F401
