"""Generate tests/golden/cosine_outputs.json by running the REAL
/root/reference/semantic_similarity_analysis/generate_similarity_matrix.py (SimilarityMatrixGenerator,
unmodified) over the miniature corpus of corpus.py and over a copy whose rows are rescaled (norms 0.5 .. 2,
exercising the `+1e-8` normalisation) with one exactly duplicated direction (a value that clips at 1).
`faiss` is replaced by the numpy oracle's index object and `plotly.graph_objects` by a no-op figure; the
normalisation, the matrix product, the diagonal / clip rules and the off-diagonal statistics are the
reference's own code (:66-100, :162-185).  This pins oracle/cosine.py (tests/test_cosine_golden.py).

    python -m tests.golden.make_cosine_golden
"""
import importlib.util
import io
import json
import os
import sys
import tempfile
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

REF_SCRIPT = "/root/reference/semantic_similarity_analysis/generate_similarity_matrix.py"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.golden import corpus, stubs  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "cosine_outputs.json")


def build_scaled_copy(vector_db_dir, dst_dir):
    """Same reports, rows multiplied by per-row factors in [0.5, 2); row 5 becomes 3 x row 2."""
    dst_dir = Path(dst_dir)
    dst_dir.mkdir(parents=True, exist_ok=True)
    for src in sorted(Path(vector_db_dir).glob("*.faiss")):
        rows = stubs.read_ixfi(src).rows.copy()
        rng = np.random.default_rng(len(src.stem) * 1000 + rows.shape[0])
        rows *= rng.uniform(0.5, 2.0, (rows.shape[0], 1)).astype(np.float32)
        if rows.shape[0] > 8:
            rows[5] = 3.0 * rows[2]
        corpus.write_ixfi(dst_dir / src.name, rows)
    return dst_dir


def _install_stubs():
    faiss = types.ModuleType("faiss")
    faiss.read_index = lambda path: stubs.read_ixfi(path)
    sys.modules["faiss"] = faiss

    class _Fig:
        def __init__(self, *a, **kw):
            pass

        def update_layout(self, *a, **kw):
            pass

        def write_html(self, path, **kw):
            with open(path, "w") as f:
                f.write("<html></html>")

    go = types.ModuleType("plotly.graph_objects")
    go.Figure = _Fig
    go.Heatmap = lambda **kw: None
    plotly = types.ModuleType("plotly")
    plotly.graph_objects = go
    sys.modules["plotly"] = plotly
    sys.modules["plotly.graph_objects"] = go


def main():
    _install_stubs()
    spec = importlib.util.spec_from_file_location("ref_generate_similarity_matrix", REF_SCRIPT)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    with tempfile.TemporaryDirectory() as tmp, redirect_stdout(io.StringIO()):
        paths = corpus.build(tmp)
        scaled = build_scaled_copy(paths["vector_db_dir"], Path(tmp) / "vector_dbs_scaled")
        for label, vec_dir in (("unit_rows", paths["vector_db_dir"]), ("scaled_rows", scaled)):
            gen = ref.SimilarityMatrixGenerator(vec_dir, paths["bm25_documents_dir"], Path(tmp) / f"out_{label}")
            out[label] = gen.process_all_documents()
            with open(Path(tmp) / f"out_{label}" / "statistics.json", encoding="utf-8") as f:
                assert json.load(f) == out[label]
    with open(GOLDEN, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, sort_keys=True, indent=1)
    print("wrote", GOLDEN, {k: sorted(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
