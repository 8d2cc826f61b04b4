"""Oracle for semantic_similarity_analysis/generate_similarity_matrix.py:84-98 (matrix) and
:162-171 (off-diagonal statistics).  Plain numpy, same operation order as the reference."""
import numpy as np


def cosine_matrix_ref(vectors):
    vectors = np.asarray(vectors, dtype=np.float32)
    norms = np.linalg.norm(vectors, axis=1, keepdims=True)
    normalized = vectors / (norms + 1e-8)
    sim = normalized @ normalized.T
    np.fill_diagonal(sim, 1.0)
    return np.clip(sim, -1.0, 1.0)


def cosine_stats_ref(sim):
    n = sim.shape[0]
    off = sim[~np.eye(n, dtype=bool)]
    return {
        "avg_similarity": float(np.mean(off)),
        "median_similarity": float(np.median(off)),
        "std_similarity": float(np.std(off)),
        "max_similarity": float(np.max(off)),
        "min_similarity": float(np.min(off)),
    }
