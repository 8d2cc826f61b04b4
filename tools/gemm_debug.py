"""Bisect helper: one tensor-core search with the given tunables, meant to run under `timeout`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rag_challenge_2_b200 as pkg
from rag_challenge_2_b200 import _lib
import oracle
n, nq = int(sys.argv[1]), int(sys.argv[2])
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    _lib.tunable_set(k, int(v))
rng = np.random.default_rng(0)
xb = rng.standard_normal((n, 1024), dtype=np.float32); xb /= np.linalg.norm(xb, axis=1, keepdims=True)
xq = xb[:nq] + 0.01 * rng.standard_normal((nq, 1024), dtype=np.float32)
idx = pkg.FlatIPIndex(1024); idx.add(xb)
print("searching", sys.argv[1:], flush=True)
D, I = idx.search(xq, 30)
oracle.assert_topk_equivalent(D, I, oracle.flat_ip_scores_ref(xb, xq, np.float64), 30)
print("OK", I[:3, :3].tolist(), flush=True)
