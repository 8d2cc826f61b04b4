"""CPU: seeded random sweep of call parameters through the REAL reference retrievers and this repo's, side by
side (tests/golden/differential.py) — identical dictionaries, including error behaviour.  Needs the reference
checkout (authoring container); skipped elsewhere.  Subprocess: the run replaces modules in sys.modules."""
import json
import os
import subprocess
import sys

import pytest

from tests.golden import differential as df

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(df.REF, "src")),
                                reason="the reference checkout is not on this machine")


@pytest.mark.parametrize("seed,extra", [(1, []), (2, []), (3, ["--corrupt"])])
def test_random_calls_agree_with_the_reference(seed, extra):
    """--corrupt: a damaged data directory (page records missing, indexes with more / fewer rows than the report has
    chunks) — partial hits, Python's chunks[-1] for faiss's -1 padding, and the reference's UnboundLocalError out of
    a failed hybrid_expansion task must all come out the same."""
    proc = subprocess.run([sys.executable, "-m", "tests.golden.differential", "--cases", "150", "--seed", str(seed)] + extra,
                          cwd=df.ROOT, capture_output=True, text=True, timeout=900)
    report = json.loads(proc.stdout)
    assert report["cases"] == 150
    assert not report["failures"], json.dumps(report["failures"][:3], ensure_ascii=False, indent=1)
    assert proc.returncode == 0
