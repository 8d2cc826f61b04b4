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


@pytest.mark.parametrize("seed", [1, 2])
def test_random_calls_agree_with_the_reference(seed):
    proc = subprocess.run([sys.executable, "-m", "tests.golden.differential", "--cases", "150", "--seed", str(seed)],
                          cwd=df.ROOT, capture_output=True, text=True, timeout=900)
    report = json.loads(proc.stdout)
    assert report["cases"] == 150
    assert not report["failures"], json.dumps(report["failures"][:3], ensure_ascii=False, indent=1)
    assert proc.returncode == 0
