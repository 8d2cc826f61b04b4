"""CPU: the reference's REAL caller (src/questions_processing.py: QuestionsProcessor, unmodified) running on
top of this repo's retriever classes installed as `src.retrieval` must produce what it
produces on top of the reference's own src/retrieval.py — answers, page groups, references, retrieved-chunk
tables, submission records, error records (tests/golden/caller_outputs.json, written by
`python -m tests.golden.caller_conformance --arm reference`).

The caller's code lives in /root/reference, which exists only in the authoring container: the test skips
elsewhere (the GPU box).  It runs in a subprocess because it replaces `src.*` modules in sys.modules."""
import json
import os
import subprocess
import sys

import pytest

from tests.golden import caller_conformance as cc

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(cc.REF, "src")),
                                reason="the reference checkout (the caller's code) is not on this machine")


def _diff(want, got, path=""):
    if type(want) is not type(got):
        return [f"{path}: type {type(want).__name__} != {type(got).__name__}"]
    if isinstance(want, dict):
        out = []
        for k in sorted(set(want) | set(got)):
            if k not in want or k not in got:
                out.append(f"{path}/{k}: only in {'golden' if k in want else 'ours'}")
            else:
                out += _diff(want[k], got[k], f"{path}/{k}")
        return out
    if isinstance(want, list):
        if len(want) != len(got):
            return [f"{path}: length {len(want)} != {len(got)}"]
        out = []
        for i, (a, b) in enumerate(zip(want, got)):
            out += _diff(a, b, f"{path}[{i}]")
        return out
    return [] if want == got else [f"{path}: {want!r:.80} != {got!r:.80}"]


def test_questions_processor_over_our_retrievers(tmp_path):
    out = tmp_path / "ours.json"
    proc = subprocess.run([sys.executable, "-m", "tests.golden.caller_conformance", "--arm", "ours", "--out", str(out)],
                          cwd=cc.ROOT, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    with open(out, encoding="utf-8") as f:
        got = json.load(f)
    with open(cc.GOLDEN, encoding="utf-8") as f:
        want = json.load(f)
    assert sorted(got["answers"]) == sorted(cc.SCENARIOS)
    problems = _diff(want, got)
    assert not problems, "\n".join(problems[:20])
    # the scenarios did exercise the caller's branches
    assert want["answers"]["rerank_upstream_groups"]["selected_groups"]
    assert want["answers"]["hybrid_expansion_method"]["algorithm_contribution"]
    assert want["batch"]["statistics"] == {"total_questions": 5, "error_count": 2, "na_count": 0, "success_count": 3}
    assert want["errors"]["unknown_company"] == ["ValueError", "No report found with 'nobody' company name."]
