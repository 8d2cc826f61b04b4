"""The C ABI from plain C: tests/abi_c/consumer.c includes include/b2r.h as C99 (-Wall -Werror), links libb2r.so
and runs.  Without a GPU (CPU test) the program checks that the create calls fail loudly; on a B200 (GPU test) it
builds indexes, searches, and compares with brute force / the Okapi formula evaluated in C."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rag-challenge-2_b200")


def build_consumer(tmp_path):
    exe = str(tmp_path / "consumer")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_c", "consumer.c"), "-o", exe, "-L", PKG, "-lb2r", "-lm",
           f"-Wl,-rpath,{PKG}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return exe


def test_c_consumer_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch

    exe = build_consumer(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "done (no GPU)" in proc.stdout


@pytest.mark.gpu
def test_c_consumer_on_gpu(tmp_path):
    exe = build_consumer(tmp_path)
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "done (GPU)" in proc.stdout
