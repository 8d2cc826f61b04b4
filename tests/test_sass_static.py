"""CPU: static proof that the built library is sm_100a-native — `cuobjdump -sass` of libb2r.so, per kernel:
the scan and BM25 streaming kernels move their data with the bulk-copy engine (UBLKCP = cp.async.bulk), the batched
kernels load operands with TMA tensor copies (UTMALDG), multiply on the 5th-generation tensor cores (UTCHMMA =
tcgen05.mma) and read accumulators from tensor memory (LDTM = tcgen05.ld); the CTA-pair kernels carry the .2CTA forms.
SASS mnemonics as listed in the B200 profiling recipe.  Needs cuobjdump (CUDA toolkit), which the image has."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rag-challenge-2_b200", "libb2r.so")
PAT = re.compile(r"\b(UBLKCP|UTMALDG|UTCHMMA|LDTM|UTCBAR)\b[A-Z0-9_.]*")


@pytest.fixture(scope="module")
def kernels():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not found")
    assert os.path.exists(LIB), "run ./build.sh first"
    elf = subprocess.run([exe, "-lelf", LIB], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), "the library must hold sm_100a code only"
    sass = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        mangled, body = part.split("\n", 1)
        counts = collections.Counter()
        for m in PAT.finditer(body):
            op = m.group(0)
            counts[op.split(".")[0]] += 1
            if "2CTA" in op:
                counts[op.split(".")[0] + ".2CTA"] += 1
            if "MULTICAST" in op:
                counts["MULTICAST"] += 1
        out[mangled.strip()] = counts
    return out


def family(kernels, needle):
    got = {k: v for k, v in kernels.items() if needle in k}
    assert got, f"no kernel named *{needle}* in the library"
    return got


def test_scan_and_bm25_stream_with_the_bulk_copy_engine(kernels):
    for name, ops in {**family(kernels, "flat_scan_tma_kernel"), **family(kernels, "bm25_stream_kernel")}.items():
        assert ops["UBLKCP"] >= 1, name
        assert ops["UTCHMMA"] == 0, name          # HBM-bound integer / fp32 work: not reshaped into tensor-core GEMMs


def test_batched_kernels_use_tma_tcgen05_and_tensor_memory(kernels):
    fams = {}
    for needle in ("flat_gemm_tc_kernel", "flat_gemm_thr_kernel", "flat_gemm_pair_kernel", "cosine_pair_kernel"):
        fams.update(family(kernels, needle))
    assert len(fams) >= 20
    for name, ops in fams.items():
        assert ops["UTMALDG"] >= 1 and ops["UTCHMMA"] >= 1 and ops["LDTM"] >= 1 and ops["UTCBAR"] >= 1, (name, dict(ops))


def test_cta_pair_kernels_issue_two_sm_instructions(kernels):
    pair = {**family(kernels, "flat_gemm_pair_kernel"), **family(kernels, "cosine_pair_kernel")}
    for name, ops in pair.items():
        assert ops["UTCHMMA.2CTA"] == ops["UTCHMMA"] >= 4, (name, dict(ops))     # every MMA is cta_group::2
        assert ops["UTMALDG.2CTA"] >= 2 and ops["UTCBAR.2CTA"] >= 1 and ops["MULTICAST"] >= 1, (name, dict(ops))
    single = {**family(kernels, "flat_gemm_tc_kernel"), **family(kernels, "flat_gemm_thr_kernel")}
    assert all(ops["UTCHMMA.2CTA"] == 0 for ops in single.values())
