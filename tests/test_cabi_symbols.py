"""CPU: libcoati_gpu.so loads and exports every symbol include/coati_gpu.h declares; with no GPU
every compute entry point must FAIL LOUDLY (no CPU fallback exists)."""
import ctypes as C
import os
import re

import pytest

import coati_b200
from coati_b200 import build as cbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    cbuild.build()
    return coati_b200.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "coati_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(coati_gpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/coati_gpu.h but not exported"


def test_strerror_messages(lib):
    # messages the C++ wrapper rethrows with (reference wording: utils.cc:507-514, align_marginal.cc:73)
    assert b"Early stop codon in ancestor/reference." == lib.coati_gpu_strerror(-7)
    assert b"Ambiguous nucleotides in ancestor/reference." == lib.coati_gpu_strerror(-6)
    assert b"exceed available memory" in lib.coati_gpu_strerror(-3)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu suite")
    with pytest.raises(coati_b200.CoatiGpuError) as ei:
        coati_b200.Context(0)
    assert ei.value.code == -1


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU implementation)."""
    pkg = os.path.join(ROOT, "coati_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".hpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "coati_oracle" not in text and "liboracle" not in text, f
