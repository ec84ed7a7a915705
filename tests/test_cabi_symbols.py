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


def test_header_is_plain_c_and_a_cxx_caller_links(lib, tmp_path):
    """The boundary is a C ABI: include/coati_gpu.h compiles as C99 on its own (plain pointers and sizes, no C++ or
    torch types), and the multi-GPU call of INTEGRATION.md -- page-locked arenas from coati_gpu_host_alloc, one
    context per device, coati_gpu_multi_alignpair_batch, the transfer counters -- compiles as C++17 and links against
    the library (not run here: there is no GPU; with one it would fail at coati_gpu_init, loudly)."""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    c_src = tmp_path / "hdr.c"
    c_src.write_text('#include "coati_gpu.h"\nint main(void) { return COATI_GPU_OK; }\n')
    subprocess.run([shutil.which("gcc") or "gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-fsyntax-only",
                    str(c_src)], check=True)
    cxx = tmp_path / "caller.cc"
    cxx.write_text(r'''
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "coati_gpu.h"
int main(int argc, char**) {
    const int n_gpus = argc;                                   // (1 when run without arguments)
    std::vector<coati_gpu_ctx*> ctxs(n_gpus, nullptr);
    std::vector<float> table(183 * 15, 0.0f);
    for(int d = 0; d < n_gpus; ++d) {
        if(int rc = coati_gpu_init(d, &ctxs[d])) { std::fprintf(stderr, "%s\n", coati_gpu_strerror(rc)); return 1; }
        coati_gpu_set_model(ctxs[d], table.data(), 0.001f, 1.0f - 1.0f / 6.0f, 1);
    }
    const char anc[] = "CTCTGGATAGTG", des[] = "CTATAGTG";
    const uint64_t a_off[2] = {0, 12}, b_off[2] = {0, 8};
    const size_t total = 12 + 8 + 1;
    char* out_a = static_cast<char*>(coati_gpu_host_alloc(total + 1));
    char* out_b = static_cast<char*>(coati_gpu_host_alloc(total + 1));
    uint64_t len = 0, h2d = 0, d2h = 0;
    float score = 0;
    int32_t status = 0;
    int rc = coati_gpu_multi_alignpair_batch(ctxs.data(), n_gpus, 1, anc, a_off, des, b_off, out_a, out_b, &len, &score,
                                             &status);
    coati_gpu_transfer_bytes(ctxs[0], &h2d, &d2h);
    std::printf("%d %s %s %g %llu %llu\n", rc, out_a, out_b, score, (unsigned long long)h2d, (unsigned long long)d2h);
    coati_gpu_host_free(out_a), coati_gpu_host_free(out_b);
    for(coati_gpu_ctx* c : ctxs) coati_gpu_shutdown(c);
    return rc;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(coati_b200.library_path())
    subprocess.run([shutil.which("g++") or "g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, str(cxx), "-o", str(exe),
                    "-L" + libdir, "-lcoati_gpu", "-Wl,-rpath," + libdir], check=True)
    import torch
    if not torch.cuda.is_available():       # no GPU here: coati_gpu_init says so
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
