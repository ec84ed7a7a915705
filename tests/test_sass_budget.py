"""CPU (needs nvcc, no GPU): instruction and register budgets of the hot step loops, read from the SASS of each
kernel compiled alone for sm_100a.  The fills are bound by instruction issue, so the number of
instructions one step issues IS their performance model; this keeps a refactoring from silently adding to it.
Budgets = the counts of the committed kernels (tools/sass_count.py) + 3 %."""
import collections
import os
import re
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
SIG = ("(const PairDesc*, uint32_t, uint32_t, unsigned int*, const uint8_t*, const uint8_t*, const float*, GapConsts, "
       "float4*, uint32_t, uint8_t*, PairResult*, const unsigned int*)")

pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not found")


def _compile(header, kernel):
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "k.cu"), "w") as f:
        f.write('#include "%s/coati_b200/csrc/%s"\nnamespace coati_gpu { template __global__ void %s%s; }\n'
                % (ROOT, header, kernel, SIG))
    r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-cubin",
                        "-Xptxas", "-v", "-o", os.path.join(d, "k.cubin"), os.path.join(d, "k.cu")],
                       capture_output=True, text=True, check=True)
    regs = int(re.search(r"Used (\d+) registers", r.stderr).group(1))
    sass = subprocess.check_output(["cuobjdump", "-sass", os.path.join(d, "k.cubin")]).decode()
    ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", sass)]
    loops = set()
    for a, t in ins:
        m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.add((int(m.group(1), 16), a))
    out = []
    for lo, hi in loops:
        body = [t for a, t in ins if lo <= a <= hi]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
        out.append(dict(n=len(body), ops=ops))
    shutil.rmtree(d, ignore_errors=True)
    return regs, out


def _step_loop(loops, fadd2_per_step, steps):
    """The innermost loop holding exactly `steps` steps' worth of packed adds."""
    cands = [l for l in loops if l["ops"]["FADD2"] == fadd2_per_step * steps]
    assert cands, sorted((l["n"], l["ops"]["FADD2"]) for l in loops)
    return min(cands, key=lambda l: l["n"])


# (header, kernel, FADD2 per step, steps per interior loop iteration, max instructions per step, max registers)
CASES = [
    ("viterbi_pipe1.cuh", "viterbi_pipe1_kernel<10, 4>", 65, 4, 204.0, 128),   # 198 at the 128-register cap
    ("viterbi_pipe1.cuh", "viterbi_pipe1_kernel<8, 4>", 52, 4, 166.0, 128),
    ("viterbi_pipe1.cuh", "viterbi_pipe1_kernel<4, 4>", 26, 4, 88.0, 128),     # 85.5
    ("viterbi_pipe3.cuh", "viterbi_pipe3_kernel<6, 4>", 57, 3, 155.0, 128),    # 150
    ("viterbi_wave1.cuh", "viterbi_wave1_kernel<4, 4>", 26, 8, 94.0, 255),     # 91.25: eight-step groups
    ("viterbi_wave1.cuh", "viterbi_wave1_kernel<10, 4>", 65, 4, 216.0, 255),
]


@pytest.mark.parametrize("header,kernel,f2,steps,max_per_step,max_regs", CASES, ids=[c[1] for c in CASES])
def test_interior_step_budget(header, kernel, f2, steps, max_per_step, max_regs):
    regs, loops = _compile(header, kernel)
    assert regs <= max_regs, "register budget: occupancy of the inter-pair fills depends on 128 (four CTAs per SM)"
    loop = _step_loop(loops, f2, steps)
    per_step = loop["n"] / steps
    assert per_step <= max_per_step, (per_step, dict(loop["ops"]))
    assert loop["ops"]["LDL"] == 0 and loop["ops"]["STL"] == 0, "spill inside the interior step loop"
    # the decisions are sign bits pushed by funnel shifts: five per cell (rows per lane = FADD2 per step / 6.5)
    cells = round(f2 / 6.5) if "pipe3" not in kernel else 6
    assert loop["ops"]["SHF"] >= 5 * cells * steps


def test_rows_to_host_kernel_writes_16_byte_words():
    """rows_to_host_kernel (traceback.cuh) is bound by the host link, which wants large writes: the bulk of a row
    must leave as 16-byte stores (STG.128), the kernel must stay small enough to sit beside the fills (<= 32
    registers, no shared memory, no spill)."""
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "k.cu"), "w") as f:
        f.write('#include "%s/coati_b200/csrc/traceback.cuh"\n' % ROOT)
    r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-cubin",
                        "-Xptxas", "-v", "-o", os.path.join(d, "k.cubin"), os.path.join(d, "k.cu")],
                       capture_output=True, text=True, check=True)
    info = r.stderr[r.stderr.index("rows_to_host_kernel"):]
    info = info[:info.index("Compile time")]
    assert int(re.search(r"Used (\d+) registers", info).group(1)) <= 32
    assert "0 bytes spill stores" in info and "smem" not in info
    sass = subprocess.check_output(["cuobjdump", "-sass", os.path.join(d, "k.cubin")]).decode()
    shutil.rmtree(d, ignore_errors=True)
    sass = sass[sass.index("rows_to_host_kernel"):]
    sass = sass[:sass.index("Function :", 20)] if "Function :" in sass[20:] else sass
    assert re.search(r"STG\.E\.128", sass), "rows leave as 16-byte words"
    assert len(re.findall(r"LDG\.E\.U8", sass)) >= 16, "sixteen independent byte loads per word"
