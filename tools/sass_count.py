"""Instruction mix of the step loop of a pipelined fill kernel (compiled alone, sm_100a).

usage: python tools/sass_count.py [R] [NC] [WAVE]   -> prints per-step instruction count and opcode histogram
"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
NC = int(sys.argv[2]) if len(sys.argv) > 2 else 4
WAVE = sys.argv[3] if len(sys.argv) > 3 else "false"
wave = WAVE == "true"
kern = os.environ.get("KERNEL", ("viterbi_wave1_kernel<%d, %d>" if wave else "viterbi_pipe1_kernel<%d, %d>") % (R, NC))
hdr = os.environ.get("HEADER", "viterbi_wave1.cuh" if wave else "viterbi_pipe1.cuh")
src = '#include "%s/coati_b200/csrc/%s"\nnamespace coati_gpu { template __global__ void %s(const PairDesc*, uint32_t, uint32_t, unsigned int*, const uint8_t*, const uint8_t*, const float*, GapConsts, float4*, uint32_t, uint8_t*, PairResult*, const unsigned int*); }\n' % (root, hdr, kern)
d = tempfile.mkdtemp()
open(d + "/k.cu", "w").write(src)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false"] + os.environ.get("DEFS", "").split() + ["-cubin", "-Xptxas", "-v", "-o", d + "/k.cubin", d + "/k.cu"])
sass = subprocess.check_output(["cuobjdump", "-sass", d + "/k.cubin"]).decode()
ins = []
for l in sass.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2)))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a: loops.append((int(m.group(1), 16), a))
print("total", len(ins), "loops", [(hex(a), hex(b), (b - a) // 16 + 1) for a, b in loops])
# the step loop: the innermost loop with the most FSETP/FADD2
best = max(loops, key=lambda ab: sum(1 for a, t in ins if ab[0] <= a <= ab[1] and ("FADD2" in t or "FSETP" in t)) / ((ab[1] - ab[0]) // 16 + 1) ** 0.01 if (ab[1]-ab[0]) < 16*600 else 0)
body = [(a, t) for a, t in ins if best[0] <= a <= best[1]]
c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for a, t in body)
print("step loop", hex(best[0]), hex(best[1]), len(body), "instructions")
print(sorted(c.items(), key=lambda kv: -kv[1]))
if os.environ.get("DUMP"):
    for a, t in body: print(hex(a), t)
if os.environ.get("ALL"):
    for ab in loops:
        bd = [(a, t) for a, t in ins if ab[0] <= a <= ab[1]]
        cc = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for a, t in bd)
        print(hex(ab[0]), len(bd), sorted(cc.items(), key=lambda kv: -kv[1])[:14])
