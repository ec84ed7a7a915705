"""Lone-warp schedule of a fill kernel's step loop, from the control words of its SASS (no GPU needed).

The intra-pair wavefront runs about one warp per scheduler, so its speed is the length of one step as a single
warp issues it: the stall fields of the instructions plus the scoreboard waits on variable-latency results
(shuffles, shared and global loads).  This tool compiles one kernel instance alone, decodes bits 105-121 of each
instruction (stall, yield, write/read barrier, wait mask -- the single-warp issue model of
/opt/skills/guides/B300_MICROARCH.md) and replays the step loop for a few iterations.

usage: python tools/sass_sched.py [R] [NC] [WAVE]        (env KERNEL / HEADER / DUMP=1 / DEFS="-DX=1 ..." as sass_count.py)
prints: instructions per step, sum of stall fields, modelled cycles per step in steady state, and where the
        cycles beyond the stall sum are waited (which producer class).
"""
import collections, os, re, subprocess, sys, tempfile

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
NC = int(sys.argv[2]) if len(sys.argv) > 2 else 4
WAVE = sys.argv[3] if len(sys.argv) > 3 else "true"
wave = WAVE == "true"
kern = os.environ.get("KERNEL", ("viterbi_wave1_kernel<%d, %d>" if wave else "viterbi_pipe1_kernel<%d, %d>") % (R, NC))
hdr = os.environ.get("HEADER", "viterbi_wave1.cuh" if wave else "viterbi_pipe1.cuh")
src = ('#include "%s/coati_b200/csrc/%s"\nnamespace coati_gpu { template __global__ void %s(const PairDesc*, uint32_t, '
       'uint32_t, unsigned int*, const uint8_t*, const uint8_t*, const float*, GapConsts, float4*, uint32_t, uint8_t*, '
       'PairResult*, const unsigned int*); }\n' % (root, hdr, kern))
d = tempfile.mkdtemp()
open(d + "/k.cu", "w").write(src)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false"]
                      + os.environ.get("DEFS", "").split() + ["-cubin", "-o", d + "/k.cubin", d + "/k.cu"])
sass = subprocess.check_output(["cuobjdump", "-sass", d + "/k.cubin"]).decode().split("\n")

# variable-latency classes: cycles from issue until the write barrier clears (B300_MICROARCH.md; SHFL measured
# ~23 on recent parts; global loads here are L1-resident broadcast lines or L2 granules fetched ahead)
LAT = {"SHFL": 23, "LDS": 29, "LDG": 34, "LD": 34, "LDC": 20, "LDCU": 20, "S2R": 20, "MUFU": 18, "VOTE": 10, "R2UR": 8,
       "NANOSLEEP": 40, "ATOMG": 320, "CS2R": 8, "POPC": 10, "REDUX": 20, "MATCH": 20, "BAR": 20}
RLAT = 6   # read barrier (operands consumed) of stores / shuffles

ins = []
for i, l in enumerate(sass):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if m:
        m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", sass[i + 1])
        hi = int(m2.group(1), 16)
        ctl = hi >> 41
        ins.append(dict(addr=int(m.group(1), 16), text=m.group(2), stall=ctl & 15, yld=(ctl >> 4) & 1,
                        wbar=(ctl >> 5) & 7, rbar=(ctl >> 8) & 7, wait=(ctl >> 11) & 63))
loops = []
for x in ins:
    m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", x["text"])
    if m and int(m.group(1), 16) < x["addr"]:
        loops.append((int(m.group(1), 16), x["addr"]))


def weight(ab):
    n = (ab[1] - ab[0]) // 16 + 1
    return sum(1 for x in ins if ab[0] <= x["addr"] <= ab[1] and ("FADD2" in x["text"] or "FSETP" in x["text"])) if n < 3000 else 0


cands = sorted((ab for ab in set(loops) if weight(ab) > 0), key=lambda ab: ab[1] - ab[0])
print("loops with cell updates (LOOP=i picks one):", [(i, hex(a), (b - a) // 16 + 1, weight((a, b))) for i, (a, b) in enumerate(cands)])
if os.environ.get("LOOP"):
    best = cands[int(os.environ["LOOP"])]
else:  # the hot inner loop: the smallest one holding at least half of the cell updates of the largest
    wmax = max(weight(ab) for ab in cands)
    best = next(ab for ab in cands if weight(ab) * 2 >= wmax)
body = [x for x in ins if best[0] <= x["addr"] <= best[1]]


def opclass(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]


def replay(body, iters=6):
    T, sb, owner = 0, [0] * 6, [None] * 6
    waited = collections.Counter()
    t_iter = []
    for it in range(iters):
        t0 = T
        for x in body:
            op = opclass(x["text"])
            arm, who = 0, None
            for s in range(6):
                if x["wait"] >> s & 1 and sb[s] > arm:
                    arm, who = sb[s], owner[s]
            t_ready = T + x["_prev_stall"]
            if arm > t_ready and it == iters - 1:
                waited[who] += arm - t_ready
            T = max(t_ready, arm)
            x["_t"] = T - t0
            if x["wbar"] < 6:
                sb[x["wbar"]] = max(sb[x["wbar"]], T + LAT.get(op, 30))
                owner[x["wbar"]] = op
            if x["rbar"] < 6:
                sb[x["rbar"]] = max(sb[x["rbar"]], T + RLAT)
                owner[x["rbar"]] = op + "(read)"
        t_iter.append(T - t0)
    return t_iter, waited


# the stall field of an instruction delays the NEXT issue
prev = 0
for x in body:
    x["_prev_stall"] = prev
    prev = max(1, x["stall"])
body[0]["_prev_stall"] = max(1, body[-1]["stall"])
t_iter, waited = replay(body)
c = collections.Counter(opclass(x["text"]) for x in body)
print("kernel", kern)
print("step loop %#x..%#x: %d instructions, stall sum %d, modelled lone-warp cycles/step %s" %
      (best[0], best[1], len(body), sum(max(1, x["stall"]) for x in body), t_iter[-3:]))
print("scoreboard waits beyond the stall fields (last iteration):", dict(waited))
print(sorted(c.items(), key=lambda kv: -kv[1]))
if os.environ.get("DUMP"):
    for x in body:
        print("%5x t=%4d st=%2d y=%d w=%d r=%d wm=%02x  %s" % (x["addr"], x["_t"], x["stall"], x["yld"], x["wbar"], x["rbar"], x["wait"], x["text"]))
