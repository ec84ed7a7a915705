"""One traced call of coati_gpu_alignpair_batch on the C5 workload (COATI_GPU_TRACE timeline on stderr)."""
import os, sys, time
if os.environ.get("TRACE", "1") == "1":
    os.environ["COATI_GPU_TRACE"] = "1"  # read once by the library
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, coati_b200
from coati_b200 import capi
from synth import synth_pairs
T = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'tables.npz'))['mg_c5'].astype(np.float32)
pinned = []
def alloc(n):
    t = torch.empty(max(1, n), dtype=torch.uint8, pin_memory=True); pinned.append(t); return t.numpy()
N = int(os.environ.get("N", 1_000_000))
w = synth_pairs(N, 5, 42, alloc=alloc)
tot = int(w["a_off"][-1] + w["b_off"][-1]) + 7 * N
out_a, out_b = alloc(tot + 1), alloc(tot + 1)
out_len = np.zeros(N, np.uint64); score = np.zeros(N, np.float32); status = np.zeros(N, np.int32)
cells = float((np.diff(w["a_off"]).astype(float) * np.diff(w["b_off"]).astype(float)).sum())
ctx = coati_b200.Context(0); ctx.set_model(T, 0.001, 5/6, 1)
def once():
    ctx._check(ctx.lib.coati_gpu_alignpair_batch(ctx.h, N, w["anc_all"].ctypes.data, w["a_off"].ctypes.data_as(capi._u64p),
        w["des_all"].ctypes.data, w["b_off"].ctypes.data_as(capi._u64p),
        out_a.ctypes.data, out_b.ctypes.data, out_len.ctypes.data_as(capi._u64p), score.ctypes.data_as(capi._fp),
        status.ctypes.data_as(capi._i32p)))
once(); once()
import ctypes; sys.stderr.flush(); print("==== traced call ====", file=sys.stderr, flush=True)
t0 = time.perf_counter(); once(); t1 = time.perf_counter()
print("e2e ms %.1f GCUPS %.0f" % (1e3 * (t1 - t0), cells / (t1 - t0) / 1e9), flush=True)
for nsub in sys.argv[1:]:
    os.environ["COATI_GPU_NSUB"] = nsub
    ts = []
    for _ in range(int(os.environ.get("REPS", 3))):
        t0 = time.perf_counter(); once(); ts.append(time.perf_counter() - t0)
    print("nsub", nsub, "e2e ms", [round(1e3 * t, 1) for t in ts], "GCUPS %.0f" % (cells / min(ts) / 1e9), flush=True)
