import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch, coati_b200
from coati_b200 import capi
from synth import synth_pairs
T = np.load('/root/repo/tests/golden/tables.npz')['mg_c5'].astype(np.float32)
pinned = []
def alloc(n):
    t = torch.empty(max(1, n), dtype=torch.uint8, pin_memory=True); pinned.append(t); return t.numpy()
N = 1_000_000
w = synth_pairs(N, 5, 42, alloc=alloc)
tot = int(w["a_off"][-1] + w["b_off"][-1]) + N
out_a, out_b = alloc(tot + 1), alloc(tot + 1)
out_len = np.zeros(N, np.uint64); score = np.zeros(N, np.float32); status = np.zeros(N, np.int32)
cells = float((np.diff(w["a_off"]).astype(float) * np.diff(w["b_off"]).astype(float)).sum())
for nsub in sys.argv[1:]:
    os.environ["COATI_GPU_NSUB"] = nsub
    ctx = coati_b200.Context(0); ctx.set_model(T, 0.001, 5/6, 1)
    def once():
        ctx._check(ctx.lib.coati_gpu_viterbi_batch(ctx.h, N, w["a_all"].ctypes.data, w["a_off"].ctypes.data_as(capi._u64p),
            w["b_all"].ctypes.data, w["b_off"].ctypes.data_as(capi._u64p), w["anc_all"].ctypes.data, w["des_all"].ctypes.data,
            out_a.ctypes.data, out_b.ctypes.data, out_len.ctypes.data_as(capi._u64p), score.ctypes.data_as(capi._fp),
            status.ctypes.data_as(capi._i32p)))
    once()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); once(); ts.append(time.perf_counter() - t0)
    # host planning only
    t0 = time.perf_counter(); bt = ctx.batch(w["a_off"], w["b_off"]); t1 = time.perf_counter(); bt.destroy()
    print("nsub", nsub, "e2e ms", [round(1e3*t,1) for t in ts], "GCUPS %.0f" % (cells/min(ts)/1e9), "plan-only(1 batch) ms %.1f" % (1e3*(t1-t0)), flush=True)
    ctx.close()
