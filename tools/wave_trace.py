"""Where the lag between consecutive bands of the wavefront accrues (diagnostics build -DCOATI_WAVE_TRACE):
per band, %globaltimer at: ticket, first columns of the row above seen, 32 / 1024 / 4096 steps done, band done.
usage: COATI_GPU_LIB=tools/gpu/ab_trace.so COATI_GPU_WAVE_R=4 python tools/wave_trace.py [la] [lb]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import coati_b200, oracle
from coati_b200.capi import PackedPairs
from tests import util
la = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
lb = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
R = int(os.environ.get("COATI_GPU_WAVE_R", "4"))
T = util.load_tables()["mg_golden"]
rng = np.random.RandomState(7)
SENSE = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT" if a + b + c not in ("TAA", "TAG", "TGA")]
des = "".join(SENSE[i] for i in rng.randint(61, size=lb // 3))
anc = "".join(SENSE[i] for i in rng.randint(61, size=la // 3))
a, b = oracle.encode_pair(anc, des)
ctx = coati_b200.Context(0); ctx.set_model(T)
pk = PackedPairs([a], [b], [anc], [des])
bt = ctx.batch(pk.a_off, pk.b_off); bt.upload(pk.a_all, pk.b_all, pk.anc_all, pk.des_all)
bt.run(); bt.run()
print(bt.timing())
nb = (len(a) + 32 * R - 1) // (32 * R)
buf = np.zeros(8 * nb, dtype=np.uint64)
rc = ctx.lib.coati_gpu_debug_wave_trace(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size))
assert rc == 0, rc
tr = buf.reshape(nb, 8).astype(np.int64)
t0 = tr[0, 0]
names = ["ticket", "row above seen", "32 steps", "1024 steps", "4096 steps", "done"]
print("band", *["%15s" % n for n in names])
for bnd in list(range(0, min(nb, 6))) + [nb // 2, nb - 1]:
    print("%4d" % bnd, *["%15.2f" % ((tr[bnd, k] - t0) / 1e3) for k in range(6)], " (us)")
d = np.diff(tr, axis=0) / 1e3
print("mean lag between consecutive bands, us:", *["%s %.2f" % (names[k], d[1:, k].mean()) for k in range(6)])
print("median:", *["%s %.2f" % (names[k], np.median(d[1:, k])) for k in range(6)])
seg = (tr[:, 1:6] - tr[:, 0:5]) / 1e3
print("segment durations of band nb/2, us: wait-for-row %.2f, first 32 steps %.2f, to 1024 %.2f, to 4096 %.2f, to end %.2f" % tuple(seg[nb // 2]))
print("band 0: %.2f %.2f %.2f %.2f %.2f" % tuple(seg[0]))
