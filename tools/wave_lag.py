"""Step time and per-band lag of the intra-pair wavefront, measured: fill time of rectangular lattices with a
fixed descendant (lb columns) and 1x / 2x / 4x / 8x as many ancestor rows.  The chain model is
    T(la) = steps_per_band * t_step + (bands - 1) * lag,   steps_per_band = lb + 31,
so the slope over the band count is the lag and the intercept is the lone-warp step time.

usage: python tools/wave_lag.py [lb] [R ...]     (one B200; prints one line per (R, la) and the fitted t_step / lag)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import coati_b200, oracle
from coati_b200.capi import PackedPairs
from tests import util

lb = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
Rs = [int(x) for x in sys.argv[2:]] or [4, 10]
T = util.load_tables()["mg_golden"]
rng = np.random.RandomState(7)
SENSE = [a + b + c for a in "ACGT" for b in "ACGT" for c in "ACGT" if a + b + c not in ("TAA", "TAG", "TGA")]
des = "".join(SENSE[i] for i in rng.randint(61, size=lb // 3))
clk = 1.965e6  # cycles per ms
for R in Rs:
    os.environ["COATI_GPU_WAVE_R"] = str(R)
    pts = []
    for la in (1920, 3840, 7680, 15360):
        anc = "".join(SENSE[i] for i in rng.randint(61, size=la // 3))
        a, b = oracle.encode_pair(anc, des)
        ctx = coati_b200.Context(0)
        ctx.set_model(T)
        pk = PackedPairs([a], [b], [anc], [des])
        bt = ctx.batch(pk.a_off, pk.b_off)
        bt.upload(pk.a_all, pk.b_all, pk.anc_all, pk.des_all)
        best = 1e9
        for _ in range(4):
            bt.run()
            best = min(best, bt.timing()["fill_ms"])
        bt.destroy(); ctx.close()
        nb = (len(a) + 32 * R - 1) // (32 * R)
        pts.append((nb, best))
        print("R %d la %d lb %d bands %d fill_ms %.3f gcups %.1f" % (R, len(a), len(b), nb, best, len(a) * len(b) / best / 1e6), flush=True)
    x = np.array([p[0] - 1 for p in pts], float); y = np.array([p[1] for p in pts])
    lag, t0 = np.polyfit(x, y, 1)
    steps = len(des) + 31
    print("R %d: t_step %.0f cycles (%.1f ns), lag per band %.0f cycles = %.1f steps" %
          (R, t0 * clk / steps, t0 * 1e6 / steps, lag * clk, lag / (t0 / steps)), flush=True)
