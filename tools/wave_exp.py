import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, coati_b200, oracle
from coati_b200.capi import PackedPairs
from tests import util
T = util.load_tables()["mg_golden"]
names = sys.argv[1:] or ("example-10k", "example-40k", "example-160k")
for name in names:
    (_, anc), (_, des) = util.load_fasta(name)
    anc = util.sanitise_ancestor(anc)
    a, b = oracle.encode_pair(anc, des)
    for R in (2, 4, 8, 10):
        os.environ["COATI_GPU_WAVE_R"] = str(R)
        ctx = coati_b200.Context(0)
        ctx.set_model(T)
        pk = PackedPairs([a], [b], [anc], [des])
        bt = ctx.batch(pk.a_off, pk.b_off)
        bt.upload(pk.a_all, pk.b_all, pk.anc_all, pk.des_all)
        bt.run(); bt.run()
        tm = bt.timing()
        bt.destroy(); ctx.close()
        print(name, "R", R, "fill_ms %.2f gcups %.1f tb_ms %.2f" % (tm["fill_ms"], len(a)*len(b)/tm["fill_ms"]/1e6, tm["traceback_ms"]), flush=True)
