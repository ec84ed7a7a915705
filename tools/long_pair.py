import sys
sys.path.insert(0, '/root/repo')
import numpy as np, coati_b200, oracle
from coati_b200.capi import PackedPairs
from tests import util
name = sys.argv[1] if len(sys.argv) > 1 else "example-40k"
T = util.load_tables()["mg_golden"]
(_, anc), (_, des) = util.load_fasta(name)
anc = util.sanitise_ancestor(anc)
a, b = oracle.encode_pair(anc, des)
ctx = coati_b200.Context(0); ctx.set_model(T)
pk = PackedPairs([a], [b], [anc], [des])
bt = ctx.batch(pk.a_off, pk.b_off); bt.upload(pk.a_all, pk.b_all, pk.anc_all, pk.des_all)
bt.run(); bt.run(); print(bt.timing())
