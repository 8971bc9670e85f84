import time, sys, numpy as np
sys.path.insert(0, '.')
from pepper_b200 import synth
from pepper_b200.variant import VariantEncoder
t=time.time()
reads, regions = synth.make_variant_workload(4, 100000, 30, synth.ONT, seed=1)
print("gen %.1fs reads %d bases %d ops %d" % (time.time()-t, reads.n_reads, reads.n_bases, int(reads.cigar_off[-1])))
enc = VariantEncoder(0)
for i in range(3):
    t=time.time(); c = enc.encode(reads, regions, synth.ont_params()); dt=time.time()-t
    print("encode host wall %.3fs cands %d" % (dt, len(c)), enc.timings())
