"""Generates the golden fixtures from the UNMODIFIED reference encoders (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
Inputs are the seeded synthetic workloads of pepper_b200/synth.py, so only outputs are stored."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from pepper_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

oracle.build()
assert oracle.have_ref(), "needs /root/reference"

reads, regions = synth.make_variant_workload(2, 5000, 30, synth.ONT, seed=21)
b = oracle.variant_encode(reads, regions, synth.ont_params(), "ref")
np.savez_compressed(os.path.join(HERE, "variant_ont_seed21.npz"), images=oracle.images_to_int8(b["images"]),
                    positions=b["positions"], depths=b["depths"], freqs=b["freqs"],
                    keys=np.array([k.encode() for k in b["keys"]]), region_of=b["region_of"])
print("variant candidates:", len(b["keys"]))

reads, regions = synth.make_polish_workload(3, 40, synth.ONT, seed=22)
c = oracle.polish_encode(reads, regions, "ref")
np.savez_compressed(os.path.join(HERE, "polish_ont_seed22.npz"), **c)
print("polish columns:", c["image"].shape)
