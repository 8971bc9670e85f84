"""Golden fixture for row f1 (read -> reference realignment): outputs of the UNMODIFIED reference
ReadAligner::align_reads_to_reference (simple_aligner.cpp + ssw_cpp.cpp + ssw.c compiled into
oracle/_ref/libref_realign.so) on a seeded synthetic polish workload.
Run in the build container:  python tests/golden/make_golden_realign.py"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from pepper_b200 import synth  # noqa: E402


def workload():
    from pepper_b200.realign import realign_regions
    reads, regions = synth.make_polish_workload(2, 10, synth.ONT, seed=41)
    return reads, realign_regions(regions, synth.make_reference(2 * 1000 + 1, 41))


if __name__ == "__main__":
    from oracle import oracle
    oracle.build()
    assert oracle.have_ref_realign(), "needs /root/reference"
    reads, regions = workload()
    out = {}
    for r in range(regions.n_regions):
        row = regions.table[r]
        ref = regions.ref[int(row[4]):int(row[4] + row[5])].tobytes().decode()
        pos, pos_end, co, cig = oracle.realign(reads, int(row[6]), int(row[7]), int(row[0]), int(row[1]) + 20, ref, impl="ref")
        out[f"r{r}_pos"], out[f"r{r}_pos_end"], out[f"r{r}_cigar_off"], out[f"r{r}_cigar"] = pos, pos_end, co, cig
        print(r, pos.shape[0], cig.shape[0])
    np.savez_compressed(os.path.join(HERE, "realign_seed41.npz"), **out)
