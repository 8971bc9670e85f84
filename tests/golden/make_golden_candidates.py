"""Pins oracle/find_candidates.py against the UNMODIFIED reference `small_chunk_stitch`
(pepper_variant/modules/python/CandidateFinder.py:356) run here with stand-in modules: an npz-backed h5py that returns
`str` objects for vlen strings (as h5py 2.10, requirements.txt:1, did) and a PEPPER_VARIANT exposing FASTA_handler /
CandidateImagePrediction.      python tests/golden/make_golden_candidates.py"""
import json
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import find_candidates as ofc  # noqa: E402

rng = np.random.default_rng(17)
L = 1200
genome = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
genome = genome[:200] + "AAAAAAAA" + genome[208:500] + "N" + genome[501:700] + "TTTTTT" + genome[706:]
n = 400
positions = np.sort(rng.integers(0, L - 1, n))
positions[:3] = [0, 3, 9]
positions[-2:] = [L - 3, L - 1]
positions = np.sort(positions)
keys, freqs, depths = [], [], []
for p in positions:
    t = rng.integers(1, 4)
    if t == 1:
        keys.append("1" + "ACGTN"[rng.integers(0, 5)])
    elif t == 2:
        keys.append("2" + "".join("ACGT"[i] for i in rng.integers(0, 4, rng.integers(2, 6))))
    else:
        keys.append("3" + genome[p:p + int(rng.integers(2, 6))])
    d = int(rng.integers(4, 60)); depths.append(d); freqs.append(int(rng.integers(1, d + 1)))
probs = rng.dirichlet([0.6, 0.5, 0.4], n)
probs[::17] = [0.4, 0.3, 0.3]
options = dict(snp_p_value=0.1, insert_p_value=0.1, delete_p_value=0.1, snp_p_value_in_lc=0.3, insert_p_value_in_lc=0.35,
               delete_p_value_in_lc=0.25, report_snp_above_freq=0.0, report_indel_above_freq=0.6)


def fetch(contig, a, b):
    return genome[max(0, a):max(0, b)]


class FakeH5:
    def __init__(self, name, mode="r"):
        obj = np.empty((n, 1), dtype=object)
        for i, k in enumerate(keys):
            obj[i, 0] = k
        self.d = {"predictions": {"batch_0": {"contigs": _L(np.array([b"ctg"] * n)), "positions": _L(np.array(positions)),
                                                "depths": _L(np.array(depths)), "candidates": _L(obj),
                                                "candidate_frequency": _L(np.array(freqs).reshape(n, 1)),
                                                "base_prediction": _L(probs)}}}

    def keys(self):
        return self.d.keys()

    def __getitem__(self, k):
        return self.d[k]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


class _L:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, k):
        return self.a


h5 = types.ModuleType("h5py"); h5.File = FakeH5
pv = types.ModuleType("pepper_variant.build.PEPPER_VARIANT")


class FASTA_handler:
    def __init__(self, path):
        pass

    def get_reference_sequence(self, contig, a, b):
        return fetch(contig, a, b)


class CandidateImagePrediction:
    def __init__(self, contig, position, depth, candidates, candidate_frequency, prediction_base, prediction_type):
        self.contig, self.position, self.depth = contig, position, depth
        self.candidates, self.candidate_frequency = candidates, candidate_frequency
        self.prediction_base, self.prediction_type = prediction_base, prediction_type


pv.FASTA_handler = FASTA_handler; pv.CandidateImagePrediction = CandidateImagePrediction
build = types.ModuleType("pepper_variant.build"); build.PEPPER_VARIANT = pv
sys.modules["h5py"] = h5
sys.modules["pepper_variant.build"] = build
sys.modules["pepper_variant.build.PEPPER_VARIANT"] = pv
sys.path.insert(0, "/root/reference")
from pepper_variant.modules.python import CandidateFinder as RefCF  # noqa: E402

opt = types.SimpleNamespace(fasta="x", **options)
want_m, want_d = RefCF.small_chunk_stitch(opt, [("f", "batch_0")])
got_m, got_d = ofc.select(options, "ctg", positions, depths, keys, freqs, probs, fetch)


def norm(rec):
    return json.dumps([x.tolist() if hasattr(x, "tolist") else (bool(x) if isinstance(x, (bool, np.bool_)) else x) for x in rec], default=float)


assert [norm(r) for r in want_m] == [norm(r) for r in got_m], (len(want_m), len(got_m))
assert [norm(r) for r in want_d] == [norm(r) for r in got_d], (len(want_d), len(got_d))
np.savez_compressed(os.path.join(HERE, "candidates_seed17.npz"), genome=np.array(genome), positions=positions, depths=np.array(depths),
                    freqs=np.array(freqs), keys=np.array(keys), probs=probs, options=np.array(json.dumps(options)),
                    margin=np.array([norm(r) for r in want_m]), deepvariant=np.array([norm(r) for r in want_d]))
print("find_candidates pinned: %d candidates -> %d margin, %d deepvariant records (%d in repeats)" %
      (n, len(want_m), len(want_d), sum(1 for r in want_d if r[11])))
