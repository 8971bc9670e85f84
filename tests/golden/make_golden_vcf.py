"""Pins oracle/vcf_records.py against the UNMODIFIED reference functions `find_candidates` (site merge + (ref, alt) dedup,
pepper_variant/modules/python/CandidateFinder.py:532-581 — run whole, process pool included) and
`VCFWriter.candidate_list_to_variant` / `write_vcf_records` (VcfWriter.py:48-218), executed in this container with stand-in
modules: the npz-free fake h5py of make_golden_candidates.py (vlen strings come back as `str`, as under h5py 2.10), a
PEPPER_VARIANT exposing FASTA_handler / CandidateImagePrediction, and a pysam whose VariantFile records every
new_record(**kw) / write() call.      python tests/golden/make_golden_vcf.py"""
import json
import os
import sys
import types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import find_candidates as ofc, vcf_records as ovr  # noqa: E402

rng = np.random.default_rng(23)
L = 3000
genome = "".join("ACGT"[i] for i in rng.integers(0, 4, L))
genome = genome[:400] + "AAAAAAAAA" + genome[409:900] + "N" + genome[901:1500] + "TTTTTTT" + genome[1507:]
# candidate records: many positions carry several alleles (multi-allelic sites, more than allowed_multiallelics at some),
# deletions of different lengths at one site (suffix normalisation), and adjacent batches repeat their boundary records
# (the shared interval end / start of the reference's tiling -> the (ref, alt) dedup)
pos_pool = np.sort(rng.choice(np.arange(5, L - 12), size=260, replace=False))
recs = []
for p in pos_pool:
    for _ in range(int(rng.choice([1, 1, 1, 2, 3, 5]))):
        t = int(rng.integers(1, 4))
        if t == 1:
            key = "1" + "ACGT"[rng.integers(0, 4)]
        elif t == 2:
            key = "2" + genome[p] + "".join("ACGT"[i] for i in rng.integers(0, 4, rng.integers(1, 5)))
        else:
            key = "3" + genome[p:p + int(rng.integers(2, 7))]
        d = int(rng.integers(4, 60))
        pr = rng.dirichlet([0.5, 0.6, 0.5])
        if rng.random() < 0.15:
            pr = np.array([0.2, 0.4, 0.4])              # ties between het and hom
        recs.append((int(p), key, d, int(rng.integers(1, d + 1)), pr.astype(np.float32).astype(np.float64)))
recs.sort(key=lambda r: r[0])
nb = 4
cuts = [0] + [len(recs) * (i + 1) // nb for i in range(nb)]
batches = []
for b in range(nb):
    lo, hi = cuts[b], cuts[b + 1]
    part = recs[lo:hi]
    if b + 1 < nb:
        part = part + recs[hi:hi + 6]                  # the next batch's first records again: duplicates
    batches.append(part)
options = dict(snp_p_value=0.1, insert_p_value=0.1, delete_p_value=0.1, snp_p_value_in_lc=0.3, insert_p_value_in_lc=0.35,
               delete_p_value_in_lc=0.25, report_snp_above_freq=0.0, report_indel_above_freq=0.6,
               allowed_multiallelics=3, snp_q_cutoff=15, indel_q_cutoff=10, snp_q_cutoff_in_lc=20, indel_q_cutoff_in_lc=12, threads=2)


def fetch(contig, a, b):
    return genome[max(0, a):max(0, b)]


class _L:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, k):
        return self.a


def _batch_group(part):
    n = len(part)
    obj = np.empty((n, 1), dtype=object)
    for i, r in enumerate(part):
        obj[i, 0] = r[1]
    return {"contigs": _L(np.array([b"ctg"] * n)), "positions": _L(np.array([r[0] for r in part])),
            "depths": _L(np.array([r[2] for r in part])), "candidates": _L(obj),
            "candidate_frequency": _L(np.array([r[3] for r in part]).reshape(n, 1)), "base_prediction": _L(np.stack([r[4] for r in part]))}


class FakeH5:
    def __init__(self, name, mode="r"):
        self.d = {"predictions": {"batch_%d" % b: _batch_group(p) for b, p in enumerate(batches)}}

    def keys(self):
        return self.d.keys()

    def __getitem__(self, k):
        return self.d[k]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


class FASTA_handler:
    def __init__(self, path):
        pass

    def get_reference_sequence(self, contig, a, b):
        return fetch(contig, a, b)

    def get_chromosome_names(self):
        return ["ctg"]

    def get_chromosome_sequence_length(self, c):
        return L


class CandidateImagePrediction:
    def __init__(self, contig, position, depth, candidates, candidate_frequency, prediction_base, prediction_type):
        self.contig, self.position, self.depth = contig, position, depth
        self.candidates, self.candidate_frequency = candidates, candidate_frequency
        self.prediction_base, self.prediction_type = prediction_base, prediction_type


WRITTEN = []


class _Header:
    def __init__(self):
        self.contigs = self

    def add_meta(self, **kw):
        pass

    def add(self, *a, **kw):
        pass

    def add_sample(self, s):
        pass


class _VariantFile:
    def __init__(self, name, mode, header=None):
        self.name = os.path.basename(name)

    def new_record(self, **kw):
        return dict(kw)

    def write(self, rec):
        WRITTEN.append((self.name, rec))

    def close(self):
        pass


h5 = types.ModuleType("h5py"); h5.File = FakeH5
pv = types.ModuleType("pepper_variant.build.PEPPER_VARIANT")
pv.FASTA_handler = FASTA_handler; pv.CandidateImagePrediction = CandidateImagePrediction
build = types.ModuleType("pepper_variant.build"); build.PEPPER_VARIANT = pv
ps = types.ModuleType("pysam"); ps.VariantFile = _VariantFile; ps.VariantHeader = _Header; ps.tabix_index = lambda *a, **k: None
sys.modules.update({"h5py": h5, "pepper_variant.build": build, "pepper_variant.build.PEPPER_VARIANT": pv, "pysam": ps})
sys.path.insert(0, "/root/reference")
from pepper_variant.modules.python import CandidateFinder as RefCF  # noqa: E402
from pepper_variant.modules.python import VcfWriter as RefVW  # noqa: E402

opt = types.SimpleNamespace(fasta="x", **options)
pairs = [("f", "batch_%d" % b) for b in range(nb)]
contigs, sites_phasing, sites_vc = RefCF.find_candidates(opt, "dir", pairs)
w = RefVW.VCFWriter(contigs, "x", "sample", "out/", "full", "pepper", "vc")
counts = w.write_vcf_records(sites_vc, opt)
tag = {"full.vcf.gz": "full", "pepper.vcf.gz": "pepper", "vc.vcf.gz": "variant_calling", "vc_SNPs.vcf.gz": "variant_calling_snp",
       "vc_INDEL.vcf.gz": "variant_calling_indel"}
want = []
for name, rec in WRITTEN:
    if tag[name] == "full":
        want.append(dict(rec, files=["full"]))
    else:
        assert want[-1]["start"] == rec["start"]
        want[-1]["files"].append(tag[name])


def norm(o):
    if isinstance(o, dict):
        return {k: norm(v) for k, v in sorted(o.items())}
    if isinstance(o, (list, tuple)):
        return [norm(x) for x in o]
    if isinstance(o, (np.floating, float)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.bool_, bool)):
        return bool(o)
    return o


# ---- the restatement on the same inputs
dv_all = []
for part in batches:
    m, d = ofc.select(options, "ctg", [r[0] for r in part], [r[2] for r in part], [r[1] for r in part], [r[3] for r in part],
                      [r[4] for r in part], fetch)
    dv_all.extend(d)
c2, sites = ovr.merge_sites(dv_all)
assert c2 == contigs and sorted(sites) == sorted(sites_vc)
got = ovr.vcf_records(sites, options)
gj, wj = [json.dumps(norm(r)) for r in got], [json.dumps(norm(dict(r, id=None))) for r in want]
gj = [json.dumps(norm(dict(r, id=None))) for r in got]
wj = [json.dumps(norm({k: (None if k == "id" else v) for k, v in r.items()})) for r in want]
assert gj == wj, next((a, b) for a, b in zip(gj, wj) if a != b)
multi = sum(1 for r in want if len(r["alleles"]) > 2)
np.savez_compressed(os.path.join(HERE, "vcf_seed23.npz"), genome=np.array(genome), options=np.array(json.dumps(options)),
                    batch_sizes=np.array([len(p) for p in batches]), positions=np.array([r[0] for p in batches for r in p]),
                    keys=np.array([r[1] for p in batches for r in p]), depths=np.array([r[2] for p in batches for r in p]),
                    freqs=np.array([r[3] for p in batches for r in p]), probs=np.stack([r[4] for p in batches for r in p]),
                    records=np.array(wj), counts=np.array(counts))
print("vcf assembly pinned: %d candidate records -> %d sites written (%d multi-allelic), counts %s" % (sum(len(p) for p in batches), len(want), multi, counts))
