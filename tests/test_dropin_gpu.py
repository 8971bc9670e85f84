"""GPU: the drop-in classes carrying the reference's own names (pepper_b200/build/PEPPER*.py) driven exactly like
pepper_variant/modules/python/AlignmentSummarizer.py:220-238 and pepper/.../AlignmentSummarizer.py:341-350 drive the
pybind modules."""
import numpy as np
import pytest

from pepper_b200 import synth

pytestmark = pytest.mark.gpu
NT16 = synth.NT16


def _to_type_reads(mod, reads):
    out = []
    codes = reads.codes()
    for r in range(reads.n_reads):
        t = mod.type_read()
        t.pos = int(reads.pos[r])
        b0, b1 = int(reads.seq_off[r]), int(reads.seq_off[r + 1])
        t.sequence = "".join(NT16[c] for c in codes[b0:b1])
        t.base_qualities = reads.qual[b0:b1].tolist()
        t.cigar_tuples = [mod.CigarOp(int(w & 15), int(w >> 4)) for w in reads.cigar[reads.cigar_off[r]:reads.cigar_off[r + 1]]]
        t.mapping_quality = int(reads.mapq[r])
        t.flags.is_reverse = bool(reads.flags[r] & 1)
        out.append(t)
    return out


def test_regional_summary_generator_dropin(oracle_built):
    from pepper_b200.build import PEPPER_VARIANT
    reads, regions = synth.make_variant_workload(1, 3000, 25, synth.ONT, seed=61)
    row = regions.table[0]
    ref = bytes(regions.ref[row[4]:row[4] + row[5]]).decode()
    p = synth.ont_params()
    gen = PEPPER_VARIANT.RegionalSummaryGenerator("chr20", int(row[0]), int(row[1]), ref)
    tr = _to_type_reads(PEPPER_VARIANT, reads)
    gen.generate_max_insert_summary(tr)
    cands = gen.generate_summary(tr, p["min_snp_baseq"], p["min_indel_baseq"], p["snp_freq_threshold"], p["insert_freq_threshold"],
                                 p["delete_freq_threshold"], p["min_coverage_threshold"], p["snp_candidate_freq_threshold"],
                                 p["indel_candidate_freq_threshold"], p["candidate_support_threshold"], False, int(row[2]), int(row[3]),
                                 32, 26, False)
    want = oracle_built.variant_encode(reads, regions, p, "port")
    assert [c.candidates[0] for c in cands] == want["keys"]
    assert [c.position for c in cands] == want["positions"].tolist()
    assert [c.depth for c in cands] == want["depths"].tolist()
    assert [c.candidate_frequency[0] for c in cands] == want["freqs"].tolist()
    assert np.array_equal(np.stack([c.image_matrix for c in cands]), oracle_built.images_to_int8(want["images"]))
    assert cands[0].contig == "chr20" and cands[0].base_label == 0


def test_summary_generator_dropin(oracle_built):
    from pepper_b200.build import PEPPER
    reads, regions = synth.make_polish_workload(1, 30, synth.ONT, seed=62)
    row = regions.table[0]
    sg = PEPPER.SummaryGenerator("", "contig_1", int(row[0]), int(row[1]))
    sg.generate_summary(_to_type_reads(PEPPER, reads), int(row[0]), int(row[1]))
    want = oracle_built.polish_encode(reads, regions, "port")
    assert np.array_equal(np.asarray(sg.image), want["image"])
    assert sg.genomic_pos == list(zip(want["pos"].tolist(), want["idx"].tolist()))
