"""Hand-written known-answer inputs, each aimed at one quirk of the reference encoders
(SURVEY.md §8c list).  Shared by the CPU tests (port vs the compiled reference) and the GPU
parity tests (CUDA vs port)."""
import numpy as np

from pepper_b200 import synth
from pepper_b200.synth import OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X, RegionTable, make_batch


def _region(ref: str, ref_start: int, n_reads: int, cand=None):
    L = len(ref)
    cs, ce = cand if cand else (ref_start, ref_start + L - 1)
    tab = np.array([[ref_start, ref_start + L - 1, cs, ce, 0, L, 0, n_reads]], dtype=np.int64)
    return RegionTable(tab, np.frombuffer(ref.encode(), dtype=np.uint8).copy())


def _ref(n, seed=1):
    rng = np.random.default_rng(seed)
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


LOOSE = dict(min_snp_baseq=1, min_indel_baseq=1, snp_freq_threshold=0.10, insert_freq_threshold=0.10,
             delete_freq_threshold=0.10, min_coverage_threshold=1, snp_candidate_freq_threshold=0.10,
             indel_candidate_freq_threshold=0.10, candidate_support_threshold=1, skip_indels=0)


def variant_kats():
    out = []
    ref = _ref(200, 11)
    S = 1000

    def mut(s, i, c):
        return s[:i] + c + s[i + 1:]

    # 1. plain SNPs on both strands + a matching read; candidate near both region edges (zero padding)
    reads = []
    for k in range(6):
        seq = ref
        seq = mut(seq, 3, "A" if ref[3] != "A" else "C")           # position < 16
        seq = mut(seq, 100, "T" if ref[100] != "T" else "G")
        seq = mut(seq, 197, "G" if ref[197] != "G" else "A")        # position > L-16
        reads.append(dict(pos=S, seq=seq, qual=30, cigar=[(OP_M, 200)], reverse=k % 2 == 1))
    reads.append(dict(pos=S, seq=ref, qual=30, cigar=[(OP_M, 200)]))
    out.append(("snp_edges", make_batch(reads), _region(ref, S, len(reads)), LOOSE))

    # 2. insert / delete anchors: ...M I M..., ...M D M..., read ending in M I and M D, =/X ops
    reads = []
    for k in range(5):
        ins = "GATTACA"[: 1 + k % 3]
        seq = ref[:50] + ins + ref[50:120] + ref[125:]
        reads.append(dict(pos=S, seq=seq, qual=25, reverse=k % 2 == 0,
                          cigar=[(OP_EQ, 30), (OP_X, 20), (OP_I, len(ins)), (OP_M, 70), (OP_D, 5), (OP_M, 75)]))
    reads.append(dict(pos=S + 10, seq=ref[10:60] + "AC", qual=20, cigar=[(OP_M, 50), (OP_I, 2)]))
    reads.append(dict(pos=S + 10, seq=ref[10:60], qual=20, cigar=[(OP_M, 50), (OP_D, 4)], reverse=True))
    out.append(("indel_anchors", make_batch(reads), _region(ref, S, len(reads)), LOOSE))

    # 3. N (ref skip) and P ops: the walker advances BOTH ref and read index (fall-through, :556-560)
    reads = []
    for k in range(4):
        reads.append(dict(pos=S + 5, seq=ref[5:85], qual=30, reverse=k % 2 == 1,
                          cigar=[(OP_M, 30), (OP_N, 10), (OP_M, 20), (OP_P, 3), (OP_M, 17)]))
    out.append(("refskip_pad_fallthrough", make_batch(reads), _region(ref, S, len(reads)), LOOSE))

    # 4. soft clip inside, hard clip, low qualities below min_snp_baseq, mapq 0 read ignored
    p = dict(LOOSE); p["min_snp_baseq"] = 10; p["min_indel_baseq"] = 10
    reads = []
    for k in range(5):
        seq = mut(ref[:90], 40, "A" if ref[40] != "A" else "C") + "TTTT"
        q = [30] * 94
        q[40] = 5 if k == 0 else 30
        q[20] = 3
        reads.append(dict(pos=S, seq=seq, qual=q, cigar=[(OP_M, 90), (OP_S, 4), (OP_H, 10)], reverse=k % 2 == 1))
    reads.append(dict(pos=S, seq=mut(ref[:90], 41, "N"), qual=30, cigar=[(OP_M, 90)], mapq=0))
    out.append(("clips_lowq_mapq0", make_batch(reads), _region(ref, S, len(reads)), p))

    # 5. insert quality rule (:453): anchor below min_snp_baseq but insert mean quality ok -> coverage += 1
    reads = []
    for k in range(4):
        seq = ref[:60] + "GG" + ref[60:100]
        q = [30] * len(seq)
        q[59] = 2                       # anchor base
        reads.append(dict(pos=S, seq=seq, qual=q, cigar=[(OP_M, 60), (OP_I, 2), (OP_M, 40)], reverse=k % 2 == 1))
    reads.append(dict(pos=S, seq=ref[:100], qual=30, cigar=[(OP_M, 100)]))
    out.append(("insert_anchor_quality", make_batch(reads), _region(ref, S, len(reads)), p))

    # 6. long indels: insert of 59/60/61 bases (key length 61/62/63), deletion of 59/60/61
    reads = []
    for n in (58, 59, 60, 61):
        ins = ("ACGT" * 20)[:n]
        for k in range(2):
            reads.append(dict(pos=S, seq=ref[:30] + ins + ref[30:60], qual=30, reverse=k == 1,
                              cigar=[(OP_M, 30), (OP_I, n), (OP_M, 30)]))
            reads.append(dict(pos=S + 60, seq=ref[60:80] + ref[80 + n:80 + n + 20], qual=30, reverse=k == 1,
                              cigar=[(OP_M, 20), (OP_D, n), (OP_M, 20)]))
    out.append(("long_indels", make_batch(reads), _region(ref, S, len(reads)), LOOSE))

    # 7. many distinct insert alleles at one site (ordering = std::set<string>, prefix before longer)
    reads = []
    alleles = ["A", "AA", "AAC", "C", "CA", "T", "G", "GT", "A", "AA", "T", "T", "ACGTACGTACGTACGTA", "ACGTACGTACGTACGTC",
               "ACGTACGTACGTACG", "ACGTACGTACGTACGTA"]
    for k, al in enumerate(alleles):
        reads.append(dict(pos=S + 20, seq=ref[20:70] + al + ref[70:120], qual=30, reverse=k % 3 == 0,
                          cigar=[(OP_M, 50), (OP_I, len(al)), (OP_M, 50)]))
    out.append(("insert_allele_order", make_batch(reads), _region(ref, S, len(reads)), LOOSE))

    # 8. deep pileup: > 125 reads on one strand (unclamped cols 4/8-10/15/25 wrap in int8, clamped ones saturate)
    reads = []
    alt = mut(ref[:60], 30, "A" if ref[30] != "A" else "C")
    for k in range(300):
        seq = alt if k % 2 == 0 else ref[:60]
        reads.append(dict(pos=S, seq=seq, qual=30, cigar=[(OP_M, 60)], reverse=(k % 5 == 0)))
    for k in range(140):
        reads.append(dict(pos=S, seq=ref[:20] + ref[23:60], qual=30, cigar=[(OP_M, 20), (OP_D, 3), (OP_M, 37)],
                          reverse=(k % 7 == 0)))
    out.append(("deep_pileup_wrap", make_batch(reads), _region(ref, S, len(reads)), LOOSE))

    # 9. non-ACGT read bases (N, IUPAC R, IUPAC D) as SNP alleles
    reads = []
    for k, b in enumerate("NNRRDDNNAA"):
        reads.append(dict(pos=S, seq=mut(ref[:80], 40, b), qual=30, cigar=[(OP_M, 80)], reverse=k % 2 == 1))
    out.append(("iupac_alleles", make_batch(reads), _region(ref, S, len(reads)), LOOSE))

    # 10. candidate region narrower than the region + reads partly outside the region + read starting before it
    reads = []
    for k in range(6):
        seq = mut(mut(ref, 60, "A" if ref[60] != "A" else "C"), 150, "T" if ref[150] != "T" else "G")
        reads.append(dict(pos=S - 20, seq="ACGTACGTACGTACGTACGT" + seq + "ACGTACGTAC", qual=30, reverse=k % 2 == 1,
                          cigar=[(OP_M, 230)]))
    out.append(("cand_window_and_overhang", make_batch(reads), _region(ref, S, len(reads), cand=(S + 50, S + 100)), LOOSE))

    # 11. skip_indels + support threshold 2
    p2 = dict(LOOSE); p2["skip_indels"] = 1; p2["candidate_support_threshold"] = 2
    reads = []
    for k in range(4):
        seq = mut(ref[:50], 25, "A" if ref[25] != "A" else "C") + "G" + ref[50:100]
        reads.append(dict(pos=S, seq=seq, qual=30, cigar=[(OP_M, 50), (OP_I, 1), (OP_M, 50)], reverse=k % 2 == 1))
    reads.append(dict(pos=S, seq=mut(ref[:100], 70, "A" if ref[70] != "A" else "C"), qual=30, cigar=[(OP_M, 100)]))
    out.append(("skip_indels_support", make_batch(reads), _region(ref, S, len(reads)), p2))

    # 12. insert directly after a deletion (anchor is a deleted base), and I as the second op after 1 M
    reads = []
    for k in range(4):
        reads.append(dict(pos=S, seq=ref[:40] + "TT" + ref[45:100], qual=30, reverse=k % 2 == 1,
                          cigar=[(OP_M, 40), (OP_D, 5), (OP_I, 2), (OP_M, 55)]))
        reads.append(dict(pos=S + 120, seq=ref[120] + "CA" + ref[121:160], qual=30, reverse=k % 2 == 0,
                          cigar=[(OP_M, 1), (OP_I, 2), (OP_M, 39)]))
    out.append(("insert_after_delete", make_batch(reads), _region(ref, S, len(reads)), LOOSE))
    return out


def polish_kats():
    out = []
    ref = _ref(300, 5)
    S = 500

    def reg(n_reads, s=S, e=S + 299):
        return RegionTable(np.array([[s, e, s, e, 0, 0, 0, n_reads]], dtype=np.int64), np.zeros(1, np.uint8))

    # 1. deletion: coverage lumped on the first deleted base; cov=0 columns wrap ((3*254)&255 = 250)
    reads = []
    for k in range(3):
        reads.append(dict(pos=S + 100, seq=ref[100:120] + ref[130:150], qual=30, reverse=k == 1,
                          cigar=[(OP_M, 20), (OP_D, 10), (OP_M, 20)]))
    out.append(("deletion_coverage_lump", make_batch(reads), reg(len(reads))))
    # 2. inserts of different lengths at one anchor, forward/reverse feature order, N bases
    reads = []
    for k, ins in enumerate(["A", "ACG", "TTN", "G", "ACGTA"]):
        reads.append(dict(pos=S + 10, seq=ref[10:60] + ins + ref[60:110], qual=30, reverse=k % 2 == 1,
                          cigar=[(OP_M, 50), (OP_I, len(ins)), (OP_M, 50)]))
    out.append(("insert_columns", make_batch(reads), reg(len(reads))))
    # 3. N / P ops counted like deletions, soft clips, mapq 0 skipped, reads overhanging both region ends
    reads = [dict(pos=S - 30, seq=_ref(400, 9), qual=30, cigar=[(OP_M, 400)]),
             dict(pos=S + 20, seq=ref[20:50] + ref[55:80], qual=30, cigar=[(OP_M, 30), (OP_N, 5), (OP_M, 25)], reverse=True),
             dict(pos=S + 20, seq="GG" + ref[20:50] + ref[53:80], qual=30, cigar=[(OP_S, 2), (OP_M, 30), (OP_P, 3), (OP_M, 27)]),
             dict(pos=S + 20, seq=ref[20:80], qual=30, cigar=[(OP_M, 60)], mapq=0),
             dict(pos=S + 290, seq=ref[290:300] + "ACGTACGT", qual=30, cigar=[(OP_M, 18)]),
             dict(pos=S + 280, seq=ref[280:300], qual=30, cigar=[(OP_M, 20), (OP_D, 15)])]
    out.append(("skips_clips_overhang", make_batch(reads), reg(len(reads))))
    # 4. insert right after a deletion (anchor has zero coverage -> count*254 wraps), insert at region end
    reads = []
    for k in range(3):
        reads.append(dict(pos=S + 50, seq=ref[50:70] + "AAC" + ref[75:100], qual=30, reverse=k == 0,
                          cigar=[(OP_M, 20), (OP_D, 5), (OP_I, 3), (OP_M, 25)]))
    reads.append(dict(pos=S + 250, seq=ref[250:300] + "TT", qual=30, cigar=[(OP_M, 50), (OP_I, 2)]))
    out.append(("insert_after_delete", make_batch(reads), reg(len(reads))))
    return out


# ----------------------------------------------------------------------------------------------------------------
# get_reads KATs (row a2): records + queries that walk every branch of bam_handler.cpp:176-303
def getreads_kats():
    """[(name, RecordBatch, [(start, stop, include_supplementary, min_mapq, min_baseq), ...])]"""
    from pepper_b200.synth import make_records, OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X
    A = "ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"

    def seq(n, off=0):
        return (A * (n // len(A) + 2))[off:off + n]
    recs = [
        # 0: plain match spanning the query on both sides
        dict(pos=100, seq=seq(200), cigar=[(OP_M, 200)]),
        # 1: leading hard + soft clip, insertion exactly at `start` (dropped: no anchor yet), match, deletion, match
        dict(pos=150, seq=seq(10 + 5 + 40 + 30), cigar=[(OP_H, 7), (OP_S, 10), (OP_I, 5), (OP_M, 40), (OP_D, 6), (OP_M, 30)], flag=16),
        # 2: starts in a deletion relative to start: M ends before start, D spans start, then M
        dict(pos=120, seq=seq(20 + 50, 3), cigar=[(OP_M, 20), (OP_D, 40), (OP_M, 50)]),
        # 3: insertion and soft clip right at stop / stop+1, = and X ops, trailing hard clip
        dict(pos=180, seq=seq(20 + 4 + 20 + 6 + 9, 1), qual=[5, 40] * 29 + [7], cigar=[(OP_EQ, 20), (OP_I, 4), (OP_X, 20), (OP_I, 6), (OP_S, 9), (OP_H, 3)]),
        # 4: ref skip (N) crossing stop, pad (P) and zero-length ops in the middle
        dict(pos=185, seq=seq(10 + 3 + 10, 2), cigar=[(OP_M, 10), (OP_P, 4), (OP_I, 3), (OP_M, 0), (OP_N, 50), (OP_M, 10)]),
        # 5..9: filtered by flag
        dict(pos=190, seq=seq(30), cigar=[(OP_M, 30)], flag=256),
        dict(pos=191, seq=seq(30), cigar=[(OP_M, 30)], flag=512),
        dict(pos=192, seq=seq(30), cigar=[(OP_M, 30)], flag=1024),
        dict(pos=193, seq=seq(30), cigar=[(OP_M, 30)], flag=4),
        dict(pos=194, seq=seq(30), cigar=[(OP_M, 30)], flag=2048 | 16),
        # 10: low MAPQ
        dict(pos=195, seq=seq(30), cigar=[(OP_M, 30)], mapq=3),
        # 11: non-ACGT bases and low qualities (bad_indicies rule)
        dict(pos=196, seq="ACGNNRYACGTACGTAAAAA", qual=[1, 2, 3, 40, 40, 40, 40, 9, 10, 11] * 2, cigar=[(OP_M, 20)]),
        # 12: overlaps the query only with a deletion -> visited by the iterator, keeps no base, dropped (:432)
        dict(pos=140, seq=seq(5 + 5), cigar=[(OP_M, 5), (OP_D, 200), (OP_M, 5)]),
        # 13: record without CIGAR (iterator gives it length 1)
        dict(pos=210, seq=seq(12), cigar=[]),
        # 14: single base at exactly stop (inclusive stop)
        dict(pos=260, seq=seq(40), cigar=[(OP_M, 40)]),
        # 15: starts exactly at stop+... beyond the query
        dict(pos=261, seq=seq(40), cigar=[(OP_S, 5), (OP_M, 35)]),
        # 16: insertion first, then match (I before any anchor inside the region)
        dict(pos=205, seq=seq(8 + 30), cigar=[(OP_I, 8), (OP_M, 30)]),
        # 17: match ending exactly at stop followed by insertion at stop+1 (cut) and deletion
        dict(pos=231, seq=seq(30 + 5 + 10), cigar=[(OP_M, 30), (OP_I, 5), (OP_D, 3), (OP_M, 10)]),
    ]
    recs.sort(key=lambda r: r["pos"])
    batch = make_records(recs)
    queries = [(160, 260, False, 0, 0), (160, 260, True, 0, 10), (160, 260, False, 20, 0), (100, 101, False, 0, 0),
               (0, 100, False, 0, 0), (299, 400, False, 0, 0), (260, 261, True, 0, 0), (205, 206, False, 0, 7),
               (150, 340, True, 0, 0), (199, 200, False, 0, 0), (139, 145, False, 0, 0), (330, 1000, True, 0, 0)]
    return [("branches", batch, queries)]
