"""TEST INFRASTRUCTURE ONLY (oracle).  Restatement of the per-candidate selection of
pepper_variant/modules/python/CandidateFinder.py:356-530 (small_chunk_stitch) for prediction records that carry one
allele each (what RegionalSummaryGenerator emits, region_summary.cpp:861-894).  Pinned against the UNMODIFIED
reference function by tests/golden/make_golden_candidates.py (stand-in modules for h5py / PEPPER_VARIANT)."""
from __future__ import annotations

import numpy as np


def repeat_annotation_hp(seq: str):
    """CandidateFinder.py:279-297 with kmer_size 1: the homopolymer run length seen from each start index."""
    n = len(seq)
    out = [1] * n
    for i in range(n):
        cnt, end = 0, i
        for j in range(i, n):
            if seq[i] == seq[j]:
                cnt += 1
            else:
                break
            end = j + 1
        for k in range(i, min(n, end)):
            out[k] = max(out[k], cnt)
    return out


def select(options: dict, contig: str, positions, depths, keys, freqs, probs, fetch):
    """fetch(contig, start, stop) -> reference string [start, stop) (clipped at the contig end).
    Returns (margin_list, deepvariant_list) with the reference's tuple layouts (:449, :519)."""
    margin, dv = [], []
    for i in range(len(positions)):
        pos = int(positions[i]); depth = int(depths[i]); key = keys[i]; freq = int(freqs[i])
        pb = np.asarray(probs[i], dtype=np.float64)
        ref_base = fetch(contig, pos, pos + 1).upper()
        up = fetch(contig, pos, pos + 10).upper()
        down = fetch(contig, max(0, pos - 10), pos).upper()
        full = (down + up).upper()
        hp = repeat_annotation_hp(full)
        pi = len(down)
        lo, hi = max(0, pi - 5), min(len(hp), pi + 4)
        in_repeat = max(hp[lo:hi]) >= 5                                            # :400-406
        if ref_base not in ("A", "C", "G", "T"):                                    # :408
            continue
        g = int(np.argmax(pb))                                                      # :411
        genotype = [0, 0] if g == 0 else ([0, 1] if g == 1 else [1, 1])
        pv = pb[g]
        t, allele = key[0], key[1:]
        valid = all(b in "ACGT" for b in allele)
        if valid and t == "1" and g != 0:                                           # :427-445
            margin.append((contig, pos, pos + 1, ref_base, [allele], genotype, depth, [freq], pv, pb))
        alts, sup, ref_allele, non_alts = [], [], ref_base, []
        if valid:
            vaf = float(freq) / float(depth)
            na = max(pb[1], pb[2])
            non_alts.append(na)
            if t == "1":
                if (not in_repeat and na >= options["snp_p_value"]) or (in_repeat and na >= options["snp_p_value_in_lc"]) \
                        or (0 < options["report_snp_above_freq"] <= vaf):
                    alts.append(allele); sup.append(freq)
            elif t == "2":
                if (not in_repeat and na >= options["insert_p_value"]) or (in_repeat and na >= options["insert_p_value_in_lc"]) \
                        or (0 < options["report_indel_above_freq"] <= vaf):
                    alts.append(allele); sup.append(freq)
            elif t == "3":
                if (not in_repeat and na >= options["delete_p_value"]) or (in_repeat and na >= options["delete_p_value_in_lc"]):
                    alts.append(ref_allele); ref_allele = allele; sup.append(freq)  # :497-505: ref/alt swap
                elif 0 < options["report_indel_above_freq"] <= vaf:
                    alts.append(allele); sup.append(freq)                           # :506-508 (no swap in the reference)
        if alts:
            dv.append((contig, pos, pos + len(ref_allele), ref_allele, alts, genotype, depth, sup, pv, pb, non_alts, in_repeat))
    return margin, dv
