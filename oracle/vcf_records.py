"""TEST INFRASTRUCTURE ONLY (oracle).  Restatement of the site-level half of `pepper_variant find_candidates`:
  * the (contig, position) merge with (ref, alt) de-duplication, pepper_variant/modules/python/CandidateFinder.py:547-581;
  * VCFWriter.candidate_list_to_variant, pepper_variant/modules/python/VcfWriter.py:48-138;
  * the record assembly of VCFWriter.write_vcf_records (QUAL, filter, GT/AP/GQ/DP/AD/VAF/REP, output files), VcfWriter.py:140-218
on the tuple lists `small_chunk_stitch` returns (oracle/find_candidates.py).  Pinned against the UNMODIFIED reference
functions by tests/golden/make_golden_vcf.py (stand-in modules for h5py / PEPPER_VARIANT / pysam)."""
from __future__ import annotations

import math
from collections import defaultdict

import numpy as np


def merge_sites(records):
    """CandidateFinder.py:547-574: records sorted by (contig, position) (stable), then per site the first record of every
    (ref, first alt) pair is kept.  Returns (contigs in order of appearance, dict (contig, pos) -> list of records)."""
    records = sorted(records, key=lambda x: (x[0], x[1]))                               # :547-548
    sites, seen, contigs = defaultdict(list), defaultdict(list), []
    for c in records:
        if c[0] not in contigs:                                                          # :564-565
            contigs.append(c[0])
        ref, alt = c[3], c[4][0]
        if (ref, alt) in seen[(c[0], c[1])]:                                             # :568-569
            continue
        seen[(c[0], c[1])].append((ref, alt))
        sites[(c[0], c[1])].append(c)
    return contigs, sites


def candidate_list_to_variant(candidates, allowed_multiallelics: int):
    """VcfWriter.py:48-138."""
    candidates = sorted(candidates, key=lambda x: (x[5], x[8]), reverse=True)           # :49
    if len(candidates) > allowed_multiallelics:                                          # :50-51
        candidates = candidates[:allowed_multiallelics]
    max_ref_length, max_ref_allele = 0, ''
    for c in candidates:                                                                 # :56-60
        if len(c[3]) > max_ref_length:
            max_ref_length, max_ref_allele = len(c[3]), c[3]
    norm = []
    for c in candidates:                                                                 # :62-74
        contig, ref_start, ref_end, ref_allele, alt_allele = c[0], c[1], c[2], c[3], c[4]
        need = max_ref_length - len(ref_allele) if len(ref_allele) < max_ref_length else 0
        if need > 0:
            suffix = max_ref_allele[-need:]
            ref_allele = ref_allele + suffix
            alt_allele = [a + suffix for a in alt_allele]
        norm.append((contig, ref_start, ref_end, ref_allele, alt_allele) + tuple(c[5:]))
    gt_qual = -1.0
    hp1, hp2 = [], []
    init = False
    site = dict(contig='', start=0, end=0, ref='', depth=0, alts=[], supports=[], non_alt=[], in_repeat=False)
    for i, c in enumerate(norm):                                                         # :92-128
        contig, ref_start, ref_end, ref_allele, alt_allele, genotype, depth, support, gp, predictions, non_alt, in_repeat = c
        site["in_repeat"] = in_repeat or site["in_repeat"]
        pg = int(np.argmax(predictions))
        if pg != 0:
            gt_qual = predictions[pg] if gt_qual < 0 else min(gt_qual, predictions[pg])
        elif gt_qual < 0:
            gt_qual = max(predictions[1], predictions[2])
        if not init:
            site.update(contig=contig, start=ref_start, end=ref_start + len(ref_allele), ref=ref_allele, depth=depth)
            init = True
        site["depth"] = min(site["depth"], depth)
        site["alts"].append(alt_allele[0])
        site["supports"].append(support[0])
        site["non_alt"].extend(non_alt)
        if pg == 1:
            hp1.append(i + 1)
        elif pg == 2:
            hp1.append(i + 1); hp2.append(i + 1)
    if 0 < len(hp1) + len(hp2) <= 2:                                                     # :130-135
        gt = hp1 + hp2
        if len(gt) == 1:
            gt = [0, gt[0]]
    else:
        gt = [0, 0]
    return (site["contig"], site["start"], site["end"], site["ref"], site["alts"], gt, site["depth"], site["supports"], gt_qual,
            site["non_alt"], site["in_repeat"])


def vcf_records(sites: dict, opt: dict):
    """VcfWriter.py:140-218 without pysam: one dict per record written, `files` = the VCFs it goes to
    (full always; pepper | variant_calling + variant_calling_snp / _indel)."""
    out = []
    last_position = -1
    for contig, position in sorted(sites):                                               # :144
        contig, ref_start, ref_end, ref_seq, alleles, genotype, depth, support, gp, non_alt, in_repeat = \
            candidate_list_to_variant(sites[(contig, position)], opt["allowed_multiallelics"])
        if len(alleles) <= 0:
            continue
        if ref_start == last_position:                                                   # :151 (not reset between contigs)
            continue
        max_alt_len = max(len(ref_seq), max(len(x) for x in alleles))
        last_position = ref_start
        alleles = tuple([ref_seq]) + tuple(alleles)
        qual = max(1, int(-10 * math.log10(max(0.000000001, 1.0 - gp))))                 # :156
        is_snp = max_alt_len == 1
        if is_snp:                                                                       # :160-171
            failed = (not in_repeat and qual <= opt["snp_q_cutoff"]) or (in_repeat and qual <= opt["snp_q_cutoff_in_lc"])
        else:
            failed = (not in_repeat and qual <= opt["indel_q_cutoff"]) or (in_repeat and qual <= opt["indel_q_cutoff_in_lc"])
        selected = genotype == [0, 0] or failed                                          # :175-177
        vafs = [round(ad / max(1, depth), 3) for ad in support]                          # :179
        files = ["full"] + (["variant_calling_snp" if is_snp else "variant_calling_indel", "variant_calling"] if selected else ["pepper"])
        out.append(dict(contig=str(contig), start=ref_start, stop=ref_end, qual=qual, filter="refCall" if genotype == [0, 0] else "PASS",
                        alleles=alleles, GT=genotype, AP=list(non_alt), GQ=qual, DP=depth, AD=list(support), VAF=vafs,
                        REP="1" if in_repeat else "0", files=files))
    return out
