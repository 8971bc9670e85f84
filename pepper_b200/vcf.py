"""Site-level half of `pepper_variant find_candidates` on the arrays the GPU path produces:

  * merge of the selected records per (contig, position) with the (ref, alt) de-duplication that removes the records adjacent
    intervals emit twice at their shared boundary      == CandidateFinder.py:547-581
  * per site: order by (genotype, genotype probability) descending, keep `allowed_multiallelics`, extend every allele to the
    longest reference allele, genotype / GQ / depth / supports                == VCFWriter.candidate_list_to_variant, VcfWriter.py:48-138
  * QUAL = max(1, int(-10 log10(max(1e-9, 1 - p)))), quality cut-offs, filter, which VCF a record goes to
                                                                              == VCFWriter.write_vcf_records, VcfWriter.py:140-218

The per-record decisions (genotype, repeat context, Margin / DeepVariant lists, delete ref/alt swap) come from the CUDA kernel
behind `candidates.select_flags`; everything here is array code over those flags (stable sort, unique, segmented min / any /
count via reduceat) plus one pass that formats the records.  Output: one dict per VCF record with the keyword arguments the
reference hands to pysam's `new_record` and `files`, the VCFs it is written to — a pysam writer consumes them unchanged."""
from __future__ import annotations

import math

import numpy as np

from .candidates import F_DV, F_REPEAT, F_SWAP, select_flags
from .synth import RegionTable

# --ont_r9_guppy5_sup defaults of find_candidates, pepper_variant/modules/argparse/SetParameters.py:40-65
VCF_OPTIONS_ONT = dict(allowed_multiallelics=4, snp_q_cutoff=20, indel_q_cutoff=15, snp_q_cutoff_in_lc=20, indel_q_cutoff_in_lc=10)
GT_OF = ([0, 0], [0, 1], [1, 1])


def _strlen(a: np.ndarray) -> np.ndarray:
    return np.char.str_len(a).astype(np.int64)


def assemble_sites(contig: str, positions, region_of, depths, freqs, keys_raw, probs, flags, geno, regions: RegionTable,
                   options: dict) -> list[dict]:
    """Records of ONE contig (arrays in the order the encoder produced them) -> VCF record dicts in position order."""
    sel = np.flatnonzero(np.asarray(flags) & F_DV)
    if sel.shape[0] == 0:
        return []
    pos = np.asarray(positions, dtype=np.int64)[sel]
    reg = np.asarray(region_of, dtype=np.int64)[sel]
    fl = np.asarray(flags)[sel]
    g = np.asarray(geno, dtype=np.int64)[sel]
    pb = np.asarray(probs, dtype=np.float32)[sel].astype(np.float64)
    dep = np.asarray(depths, dtype=np.int64)[sel]
    sup = np.asarray(freqs, dtype=np.int64)[sel]
    tab = regions.table
    ref_base = np.char.upper(regions.ref[tab[reg, 4] + pos - tab[reg, 0]].astype(np.uint8).view("S1"))
    kr = np.ascontiguousarray(np.asarray(keys_raw, dtype=np.uint8)[sel][:, 1:63]).view("S62")[:, 0]     # the allele = key[1:]
    swap = (fl & F_SWAP) != 0
    ref = np.where(swap, kr, ref_base.astype("S62"))
    alt = np.where(swap, ref_base.astype("S62"), kr)
    pv = pb[np.arange(pb.shape[0]), g]
    non_alt = np.maximum(pb[:, 1], pb[:, 2])
    rep = (fl & F_REPEAT) != 0

    # ---- sorted by position (stable), first record of every (position, ref, alt)                      CandidateFinder.py:547-574
    o = np.argsort(pos, kind="stable")
    st = np.zeros(o.shape[0], dtype=[("pos", np.int64), ("ref", "S62"), ("alt", "S62")])
    st["pos"], st["ref"], st["alt"] = pos[o], ref[o], alt[o]
    _, first = np.unique(st, return_index=True)
    o = o[np.sort(first)]
    pos, ref, alt, g, pv, pb, dep, sup, non_alt, rep = (a[o] for a in (pos, ref, alt, g, pv, pb, dep, sup, non_alt, rep))
    n = pos.shape[0]
    site = np.concatenate([[0], np.cumsum(pos[1:] != pos[:-1])])

    # ---- per site: by (genotype, probability) descending, ties in arrival order; keep allowed_multiallelics   VcfWriter.py:49-51
    o = np.lexsort((np.arange(n), -pv, -g, site))
    pos, ref, alt, g, pv, pb, dep, sup, non_alt, rep, site = (a[o] for a in (pos, ref, alt, g, pv, pb, dep, sup, non_alt, rep, site))
    start_all = np.flatnonzero(np.concatenate([[True], site[1:] != site[:-1]]))
    rank = np.arange(n) - np.repeat(start_all, np.diff(np.concatenate([start_all, [n]])))
    keep = rank < int(options["allowed_multiallelics"])
    pos, ref, alt, g, pv, pb, dep, sup, non_alt, rep, site = (a[keep] for a in (pos, ref, alt, g, pv, pb, dep, sup, non_alt, rep, site))
    n = pos.shape[0]
    s0 = np.flatnonzero(np.concatenate([[True], site[1:] != site[:-1]]))
    cnt = np.diff(np.concatenate([s0, [n]]))

    # ---- segmented reductions                                                                         VcfWriter.py:92-135
    s_depth = np.minimum.reduceat(dep, s0)
    s_rep = np.maximum.reduceat(rep.astype(np.int8), s0) > 0
    nz_min = np.minimum.reduceat(np.where(g != 0, pv, np.inf), s0)
    gt_qual = np.where(g[s0] != 0, nz_min, non_alt[s0])          # first of a site has the largest genotype: non-zero ones come first
    n_hom = np.add.reduceat((g == 2).astype(np.int64), s0)
    n_het = np.add.reduceat((g == 1).astype(np.int64), s0)
    rlen, alen = _strlen(ref), _strlen(alt)
    max_rlen = np.maximum.reduceat(rlen, s0)
    ragged = np.flatnonzero(np.minimum.reduceat(rlen, s0) != max_rlen)
    ref_s = [r.decode() for r in ref]
    alt_s = [a.decode() for a in alt]
    for k in ragged:                                             # extend to the longest reference allele        VcfWriter.py:53-74
        a, b = int(s0[k]), int(s0[k] + cnt[k])
        longest = ref_s[a + int(np.argmax(rlen[a:b]))]
        for i in range(a, b):
            need = len(longest) - len(ref_s[i])
            if need > 0:
                ref_s[i] += longest[-need:]
                alt_s[i] += longest[-need:]
    # ---- records                                                                                      VcfWriter.py:140-218
    out = []
    last_position = -1
    snp_cut, indel_cut = options["snp_q_cutoff"], options["indel_q_cutoff"]
    snp_cut_lc, indel_cut_lc = options["snp_q_cutoff_in_lc"], options["indel_q_cutoff_in_lc"]
    for k in range(s0.shape[0]):
        a, b = int(s0[k]), int(s0[k] + cnt[k])
        ref_start = int(pos[a])
        if ref_start == last_position:
            continue
        last_position = ref_start
        ref_seq = ref_s[a]
        alleles = alt_s[a:b]
        nh, nm = int(n_het[k]), int(n_hom[k])
        if nm == 1 and nh == 0:
            gt = [1, 1]
        elif nm == 0 and nh == 1:
            gt = [0, 1]
        elif nm == 0 and nh == 2:
            gt = [1, 2]
        else:
            gt = [0, 0]
        depth = int(s_depth[k])
        gp = float(gt_qual[k])
        qual = max(1, int(-10 * math.log10(max(0.000000001, 1.0 - gp))))
        in_rep = bool(s_rep[k])
        is_snp = max(len(ref_seq), max(len(x) for x in alleles)) == 1
        if is_snp:
            failed = (not in_rep and qual <= snp_cut) or (in_rep and qual <= snp_cut_lc)
        else:
            failed = (not in_rep and qual <= indel_cut) or (in_rep and qual <= indel_cut_lc)
        selected = gt == [0, 0] or failed
        ad = [int(x) for x in sup[a:b]]
        files = ["full"] + (["variant_calling_snp" if is_snp else "variant_calling_indel", "variant_calling"] if selected else ["pepper"])
        out.append(dict(contig=str(contig), start=ref_start, stop=ref_start + len(ref_seq), qual=qual,
                        filter="refCall" if gt == [0, 0] else "PASS", alleles=(ref_seq,) + tuple(alleles), GT=gt,
                        AP=[float(x) for x in non_alt[a:b]], GQ=qual, DP=depth, AD=ad, VAF=[round(x / max(1, depth), 3) for x in ad],
                        REP="1" if in_rep else "0", files=files))
    return out


def find_site_records(contig: str, positions, region_of, depths, freqs, keys_raw, probs, regions: RegionTable, options: dict,
                      vcf_options: dict | None = None) -> list[dict]:
    """predictions -> VCF records of one contig: the CUDA per-record selection + the site assembly."""
    flags, geno = select_flags(positions, region_of, depths, freqs, keys_raw, probs, regions, options)
    return assemble_sites(contig, positions, region_of, depths, freqs, keys_raw, probs, flags, geno, regions, vcf_options or VCF_OPTIONS_ONT)


def format_vcf_line(rec: dict, sample_fields=("GT", "AP", "GQ", "DP", "AD", "VAF", "REP")) -> str:
    """One VCF 4.2 data line for a record dict (text form of what pysam writes; positions are 1-based in VCF)."""
    def fmt(v):
        if isinstance(v, (list, tuple)):
            return ",".join(fmt(x) for x in v)
        if isinstance(v, float):
            return ("%.6g" % v)
        return str(v)
    sample = [("/".join(str(x) for x in rec["GT"]))] + [fmt(rec[f]) for f in sample_fields[1:]]
    return "\t".join([rec["contig"], str(rec["start"] + 1), ".", rec["alleles"][0], ",".join(rec["alleles"][1:]), str(rec["qual"]),
                      rec["filter"], ".", ":".join(sample_fields), ":".join(sample)])
