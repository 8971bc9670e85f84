"""Candidate selection after inference (== small_chunk_stitch of pepper_variant/modules/python/CandidateFinder.py:356-530)
on the GPU: the kernel decides, per prediction record, whether it goes to the Margin (phasing) list and/or the DeepVariant
(re-genotyping) list, its genotype and repeat context; this module only assembles the reference's Python tuples from
those flags so that `VcfWriter.write_vcf_records` (pepper_variant VcfWriter.py:140) can consume them unchanged.

Option names are the reference's (`--snp_p_value`, ... ; platform defaults in SetParameters.py:38-65)."""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import _lib
from .abi import PbRegion, regions_array, ALLELE_STRIDE
from .synth import RegionTable


class PbCandidateOptions(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("snp_p_value", "insert_p_value", "delete_p_value", "snp_p_value_in_lc",
                                          "insert_p_value_in_lc", "delete_p_value_in_lc", "report_snp_above_freq",
                                          "report_indel_above_freq")]


ONT_OPTIONS = dict(snp_p_value=0.1, insert_p_value=0.1, delete_p_value=0.1, snp_p_value_in_lc=0.1, insert_p_value_in_lc=0.15,
                   delete_p_value_in_lc=0.1, report_snp_above_freq=0.0, report_indel_above_freq=0.0)   # SetParameters.py:42-63

F_MARGIN, F_DV, F_REPEAT, F_SWAP, F_REF_OK = 1, 2, 4, 8, 16


def select_flags(positions, region_of, depths, freqs, keys_raw, probs, regions: RegionTable, options: dict, stream: int = 0):
    """-> (flags uint8 [n], genotype uint8 [n]) from pb_variant_find_candidates_host."""
    _lib.require_gpu()
    L = _lib.lib()
    vp = C.c_void_p
    L.pb_variant_find_candidates_host.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, C.POINTER(PbRegion), C.c_int64, vp, C.c_int64,
                                                  C.POINTER(PbCandidateOptions), vp, vp, vp]
    n = int(len(positions))
    pos = np.ascontiguousarray(positions, dtype=np.int64)
    rof = np.ascontiguousarray(region_of, dtype=np.int32)
    dep = np.ascontiguousarray(depths, dtype=np.uint8)
    frq = np.ascontiguousarray(freqs, dtype=np.uint8)
    keys = np.ascontiguousarray(keys_raw, dtype=np.uint8).reshape(n, ALLELE_STRIDE)
    pr = np.ascontiguousarray(probs, dtype=np.float32).reshape(n, 3)
    regs, keep = regions_array(regions)
    ref = np.ascontiguousarray(regions.ref, dtype=np.uint8)
    o = PbCandidateOptions(**options)
    flags = np.zeros(n, dtype=np.uint8)
    geno = np.zeros(n, dtype=np.uint8)
    _lib.check(L.pb_variant_find_candidates_host(pos.ctypes.data, rof.ctypes.data, dep.ctypes.data, frq.ctypes.data, keys.ctypes.data,
                                                 pr.ctypes.data, n, regs, regions.n_regions, ref.ctypes.data, ref.shape[0], C.byref(o),
                                                 flags.ctypes.data, geno.ctypes.data, C.c_void_p(stream)),
               "pb_variant_find_candidates_host")
    return flags, geno


def find_candidates(contig: str, positions, region_of, depths, freqs, keys_raw, probs, regions: RegionTable, options: dict):
    """(margin_list, deepvariant_list) with the reference's tuple layouts (CandidateFinder.py:449, :519)."""
    flags, geno = select_flags(positions, region_of, depths, freqs, keys_raw, probs, regions, options)
    GT = ([0, 0], [0, 1], [1, 1])
    margin, dv = [], []
    tab = regions.table
    for i in np.nonzero(flags & (F_MARGIN | F_DV))[0]:
        key = bytes(keys_raw[i]).split(b"\0", 1)[0].decode()
        pos = int(positions[i]); r = int(region_of[i])
        ref_base = chr(regions.ref[int(tab[r, 4] + pos - tab[r, 0])]).upper()
        pb = np.asarray(probs[i], dtype=np.float64)
        g = int(geno[i])
        if flags[i] & F_MARGIN:
            margin.append((contig, pos, pos + 1, ref_base, [key[1:]], list(GT[g]), int(depths[i]), [int(freqs[i])], pb[g], pb))
        if flags[i] & F_DV:
            if flags[i] & F_SWAP:
                ref_allele, alts = key[1:], [ref_base]
            else:
                ref_allele, alts = ref_base, [key[1:]]
            dv.append((contig, pos, pos + len(ref_allele), ref_allele, alts, list(GT[g]), int(depths[i]), [int(freqs[i])], pb[g], pb,
                       [max(pb[1], pb[2])], bool(flags[i] & F_REPEAT)))
    return margin, dv
