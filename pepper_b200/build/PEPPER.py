"""Drop-in for the reference's pybind11 module ``PEPPER`` (pepper/modules/headers/pybind_api.h:17-118), restricted to
the classes on the hot path: ``SummaryGenerator`` runs on the GPU through libpepper_b200.  Usage in the reference's
pepper/modules/python/AlignmentSummarizer.py:1 :

    from pepper_b200.build import PEPPER                 # instead of: from pepper.build import PEPPER
"""
from __future__ import annotations

import numpy as np

from .. import synth
from ..polish import PolishEncoder
from .PEPPER_VARIANT import (CigarOp, type_read, type_read_flags, _reads_to_batch,  # noqa: F401  (same structs)
                             BAM_handler, FASTA_handler)                            # noqa: F401  (same handlers)

_encoder = None


def _enc() -> PolishEncoder:
    global _encoder
    if _encoder is None:
        _encoder = PolishEncoder(0)
    return _encoder


class SummaryGenerator:              # pepper/modules/headers/pileup_summary/summary_generator.h:20
    def __init__(self, reference_sequence: str, chromosome_name: str, ref_start: int, ref_end: int):
        self.reference_sequence = reference_sequence
        self.chromosome_name = chromosome_name
        self.ref_start = int(ref_start)
        self.ref_end = int(ref_end)
        self.image = []
        self.labels = []
        self.genomic_pos = []
        self.bad_label_positions = []

    def generate_summary(self, reads, start_pos: int, end_pos: int):
        if int(start_pos) != self.ref_start or int(end_pos) != self.ref_end:
            raise ValueError("the reference always calls generate_summary with the constructor's start/end "
                             "(AlignmentSummarizer.py:341-348); other ranges are not supported")
        batch = _reads_to_batch(reads)
        tab = np.array([[self.ref_start, self.ref_end, self.ref_start, self.ref_end, 0, 0, 0, batch.n_reads]], dtype=np.int64)
        s = _enc().encode(batch, synth.RegionTable(tab, np.zeros(1, np.uint8)))
        self.image = s.image                     # uint8 [cols,10] (the reference returns list[list[uint8]])
        self.genomic_pos = list(zip(s.pos.tolist(), s.idx.tolist()))


# ------------------------------------------------------------------------------------------------ realignment
_realigner = None


class ReadAligner:                   # realignment/simple_aligner.h:37; pybind_api.h ReadAligner(ref_start, ref_end, ref_seq)
    """align_reads_to_reference on the GPU (pepper_b200/csrc/realign.cu), bit-identical to the vendored SSW library.
    One region per call here; the batched path is pepper_b200.realign.Realigner."""

    def __init__(self, ref_start: int, ref_end: int, ref_seq: str):
        self.region_start = int(ref_start)
        self.region_end = int(ref_end)
        self.reference_sequence = ref_seq

    def align_reads_to_reference(self, reads):
        global _realigner
        from ..realign import Realigner
        if _realigner is None:
            _realigner = Realigner(0)
        kept = [r for r in reads if int(r.pos) >= self.region_start]        # simple_aligner.cpp:73-77 (the rest is dropped)
        if not kept:
            return []
        batch = _reads_to_batch(kept)
        ref = np.frombuffer(self.reference_sequence.encode(), dtype=np.uint8).copy()
        tab = np.array([[self.region_start, self.region_end, self.region_start, self.region_end, 0, ref.shape[0], 0, batch.n_reads]],
                       dtype=np.int64)
        out = _realigner.realign(batch, synth.RegionTable(tab, ref if ref.shape[0] else np.zeros(1, np.uint8)))
        res = []
        for i, r in enumerate(kept):
            new_cig = out.cigar[out.cigar_off[i]:out.cigar_off[i + 1]]
            old_cig = batch.cigar[batch.cigar_off[i]:batch.cigar_off[i + 1]]
            if out.pos[i] != batch.pos[i] or new_cig.shape[0] != old_cig.shape[0] or not np.array_equal(new_cig, old_cig):
                import copy
                r = copy.copy(r)
                r.cigar_tuples = [CigarOp(int(w & 15), int(w >> 4)) for w in new_cig]
                r.pos = int(out.pos[i])
                r.pos_end = r.pos + sum(int(w >> 4) for w in new_cig if int(w & 15) in (0, 2, 3, 7, 8)) - 1
            res.append(r)
        return res
