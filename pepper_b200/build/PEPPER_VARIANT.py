"""Drop-in for the reference's pybind11 module ``PEPPER_VARIANT`` (pepper_variant/modules/cpp/pybind_api.h:23-278),
restricted to the classes on the hot path.  Same class names, constructor arguments, method names and attribute
names; the work is done by libpepper_b200 on the GPU (one region per call here — the batched API is
``pepper_b200.variant.VariantEncoder``).  Usage in the reference's AlignmentSummarizer.py:1 :

    from pepper_b200.build import PEPPER_VARIANT        # instead of: from pepper_variant.build import PEPPER_VARIANT
"""
from __future__ import annotations

import numpy as np

from .. import synth
from ..variant import VariantEncoder


class CigarOp:                       # cigar.h:30
    def __init__(self, operation: int = -1, length: int = 0):
        self.operation = operation
        self.length = length


class type_read_flags:               # read.h:13
    def __init__(self):
        for f in ("is_paired", "is_proper_pair", "is_unmapped", "is_mate_unmapped", "is_reverse", "is_mate_is_reverse",
                  "is_read1", "is_read2", "is_secondary", "is_qc_failed", "is_duplicate", "is_supplementary"):
            setattr(self, f, False)


class type_read:                     # read.h:60
    def __init__(self):
        self.pos = 0
        self.pos_end = 0
        self.query_name = ""
        self.flags = type_read_flags()
        self.sequence = ""
        self.cigar_tuples = []
        self.bad_indicies = []
        self.mapping_quality = 0
        self.base_qualities = []
        self.read_id = 0
        self.hp_tag = 0


class CandidateImageSummary:         # region_summary.h:88 (picklable: it crosses ProcessPool boundaries)
    def __init__(self, contig="", position=0, depth=0, candidates=None, candidate_frequency=None, image_matrix=None,
                 base_label=0, type_label=0):
        self.contig = contig
        self.position = position
        self.depth = depth
        self.candidates = candidates or []
        self.candidate_frequency = candidate_frequency or []
        self.image_matrix = image_matrix
        self.base_label = base_label
        self.type_label = type_label


def _reads_to_batch(reads) -> synth.ReadBatch:
    recs = []
    for r in reads:
        recs.append(dict(pos=int(r.pos), seq=r.sequence, qual=[int(q) for q in r.base_qualities],
                         cigar=[(int(c.operation), int(c.length)) for c in r.cigar_tuples],
                         reverse=bool(r.flags.is_reverse), mapq=int(r.mapping_quality)))
    return synth.make_batch(recs)


_encoder = None


def _enc() -> VariantEncoder:
    global _encoder
    if _encoder is None:
        _encoder = VariantEncoder(0)
    return _encoder


class RegionalSummaryGenerator:      # region_summary.h:138
    def __init__(self, contig: str, region_start: int, region_end: int, reference_sequence: str):
        self.contig = contig
        self.ref_start = int(region_start)
        self.ref_end = int(region_end)
        self.reference_sequence = reference_sequence
        n = self.ref_end - self.ref_start + 1
        self.max_observed_insert = [0] * n              # GENERATE_INDELS = false, region_summary.h:50
        self.cumulative_observed_insert = [0] * n
        self.positions = []
        self.index = []
        self.total_observered_insert_bases = 0

    def generate_max_insert_summary(self, reads):
        n = self.ref_end - self.ref_start + 1
        self.positions = [self.ref_start + i for i in range(n)]    # region_summary.cpp:69-96 with no inserts
        self.index = [0] * n

    def generate_summary(self, reads, min_snp_baseq, min_indel_baseq, snp_freq_threshold, insert_freq_threshold,
                         delete_freq_threshold, min_coverage_threshold, snp_candidate_freq_threshold,
                         indel_candidate_freq_threshold, candidate_support_threshold, skip_indels, candidate_region_start,
                         candidate_region_end, candidate_window_size, feature_size, train_mode):
        if train_mode:
            raise NotImplementedError("train_mode (label generation) is outside the inference hot path")
        if candidate_window_size != 32 or feature_size != 26:
            raise ValueError("libpepper_b200 is specialised on CANDIDATE_WINDOW_SIZE=32, IMAGE_HEIGHT=26")
        batch = _reads_to_batch(reads)
        ref = np.frombuffer(self.reference_sequence.encode(), dtype=np.uint8).copy()
        tab = np.array([[self.ref_start, self.ref_end, candidate_region_start, candidate_region_end, 0, ref.shape[0], 0,
                         batch.n_reads]], dtype=np.int64)
        params = dict(min_snp_baseq=min_snp_baseq, min_indel_baseq=min_indel_baseq, snp_freq_threshold=snp_freq_threshold,
                      insert_freq_threshold=insert_freq_threshold, delete_freq_threshold=delete_freq_threshold,
                      min_coverage_threshold=min_coverage_threshold, snp_candidate_freq_threshold=snp_candidate_freq_threshold,
                      indel_candidate_freq_threshold=indel_candidate_freq_threshold,
                      candidate_support_threshold=candidate_support_threshold, skip_indels=int(bool(skip_indels)))
        c = _enc().encode(batch, synth.RegionTable(tab, ref if ref.shape[0] else np.zeros(1, np.uint8)), params)
        keys = c.keys
        return [CandidateImageSummary(self.contig, int(c.positions[i]), int(c.depths[i]), [keys[i]], [int(c.freqs[i])],
                                      c.images[i], 0, 0) for i in range(len(c))]


# ------------------------------------------------------------------------------------------------ file handlers
_trimmer = None


def _trim():
    global _trimmer
    if _trimmer is None:
        from ..reads import ReadTrimmer
        _trimmer = ReadTrimmer(0)
    return _trimmer


def _batch_to_reads(b: synth.ReadBatch, min_baseq: int = 0) -> list:
    """pb_reads_t -> list[type_read] as BAM_handler::get_reads builds them (bam_handler.cpp:434-446)."""
    codes = b.codes()
    ascii_of = np.frombuffer(synth.NT16.encode(), dtype=np.uint8)
    out = []
    for i in range(b.n_reads):
        so, se = int(b.seq_off[i]), int(b.seq_off[i + 1])
        r = type_read()
        c = codes[so:se]
        q = b.qual[so:se]
        r.pos = int(b.pos[i])
        r.sequence = ascii_of[c].tobytes().decode()
        r.base_qualities = q.astype(int).tolist()
        ref_len = 0
        for w in b.cigar[b.cigar_off[i]:b.cigar_off[i + 1]]:
            op, ln = int(w & 15), int(w >> 4)
            r.cigar_tuples.append(CigarOp(op, ln))
            if op in (0, 2, 3, 7, 8):
                ref_len += ln
        r.pos_end = r.pos + ref_len
        r.flags.is_reverse = bool(b.flags[i] & 1)
        r.mapping_quality = int(b.mapq[i])
        bad = np.nonzero((q < min_baseq) | ~np.isin(c, synth.ACGT_CODES))[0].tolist()      # :216-222
        r.bad_indicies = bad + [len(r.sequence) + 1]                                       # :307
        r.read_id = i
        out.append(r)
    return out


class BAM_handler:                   # bam_handler.h:155; pybind_api.h BAM_handler(path).get_reads(...)
    """File I/O by the host BGZF/BAM/BAI reader of libpepper_b200, the trim of get_reads on the GPU.  Differences from the
    reference objects: query_name is empty and hp_tag is 0 (aux fields are not materialised; no encoder on the hot path
    reads them).  The fast path keeps the reads in HBM: pepper_b200.reads.ReadTrimmer + pepper_b200.bamio.BamReader."""

    def __init__(self, path: str):
        from ..bamio import BamReader
        self._reader = BamReader(path)

    def get_chromosome_sequence_names(self):
        return self._reader.get_chromosome_sequence_names()

    def get_chromosome_sequence_names_with_length(self):
        return self._reader.get_chromosome_sequence_names_with_length()

    def get_sample_names(self):
        return self._reader.get_sample_names()

    def get_reads(self, chromosome: str, start: int, stop: int, include_supplementary: bool, min_mapq: int = 0,
                  min_baseq: int = 0):
        view = self._reader.fetch(chromosome, int(start), int(stop))
        t = _trim().get_reads(view, [(int(start), int(stop))], include_supplementary, min_mapq, min_baseq)
        return _batch_to_reads(t.to_host(), min_baseq)


class FASTA_handler:                 # fasta_handler.h; pybind_api.h FASTA_handler(path)
    def __init__(self, path: str):
        from ..bamio import FastaReader
        self._reader = FastaReader(path)

    def get_reference_sequence(self, region: str, start: int, stop: int) -> str:
        return self._reader.get_reference_sequence(region, int(start), int(stop))

    def get_chromosome_sequence_length(self, chromosome_name: str) -> int:
        return self._reader.get_chromosome_sequence_length(chromosome_name)

    def get_chromosome_names(self):
        return self._reader.get_chromosome_names()
