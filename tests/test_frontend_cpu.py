"""CPU: host logic of the file front end (interval tiling, reservoir sampling, region / reference tables)."""
import numpy as np

from pepper_b200 import synth, synth_files


def test_intervals_match_reference_tiling():
    from pepper_b200.frontend import polish_intervals, variant_intervals
    # pepper ImageGenerationUI.py:269-272 (max_size 1000, MIN_IMAGE_OVERLAP 100)
    assert polish_intervals(0, 2499) == [(0, 1100), (900, 2100), (1900, 2499)]
    assert polish_intervals(500, 1400) == [(500, 1400)]
    # pepper_variant ImageGenerationUI.py:307-316 (region_size 100000)
    assert variant_intervals(0, 250_000) == [(0, 100_000), (100_000, 200_000), (200_000, 250_000)]
    assert variant_intervals(10, 20, 100) == [(10, 20)]


def test_reference_table_is_one_fetch_with_clamped_lengths(tmp_path):
    """_FromFiles._ref_table: per-region strings are offsets into ONE faidx fetch; lengths clamp at the contig end like
    get_reference_sequence does."""
    from pepper_b200.bamio import FastaReader
    from pepper_b200.frontend import _FromFiles
    genome = synth.make_reference(5000, 3)
    fa = str(tmp_path / "g.fa")
    synth_files.write_fasta(fa, [("c", genome)])

    class Stub(_FromFiles):
        def __init__(self):
            self.fasta = FastaReader(fa)
    rows = [[100, 1300, 100, 1300, 0, 0, 0, 0], [4000, 5100, 4000, 5100, 0, 0, 0, 0]]
    spans = [(100, 1321), (4000, 5121)]
    tab = Stub()._ref_table("c", rows, spans)
    for r, (a, b) in enumerate(spans):
        off, ln = int(tab.table[r, 4]), int(tab.table[r, 5])
        assert ln == min(b, 5000) - a
        assert np.array_equal(tab.ref[off:off + ln], genome[a:min(b, 5000)])


def test_reservoir_matches_numpy_stream_for_region_limits():
    from pepper_b200.reads import reservoir_select
    for total, allowed in [(1501, 1500), (6000, 5000)]:
        s = reservoir_select(total, allowed)
        assert s.shape[0] == allowed and len(set(s.tolist())) == allowed and s.max() < total
