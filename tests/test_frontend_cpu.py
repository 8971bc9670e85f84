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


def test_prediction_store_writers(tmp_path):
    """frontend.write_*_predictions lay the calls out the way the reference's predict loops do (DataStorePredict layouts)."""
    from pepper_b200.datastore import VariantPredictionStore, PolishPredictionStore
    from pepper_b200.frontend import write_variant_predictions, write_polish_predictions
    from pepper_b200.pipeline import VariantCalls, PolishCalls
    n = 1030
    keys = np.zeros((n, 64), dtype=np.uint8)
    keys[:, 0], keys[:, 1] = ord("1"), ord("T")
    vc = VariantCalls(np.arange(n, dtype=np.int64), np.full(n, 30, np.uint8), np.full(n, 9, np.uint8), keys, np.zeros(n, np.int32),
                      np.tile(np.array([[0.1, 0.2, 0.7]], np.float32), (n, 1)))
    with VariantPredictionStore(str(tmp_path / "v.hdf"), "w", backend="npz") as st:
        assert write_variant_predictions(st, "chr20", vc, batch_size=512) == 3
    st = VariantPredictionStore(str(tmp_path / "v.hdf"), "r", backend="npz")
    assert st.keys("predictions") == ["batch_0", "batch_1", "batch_2"]
    assert st.get("predictions/batch_2/positions").tolist() == list(range(1024, 1030))
    assert st.get("predictions/batch_0/base_prediction").shape == (512, 3) and st.get("predictions/batch_0/candidates")[0, 0] == "1T"
    pc = PolishCalls(np.ones((3, 1000), np.uint8), np.full((3, 1000), 20, np.uint8), np.tile(np.arange(1000, dtype=np.int64), (3, 1)),
                     np.zeros((3, 1000), np.int32), np.array([0, 0, 1], np.int32), np.array([0, 1, 0], np.int32))
    with PolishPredictionStore(str(tmp_path / "p.hdf"), "w", backend="npz") as st:
        write_polish_predictions(st, "ctg", pc, [(0, 1100), (900, 2100)])
    st = PolishPredictionStore(str(tmp_path / "p.hdf"), "r", backend="npz")
    assert st.keys("predictions/ctg") == ["ctg-0-1100", "ctg-900-2100"]
    assert st.get("predictions/ctg/ctg-0-1100/1/bases").shape == (1000,) and int(st.get("predictions/ctg/ctg-900-2100/contig_end")) == 2100
