"""CPU: the image / prediction stores keep the reference's group / dataset layout, dtypes and shapes
(pepper_variant DataStore.py:54-71, DataStorePredict.py:49-66; pepper DataStore.py:53-67, DataStorePredict.py:49-76)."""
import numpy as np

from pepper_b200 import datastore as ds


def test_variant_stores_roundtrip(tmp_path):
    f = str(tmp_path / "img")
    imgs = (np.arange(2 * 33 * 26) % 251 - 125).astype(np.int8).reshape(2, 33, 26)
    with ds.VariantImageStore(f, backend="npz") as s:
        s.write_summary("chr20_1000_2000", "chr20", [1001, 1500], [30, 125], ["1T", "2ACG"], [5, 12], imgs)
    r = ds._Store(f, mode="r", backend="npz")
    g = "summaries/chr20_1000_2000/"
    assert r.keys("summaries") == ["chr20_1000_2000"]
    assert r.get(g + "positions").dtype == np.int32 and r.get(g + "depths").dtype == np.uint8
    assert r.get(g + "images").dtype == np.int8 and r.get(g + "images").shape == (2, 33, 26)
    assert r.get(g + "candidates").shape == (2, 1) and r.get(g + "candidates")[1, 0] == "2ACG"
    assert r.get(g + "candidate_frequency").shape == (2, 1)
    p = str(tmp_path / "pred")
    with ds.VariantPredictionStore(p, backend="npz") as s:
        s.write_prediction(0, ["chr20", "chr20"], [1001, 1500], [30, 125], ["1T", "2ACG"], [5, 12],
                           np.array([[0.1, 0.8, 0.1], [0.9, 0.05, 0.05]], np.float32))
    r = ds._Store(p, mode="r", backend="npz")
    bp = r.get("predictions/batch_0/base_prediction")
    assert bp.dtype == np.float64 and bp.shape == (2, 3)


def test_polish_stores_roundtrip(tmp_path):
    f = str(tmp_path / "pimg")
    with ds.PolishImageStore(f, backend="npz") as s:
        s.write_summary("ctg1", 0, 1100, 1, np.zeros((1000, 10), np.uint8), np.arange(1000), np.zeros(1000, np.int64))
    r = ds._Store(f, mode="r", backend="npz")
    g = "summaries/ctg1_0_1100_1/"
    assert sorted(k for k in r.data if k.startswith(g)) == sorted(g + k for k in
                                                                   ("image", "label", "position", "index", "contig", "region_start", "region_end", "chunk_id"))
    assert r.get(g + "position").dtype == np.int64 and r.get(g + "image").shape == (1000, 10)
    p = str(tmp_path / "ppred")
    with ds.PolishPredictionStore(p, backend="npz") as s:
        s.write_prediction("ctg1", 0, 1100, 0, np.arange(1000), np.zeros(1000), np.ones(1000), np.full(1000, 30))
        s.write_prediction("ctg1", 0, 1100, 1, np.arange(1000), np.zeros(1000), np.ones(1000), np.full(1000, 30))
    r = ds._Store(p, mode="r", backend="npz")
    assert r.keys("predictions/ctg1") == ["ctg1-0-1100"]
    assert set(r.keys("predictions/ctg1/ctg1-0-1100")) == {"contig_start", "contig_end", "0", "1"}
    assert r.get("predictions/ctg1/ctg1-0-1100/1/bases").dtype == np.uint8
