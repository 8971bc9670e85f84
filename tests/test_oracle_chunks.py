"""CPU: oracle/chunk_images.py against the fixture produced by the UNMODIFIED reference AlignmentSummarizer.chunk_images
(tests/golden/make_golden_chunks.py), and the product's host-side chunk_images against that oracle."""
import os
import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load():
    return np.load(os.path.join(GOLD, "chunks_seed13.npz"))


def test_oracle_matches_reference_fixture():
    from oracle import chunk_images as och
    g = _load()
    gi, gp, gk, gc, gr = och.chunk_images(g["image"], g["pos"], g["idx"], g["col_off"])
    assert np.array_equal(gp, g["position"]) and np.array_equal(gk, g["index"])
    assert np.array_equal(gc, g["chunk_id"]) and np.array_equal(gr, g["region"])
    assert np.array_equal(gi.astype(np.int64).sum(axis=(1, 2)), g["image_sum"])
    # chunk table: valid prefix comes from [start, start + nvalid), the rest is zero / (-1, -1)
    for k in range(gc.shape[0]):
        s, n = int(g["start"][k]), int(g["nvalid"][k])
        assert np.array_equal(gi[k, :n], g["image"][s:s + n]) and not gi[k, n:].any()
        assert np.all(gp[k, n:] == -1) and np.all(gk[k, n:] == -1)


def test_product_host_chunking_matches_oracle():
    from oracle import chunk_images as och
    from pepper_b200.polish import PolishSummary, chunk_images
    g = _load()
    imgs, pos, idx, cids, regs = chunk_images(PolishSummary(g["image"], g["pos"], g["idx"], g["col_off"]))
    gi, gp, gk, gc, gr = och.chunk_images(g["image"], g["pos"], g["idx"], g["col_off"])
    assert np.array_equal(imgs, gi) and np.array_equal(pos, gp) and np.array_equal(idx, gk)
    assert np.array_equal(cids, gc) and np.array_equal(regs, gr)
