"""Polish stitch: oracle vs the golden vector produced by the UNMODIFIED reference function
(tests/golden/make_golden_stitch.py), and the CUDA kernel vs both."""
import os
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold():
    return np.load(os.path.join(GOLD, "stitch_seed11.npz"))


def test_oracle_matches_reference_golden():
    from oracle import stitch as ostitch
    g = _gold()
    got = ostitch.stitch(g["bases"], g["position"], g["index"], g["image_region"], g["chunk_id"], g["region_starts"], g["region_ends"])
    assert got == str(g["consensus"])


def _many_chunks(seed):
    """one region with 12 chunks: chunk ids 10, 11 sort before 2..9 as strings (Stitch.py:50)."""
    rng = np.random.default_rng(seed)
    cols = [(p, k) for p in range(300, 300 + 9000) for k in range(1 + int(rng.random() < 0.3))]
    imgs, n, start, end, cid = [], len(cols), 0, 1000, 0
    while True:
        pos = np.full(1000, -1, np.int64); idx = np.full(1000, -1, np.int64)
        m = end - start
        pos[:m] = [c[0] for c in cols[start:end]]; idx[:m] = [c[1] for c in cols[start:end]]
        imgs.append((0, cid, pos, idx, rng.integers(0, 5, 1000).astype(np.uint8)))
        cid += 1
        if end == n:
            break
        start = end - 50
        end = min(n, start + 1000)
    return imgs


@pytest.mark.gpu
def test_cuda_stitch_golden_and_oracle():
    from oracle import stitch as ostitch
    from pepper_b200.polish import stitch
    g = _gold()
    got = stitch(g["bases"], g["position"], g["index"], g["image_region"], g["chunk_id"], g["region_starts"])
    assert got == str(g["consensus"])
    imgs = _many_chunks(5)
    assert len(imgs) >= 11
    b = np.stack([i[4] for i in imgs]); p = np.stack([i[2] for i in imgs]); x = np.stack([i[3] for i in imgs])
    r = np.zeros(len(imgs), np.int32); c = np.arange(len(imgs), dtype=np.int32)
    want = ostitch.stitch(b, p, x, r, c, [300], [9299])
    assert stitch(b, p, x, r, c, [300]) == want
    assert want and set(want) <= set("ACGT")


@pytest.mark.gpu
def test_polish_end_to_end_consensus(oracle_built):
    """make_images -> call_consensus -> stitch through the public API equals the oracle chain."""
    from oracle import nets, stitch as ostitch
    from pepper_b200 import synth
    from pepper_b200.pipeline import PolishCaller
    from pepper_b200.polish import PolishSummary, chunk_images, stitch
    state = nets.make_polish_weights(3)
    reads, regions = synth.make_polish_workload(4, 30, synth.ONT, seed=44)
    pc = PolishCaller(state)
    calls = pc.call(reads, regions)
    got = stitch(calls.bases, calls.position, calls.index, calls.image_region, calls.chunk_id, regions.col("ref_start"))
    w = oracle_built.polish_encode(reads, regions, "port")
    imgs, pos, idx, cids, regs = chunk_images(PolishSummary(w["image"], w["pos"], w["idx"], w["col_off"]))
    want = ostitch.stitch(calls.bases, pos, idx, regs, cids, regions.col("ref_start"), regions.col("ref_end"))
    assert got == want and len(got) > 3000
    pc.close()
