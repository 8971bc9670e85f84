"""GPU: the fused public API (pb_*_call_host) equals encoder-then-network run separately and the oracle."""
import numpy as np
import pytest

from pepper_b200 import synth

pytestmark = pytest.mark.gpu


def test_variant_call_matches_separate_stages(oracle_built):
    from oracle import nets
    from pepper_b200.pipeline import VariantCaller
    state = nets.make_variant_weights(1)
    reads, regions = synth.make_variant_workload(3, 7000, 30, synth.ONT, seed=31)
    caller = VariantCaller(state)
    calls = caller.call(reads, regions, synth.ont_params(), want_images=True)
    want = oracle_built.variant_encode(reads, regions, synth.ont_params(), "port")
    assert calls.keys == want["keys"]
    assert np.array_equal(calls.images, oracle_built.images_to_int8(want["images"]))
    assert np.array_equal(calls.region_of, want["region_of"])
    probs = nets.variant_predict(state, calls.images, threads=8)
    assert np.abs(probs - calls.probs).max() < 1e-3
    sep = caller.net.predict(calls.images)
    assert np.array_equal(sep, calls.probs)          # same kernels, same order -> bitwise
    t = caller.timings()
    assert t["encode_ms"] > 0 and t["network_ms"] > 0
    small = caller.call(reads, regions, synth.ont_params(), capacity=5)
    assert np.array_equal(small.probs, calls.probs)
    caller.close()


def test_polish_call_matches_separate_stages(oracle_built):
    from oracle import nets
    from oracle import chunk_images as och          # pinned to the unmodified AlignmentSummarizer.chunk_images
    from pepper_b200.pipeline import PolishCaller
    state = nets.make_polish_weights(2)
    reads, regions = synth.make_polish_workload(6, 35, synth.ONT, seed=32)
    pc = PolishCaller(state)
    calls = pc.call(reads, regions)
    w = oracle_built.polish_encode(reads, regions, "port")
    imgs, pos, idx, cids, regs = och.chunk_images(w["image"], w["pos"], w["idx"], w["col_off"])
    assert len(set(cids.tolist())) > 1               # multi-chunk regions present (k_polish_chunk's overlap rule exercised)
    assert np.array_equal(calls.position, pos) and np.array_equal(calls.index.astype(np.int64), idx)
    assert np.array_equal(calls.chunk_id, cids) and np.array_equal(calls.image_region, regs)
    b, p = pc.net.predict(imgs)
    assert np.array_equal(b, calls.bases) and np.array_equal(p, calls.phred)
    wb, wp, wh, wa = nets.polish_predict(state, imgs, threads=8)
    srt = np.sort(wa, axis=2)
    clear = (srt[:, :, -1] - srt[:, :, -2]) > 1e-4
    assert np.array_equal(calls.bases[clear], wb[clear]) and clear.mean() > 0.99
    pc.close()


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_variant_call_pipelined_groups_match_single_shot(oracle_built):
    """> 96 regions: the host entry stages region groups on a copy stream while the previous group computes; the result
    must equal the encoder + network run in one shot (and the oracle's candidate list)."""
    from oracle import nets
    from pepper_b200.pipeline import VariantCaller
    state = nets.make_variant_weights(1)
    reads, regions = synth.make_variant_workload(6, 1500, 25, synth.ONT, seed=33)
    reads, regions = synth.tile_workload(reads, regions, 40)             # 240 regions -> 3 groups
    caller = VariantCaller(state)
    calls = caller.call(reads, regions, synth.ont_params(), want_images=True)
    enc = caller.enc.encode(reads, regions, synth.ont_params())
    assert calls.keys == enc.keys and np.array_equal(calls.positions, enc.positions)
    assert np.array_equal(calls.region_of, enc.region_of) and np.array_equal(calls.images, enc.images)
    assert np.array_equal(calls.probs, caller.net.predict(enc.images))
    small = caller.call(reads, regions, synth.ont_params(), capacity=7)   # capacity retry across groups
    assert np.array_equal(small.probs, calls.probs)
    one = synth.RegionTable(regions.table[:6].copy(), regions.ref)
    want = oracle_built.variant_encode(reads, one, synth.ont_params(), "port")
    assert calls.keys[:len(want["keys"])] == want["keys"]
    caller.close()
