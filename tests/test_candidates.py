"""Candidate selection: oracle vs the golden records produced by the UNMODIFIED reference
small_chunk_stitch (tests/golden/make_golden_candidates.py), and the CUDA kernel vs both."""
import json
import os
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _norm(rec):
    return json.dumps([x.tolist() if hasattr(x, "tolist") else (bool(x) if isinstance(x, (bool, np.bool_)) else x) for x in rec], default=float)


def _load():
    g = np.load(os.path.join(GOLD, "candidates_seed17.npz"))
    return g, str(g["genome"]), json.loads(str(g["options"]))


def test_oracle_matches_reference_golden():
    from oracle import find_candidates as ofc
    g, genome, options = _load()
    m, d = ofc.select(options, "ctg", g["positions"], g["depths"], [str(k) for k in g["keys"]], g["freqs"], g["probs"],
                      lambda c, a, b: genome[max(0, a):max(0, b)])
    assert [_norm(r) for r in m] == list(g["margin"]) and [_norm(r) for r in d] == list(g["deepvariant"])


@pytest.mark.gpu
def test_cuda_matches_reference_golden():
    from pepper_b200 import synth
    from pepper_b200.candidates import find_candidates
    g, genome, options = _load()
    n = len(g["positions"])
    keys = np.zeros((n, 64), np.uint8)
    for i, k in enumerate(g["keys"]):
        b = str(k).encode(); keys[i, :len(b)] = np.frombuffer(b, np.uint8)
    ref = np.frombuffer(genome.encode(), np.uint8).copy()
    regions = synth.RegionTable(np.array([[0, len(genome) - 1, 0, len(genome) - 1, 0, len(genome), 0, 0]], np.int64), ref)
    # the golden probabilities are float64; the pipeline carries float32 -> compare on float32-rounded inputs both ways
    m, d = find_candidates("ctg", g["positions"], np.zeros(n, np.int32), g["depths"], g["freqs"], keys, g["probs"].astype(np.float32), regions, options)
    want_m = [json.loads(r) for r in g["margin"]]
    want_d = [json.loads(r) for r in g["deepvariant"]]
    assert [(r[1], r[3], r[4], r[5], r[6], r[7]) for r in m] == [(r[1], r[3], r[4], r[5], r[6], r[7]) for r in want_m]
    assert [(r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[11]) for r in d] == [(r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[11]) for r in want_d]
    assert np.allclose([r[8] for r in d], [r[8] for r in want_d], atol=1e-6)


@pytest.mark.gpu
def test_cuda_on_pipeline_output(oracle_built):
    """encoder -> network -> selection on the GPU against the oracle chain on the same records."""
    from oracle import nets, find_candidates as ofc
    from pepper_b200 import synth
    from pepper_b200.pipeline import VariantCaller
    from pepper_b200.candidates import find_candidates, ONT_OPTIONS
    reads, regions = synth.make_variant_workload(2, 6000, 30, synth.ONT, seed=51)
    caller = VariantCaller(nets.make_variant_weights(2))
    calls = caller.call(reads, regions, synth.ont_params())
    # one contig: the regions' reference strings are slices of one genome
    start = int(regions.table[0, 0]); end = int(regions.table[-1, 1])
    genome = np.zeros(end - start + 1, np.uint8)
    for r in range(regions.n_regions):
        t = regions.table[r]
        genome[t[0] - start:t[0] - start + t[5]] = regions.ref[t[4]:t[4] + t[5]]
    gs = bytes(genome).decode()

    def fetch(c, a, b):
        return gs[max(0, a - start):max(0, b - start)]
    opts = dict(ONT_OPTIONS); opts["report_indel_above_freq"] = 0.5
    m, d = find_candidates("chr20", calls.positions, calls.region_of, calls.depths, calls.freqs, calls.keys_raw, calls.probs, regions, opts)
    wm, wd = ofc.select(opts, "chr20", calls.positions, calls.depths, calls.keys, calls.freqs, calls.probs, fetch)
    assert [_norm(r) for r in m] == [_norm(r) for r in wm]
    assert [_norm(r) for r in d] == [_norm(r) for r in wd]
    assert len(d) > 20
    caller.close()
