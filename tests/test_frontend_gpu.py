"""GPU: the one-call file -> predictions front ends (pepper_b200/frontend.py) against the oracle chain: get_reads (+ reservoir
down-sampling) -> [realign] -> encoder, all oracle restatements, then the same network weights."""
import numpy as np
import pytest

from pepper_b200 import synth, synth_files

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("frontend_gpu")
    rec, genome = synth.simulate_contig_records(26000, 30, synth.ONT, 23)
    bam, fa = str(d / "f.bam"), str(d / "f.fa")
    synth_files.write_bam(bam, [("ctg", genome.shape[0])], {0: rec})
    synth_files.write_fasta(fa, [("ctg", genome)])
    return dict(bam=bam, fa=fa, rec=rec, genome=genome)


def oracle_reads(oracle, rec, queries, supp, mq, bq, max_reads, rate=1.0):
    from pepper_b200.reads import reservoir_select
    batches, counts = [], []
    for (s, e) in queries:
        b, _, _ = oracle.get_reads(rec, s, e, supp, mq, bq, impl="port")
        sel = reservoir_select(b.n_reads, int(min(max_reads, rate * b.n_reads)))
        if sel is not None:
            b = synth.take_reads(b, sel)
        batches.append(b)
        counts.append(b.n_reads)
    return synth.concat_batches(batches), np.array(counts)


def test_intervals_match_reference_tiling():
    from pepper_b200.frontend import polish_intervals, variant_intervals
    assert polish_intervals(0, 2499) == [(0, 1100), (900, 2100), (1900, 2499)]
    assert variant_intervals(0, 250_000) == [(0, 100_000), (100_000, 200_000), (200_000, 250_000)]


@pytest.mark.parametrize("max_reads", [5000, 25])
def test_variant_from_files(oracle_built, files, max_reads):
    from pepper_b200 import weights
    from pepper_b200.frontend import VariantFromFiles, variant_intervals
    params = synth.ont_params()
    iv = variant_intervals(2000, 24000, 8000)
    vf = VariantFromFiles(files["bam"], files["fa"], weights.random_variant_state(0))
    calls, table = vf.call("ctg", iv, params, max_reads=max_reads)
    queries = [(max(0, s - 100), e + 100) for s, e in iv]
    reads, counts = oracle_reads(oracle_built, files["rec"], queries, False, 0, int(params["min_snp_baseq"]), max_reads)
    assert np.array_equal(table.table[:, 7] - table.table[:, 6], counts)
    if max_reads == 25:
        assert (counts == 25).all()
    tab = table.table.copy()
    want = oracle_built.variant_encode(reads, synth.RegionTable(tab, table.ref), params, "port")
    assert calls.keys == want["keys"] and np.array_equal(calls.positions, want["positions"])
    assert np.array_equal(calls.images, oracle_built.images_to_int8(want["images"]))
    assert len(calls) > (5 if max_reads == 25 else 20)


def test_polish_from_files(oracle_built, files):
    from pepper_b200 import weights
    from pepper_b200.frontend import PolishFromFiles, polish_intervals
    from pepper_b200.pipeline import PolishCaller
    regs = polish_intervals(3000, 8000)
    state = weights.random_polish_state(0)
    pf = PolishFromFiles(files["bam"], files["fa"], state)
    calls, table = pf.call("ctg", regs, realign=True, max_reads=30)
    reads, counts = oracle_reads(oracle_built, files["rec"], regs, False, 0, 0, 30)
    assert np.array_equal(table.table[:, 7] - table.table[:, 6], counts) and counts.max() == 30
    otab = synth.RegionTable(table.table.copy(), table.ref)
    pos, off, cig = [], [0], []
    for r in range(len(regs)):
        row = otab.table[r]
        ref = otab.ref[int(row[4]):int(row[4] + row[5])].tobytes().decode()
        p_, _, co, c = oracle_built.realign(reads, int(row[6]), int(row[7]), int(row[0]), int(row[1]) + 20, ref, impl="port")
        pos.append(p_)
        cig.append(c)
        off.extend((co[1:] + off[-1]).tolist())
    realigned = synth.ReadBatch(np.concatenate(pos), reads.seq_off, np.array(off, dtype=np.int64), reads.flags, reads.mapq, reads.seq,
                                reads.qual, np.concatenate(cig))
    pc = PolishCaller(state)
    want = pc.call(realigned, otab)                    # encoder + GRU parity of this caller is covered by test_pipeline_gpu / smoke
    assert np.array_equal(calls.position, want.position) and np.array_equal(calls.chunk_id, want.chunk_id)
    assert np.array_equal(calls.bases, want.bases) and np.array_equal(calls.phred, want.phred)
    # and the images behind it: the oracle polish encoder on the oracle-realigned reads
    from pepper_b200.polish import PolishEncoder
    s_gpu = PolishEncoder(0).encode(realigned, otab)
    s_or = oracle_built.polish_encode(realigned, otab, "port")
    assert np.array_equal(s_gpu.image, s_or["image"])
    pc.close()


def test_variant_batches_equal_single_call(files):
    from pepper_b200 import weights
    from pepper_b200.frontend import VariantFromFiles, variant_intervals
    params = synth.ont_params()
    iv = variant_intervals(1000, 25000, 4000)
    vf = VariantFromFiles(files["bam"], files["fa"], weights.random_variant_state(0))
    whole, _ = vf.call("ctg", iv, params)
    parts = list(vf.call_batches("ctg", iv, params, batch=2))
    assert len(parts) == 3
    assert np.array_equal(np.concatenate([p[0].positions for p in parts]), whole.positions)
    assert sum((p[0].keys for p in parts), []) == whole.keys
    assert np.array_equal(np.concatenate([p[0].images for p in parts]), whole.images)
    assert np.abs(np.concatenate([p[0].probs for p in parts]) - whole.probs).max() < 1e-5


def test_variant_stream_equals_single_call(files):
    """call_stream (one streaming session over batches, GPU inflate prefetched by a helper thread) == one call, for the GPU
    and the host inflate paths."""
    from pepper_b200 import weights
    from pepper_b200.frontend import VariantFromFiles, variant_intervals
    params = synth.ont_params()
    iv = variant_intervals(1000, 25000, 3000)
    for gpu_inflate in (True, False):
        vf = VariantFromFiles(files["bam"], files["fa"], weights.random_variant_state(0), gpu_inflate=gpu_inflate)
        whole, _ = vf.call("ctg", iv, params)
        got = vf.call_stream("ctg", iv, params, batch=3, want_images=True)
        assert np.array_equal(got.positions, whole.positions) and got.keys == whole.keys
        assert np.array_equal(got.region_of, whole.region_of) and np.array_equal(got.images, whole.images)
        assert np.array_equal(got.probs, whole.probs)
        small = vf.call_stream("ctg", iv, params, batch=3, capacity=10)          # capacity retry restarts the session
        assert np.array_equal(small.probs, whole.probs)
        vf.close()


def test_file_source_through_distributed_caller_equals_call_stream(files):
    """The N-GPU from-files path at world size 1: DistributedVariantCaller over a VariantFileSource (groups of 3 intervals, fetched two
    ahead by the helper thread) writes the records of VariantFromFiles.call_stream, bit for bit; a second pass over the file
    (`replicas`, the weak-scaling job of bench.py) repeats them with shifted region ids."""
    from pepper_b200 import weights
    from pepper_b200.dist import DistributedVariantCaller, records_from_calls
    from pepper_b200.frontend import VariantFileSource, VariantFromFiles, variant_intervals
    params = synth.ont_params()
    iv = variant_intervals(1000, 25000, 3000)
    state = weights.random_variant_state(0)
    with VariantFromFiles(files["bam"], files["fa"], state) as vf:
        want = records_from_calls(vf.call_stream("ctg", iv, params, batch=3))
    assert want.shape[0] > 50
    for gpu_inflate in (True, False):
        dvc = DistributedVariantCaller(state, 0, capacity=4 * want.shape[0], group_regions=3)
        src = VariantFileSource(files["bam"], files["fa"], "ctg", iv, int(params["min_snp_baseq"]), gpu_inflate=gpu_inflate)
        assert dvc.run(src, None, params) == want.shape[0]
        assert np.array_equal(dvc.buffer.to_host(), want)
        assert dvc.run(src, None, params, replicas=2) == 2 * want.shape[0]
        twice = dvc.buffer.to_host()
        assert np.array_equal(twice[:want.shape[0]], want)
        tail = twice[want.shape[0]:].copy()
        tail["region"] -= len(iv)
        assert np.array_equal(tail, want)
        src.close()
        dvc.close()


def test_polish_contig_from_files_to_consensus(files):
    """files -> tiling -> get_reads -> realign -> encoder -> GRU -> stitch, in batches, equals the oracle stitch of the same calls
    and does not depend on the batch size."""
    from oracle import stitch as ostitch
    from pepper_b200 import weights
    from pepper_b200.frontend import PolishFromFiles
    pf = PolishFromFiles(files["bam"], files["fa"], weights.random_polish_state(1))
    seq, calls, regs = pf.polish_contig("ctg", 2000, 9999, batch=3, return_calls=True)
    starts = np.array([r[0] for r in regs], dtype=np.int64)
    ends = np.array([r[1] for r in regs], dtype=np.int64)
    want = ostitch.stitch(calls.bases, calls.position, calls.index, calls.image_region, calls.chunk_id, starts, ends)
    assert seq == want and len(seq) > 6000
    assert pf.polish_contig("ctg", 2000, 9999, batch=100) == seq


def test_variant_files_to_vcf_records(oracle_built, files):
    """BAM + FASTA -> candidates -> LSTM -> candidate selection == the oracle selection on the same prediction records."""
    from oracle import nets, find_candidates as ofc
    from pepper_b200.candidates import ONT_OPTIONS
    from pepper_b200.frontend import VariantFromFiles, variant_intervals
    from tests.test_candidates import _norm
    params = synth.ont_params()
    iv = variant_intervals(2000, 24000, 8000)
    vf = VariantFromFiles(files["bam"], files["fa"], nets.make_variant_weights(2))
    opts = dict(ONT_OPTIONS); opts["report_indel_above_freq"] = 0.5
    m, d = vf.find_candidate_tuples("ctg", iv, params, opts)
    calls, _ = vf.call("ctg", iv, params, want_images=False)
    gs = files["genome"].tobytes().decode()
    wm, wd = ofc.select(opts, "ctg", calls.positions, calls.depths, calls.keys, calls.freqs, calls.probs, lambda c, a, b: gs[max(0, a):max(0, b)])
    assert [_norm(r) for r in m] == [_norm(r) for r in wm]
    assert [_norm(r) for r in d] == [_norm(r) for r in wd]
    assert len(d) > 10
    # the per-site records (site merge across the shared interval boundaries, dedup, QUAL, GT): oracle restatement of the
    # reference's merge + VCFWriter on the oracle's tuples
    from oracle import vcf_records as ovr
    from pepper_b200.vcf import VCF_OPTIONS_ONT
    from tests.test_vcf_records import _norm as _n2
    recs = vf.find_candidates("ctg", iv, params, opts)
    contigs, sites = ovr.merge_sites(wd)
    want = ovr.vcf_records(sites, VCF_OPTIONS_ONT)
    assert [_n2(r) for r in recs] == [_n2(r) for r in want] and len(recs) > 10
