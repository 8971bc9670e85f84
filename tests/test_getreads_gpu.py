"""GPU parity for row a2: the batched get_reads kernels (pepper_b200/csrc/get_reads.cu, through the C-ABI) against the
plain-C restatement of BAM_handler::get_reads, bit-exact; down-sampling; and the chain records -> get_reads -> encoder ->
network in HBM against the same chain fed with the oracle's trimmed reads."""
import numpy as np
import pytest

from pepper_b200 import synth
from tests import kats
from tests.golden import make_golden_getreads as gold
from tests.test_oracle_getreads import FIELDS, GOLD

pytestmark = pytest.mark.gpu


def same(a, b, ctx=""):
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert x.shape == y.shape and np.array_equal(x, y), (ctx, f)


@pytest.fixture(scope="module")
def trimmer():
    from pepper_b200.reads import ReadTrimmer
    t = ReadTrimmer(0)
    yield t
    t.close()


def oracle_batch(oracle, rec, queries, supp, mq, bq, select=None):
    outs, counts = [], []
    for qi, (s, e) in enumerate(queries):
        b, _, _ = oracle.get_reads(rec, s, e, supp, mq, bq, impl="port")
        counts.append(b.n_reads)
        if select is not None and select[qi] is not None:
            b = synth.take_reads(b, select[qi])
        outs.append(b)
    return synth.concat_batches(outs), np.array(counts)


def test_kats_single_queries(oracle_built, trimmer):
    for name, rec, queries in kats.getreads_kats():
        for (s, e, supp, mq, bq) in queries:
            want, _, _ = oracle_built.get_reads(rec, s, e, supp, mq, bq, impl="port")
            got = trimmer.get_reads(rec, [(s, e)], supp, mq, bq)
            assert got.total_reads[0] == want.n_reads, (name, s, e)
            same(got.to_host(), want, (name, s, e, supp, mq, bq))


def test_kats_batched(oracle_built, trimmer):
    name, rec, queries = kats.getreads_kats()[0]
    iv = [(q[0], q[1]) for q in queries]
    want, counts = oracle_batch(oracle_built, rec, iv, True, 0, 0)
    got = trimmer.get_reads(rec, iv, True, 0, 0)
    assert np.array_equal(got.total_reads, counts)
    assert np.array_equal(got.read_end - got.read_begin, counts)
    same(got.to_host(), want)


@pytest.mark.parametrize("seed,platform,n_iv", [(3, synth.ONT, 40), (4, synth.HIFI, 40)])
def test_synthetic_batched_device_records(oracle_built, trimmer, seed, platform, n_iv):
    from pepper_b200.reads import DeviceRecords
    start = 7000
    rec, _ = synth.simulate_contig_records(30000, 20, platform, seed, contig_start=start)
    rng = np.random.default_rng(seed)
    iv = []
    for _ in range(n_iv):
        s = int(rng.integers(start - 500, start + 30500))
        iv.append((s, s + int(rng.choice([1, 2, 50, 1201, 6000]))))
    for supp, mq in [(False, 0), (True, 10)]:
        want, counts = oracle_batch(oracle_built, rec, iv, supp, mq, 0)
        got = trimmer.get_reads(DeviceRecords(rec), iv, supp, mq, 0)
        assert np.array_equal(got.total_reads, counts)
        same(got.to_host(), want, (supp, mq))
        assert want.n_reads > 100


def test_golden(trimmer):
    g = np.load(GOLD)
    rec, _ = synth.simulate_contig_records(gold.CONTIG, gold.COV, synth.ONT, gold.SEED, contig_start=gold.START)
    for qi, (s, e, supp, mq, bq) in enumerate(gold.QUERIES):
        got = trimmer.get_reads(rec, [(s, e)], supp, mq, bq).to_host()
        for f in FIELDS:
            assert np.array_equal(getattr(got, f), g[f"q{qi}_{f}"]), (qi, f)


def test_polish_tiling_many_overlaps(oracle_built, trimmer):
    """1 kb polish regions against 10 kb reads: every record is returned by ~10 queries."""
    start = 0
    rec, _ = synth.simulate_contig_records(40000, 30, synth.ONT, 12, contig_start=start)
    iv = [(max(0, p - 100), p + 1100) for p in range(0, 40000, 1000)]
    want, counts = oracle_batch(oracle_built, rec, iv, False, 0, 0)
    got = trimmer.get_reads(rec, iv, False, 0, 0)
    assert np.array_equal(got.total_reads, counts)
    same(got.to_host(), want)


def test_downsampling_reservoir(oracle_built, trimmer):
    from pepper_b200.reads import reservoir_select
    rec, _ = synth.simulate_contig_records(12000, 60, synth.ONT, 5, contig_start=0)
    iv = [(1000, 3000), (4000, 4100), (6000, 9000), (20000, 21000)]
    _, counts = oracle_batch(oracle_built, rec, iv, False, 0, 0)
    max_reads = 40
    sel = [reservoir_select(int(c), int(min(max_reads, 1.0 * int(c)))) for c in counts]
    assert any(s is not None for s in sel) and any(s is None for s in sel)
    want, _ = oracle_batch(oracle_built, rec, iv, False, 0, 0, select=sel)
    got = trimmer.get_reads(rec, iv, False, 0, 0, max_reads=max_reads, downsample_rate=1.0)
    assert np.array_equal(got.total_reads, counts)
    assert np.array_equal(got.read_end - got.read_begin, [min(int(c), max_reads) for c in counts])
    same(got.to_host(), want)


def test_empty_and_errors(trimmer):
    from pepper_b200._lib import PepperB200Error
    rec = synth.make_records([])
    got = trimmer.get_reads(rec, [(0, 100)], False, 0, 0)
    assert got.total_reads.tolist() == [0] and got.to_host().n_reads == 0
    rec, _ = synth.simulate_contig_records(5000, 5, synth.HIFI, 1)
    got = trimmer.get_reads(rec, [], False, 0, 0)
    assert got.to_host().n_reads == 0
    bad = synth.make_records([dict(pos=50, seq="ACGT", cigar=[(0, 4)]), dict(pos=10, seq="ACGT", cigar=[(0, 4)])])
    with pytest.raises(PepperB200Error):
        trimmer.get_reads(bad, [(0, 100)], False, 0, 0)


def test_chain_records_to_variant_calls(oracle_built, trimmer):
    """records in HBM -> get_reads -> variant encoder -> LSTM, all on the device, equals the same caller fed with the
    oracle's trimmed reads (and its candidates equal the oracle encoder's on those reads)."""
    import torch
    from pepper_b200 import weights
    from pepper_b200.pipeline import VariantCaller, DeviceReads, FetchedReads
    from pepper_b200.reads import DeviceRecords
    start, size, safe = 50_000, 6000, 100
    rec, genome = synth.simulate_contig_records(3 * size + 2 * safe, 30, synth.ONT, 8, contig_start=start)
    rows, refs, iv, roff = [], [], [], 0
    for r in range(3):
        s = start + safe + r * size
        e = s + size
        rs, re_ = s - safe, e + safe                                   # AlignmentSummarizer.py:181-189
        iv.append((rs, re_))
        ref = genome[rs - start: re_ - start + 1]
        rows.append([rs, re_, s, e, roff, ref.shape[0], 0, 0])
        refs.append(ref)
        roff += ref.shape[0]
    regions = synth.RegionTable(np.array(rows, dtype=np.int64), np.concatenate(refs))
    params = synth.ont_params()
    got = trimmer.get_reads(DeviceRecords(rec), iv, False, 0, 0)
    fetched = FetchedReads(got, regions)
    # oracle chain: per-query port get_reads, concatenated, ranges into the table
    want_reads, counts = oracle_batch(oracle_built, rec, iv, False, 0, 0)
    tab = regions.table.copy()
    tab[:, 7] = np.cumsum(counts)
    tab[:, 6] = tab[:, 7] - counts
    oregions = synth.RegionTable(tab, regions.ref)
    vc = VariantCaller(weights.random_variant_state(0))
    dev = torch.device("cuda", 0)

    def outs(cap):
        return dict(images=torch.empty((cap, 33, 26), dtype=torch.int8, device=dev), positions=torch.empty(cap, dtype=torch.int64, device=dev),
                    depths=torch.empty(cap, dtype=torch.uint8, device=dev), freqs=torch.empty(cap, dtype=torch.uint8, device=dev),
                    keys=torch.empty((cap, 64), dtype=torch.uint8, device=dev), region_of=torch.empty(cap, dtype=torch.int32, device=dev),
                    probs=torch.empty((cap, 3), dtype=torch.float32, device=dev))
    a, b = outs(4000), outs(4000)
    na = vc.call_device(fetched, params, a)
    nb = vc.call_device(DeviceReads(want_reads, oregions), params, b)
    assert na == nb and na > 20
    for k in a:
        assert torch.equal(a[k][:na], b[k][:nb]), k
    o = oracle_built.variant_encode(want_reads, oregions, params, "port")
    assert np.array_equal(a["positions"][:na].cpu().numpy(), o["positions"])
    assert np.array_equal(a["images"][:na].cpu().numpy(), oracle_built.images_to_int8(o["images"]))
    vc.close()


def test_properties_large_batch(trimmer):
    """Size-independent properties at a size the oracle is not run at: every read starts on a match inside its query, CIGAR and
    sequence lengths agree, the batch is idempotent (trimming the trimmed reads again with the same queries changes nothing)."""
    from pepper_b200.reads import DeviceRecords
    start = 0
    rec, _ = synth.simulate_contig_records(300_000, 30, synth.ONT, 21, contig_start=start)
    iv = [(max(0, p - 100), p + 1100) for p in range(0, 300_000, 1000)]
    got = trimmer.get_reads(DeviceRecords(rec), iv, False, 0, 0)
    b = got.to_host()
    assert b.n_reads == int(got.total_reads.sum()) > 5000
    ops, lens = b.cigar & 15, (b.cigar >> 4).astype(np.int64)
    first = ops[b.cigar_off[:-1]]
    assert np.isin(first, [0, 7, 8]).all()
    read_cons = np.where(np.isin(ops, [0, 1, 4, 7, 8]), lens, 0)
    ref_cons = np.where(np.isin(ops, [0, 2, 3, 7, 8]), lens, 0)
    csum = np.concatenate([[0], np.cumsum(read_cons)])
    rsum = np.concatenate([[0], np.cumsum(ref_cons)])
    assert np.array_equal(csum[b.cigar_off[1:]] - csum[b.cigar_off[:-1]], np.diff(b.seq_off))
    span = rsum[b.cigar_off[1:]] - rsum[b.cigar_off[:-1]]
    qi = np.repeat(np.arange(len(iv)), got.read_end - got.read_begin)
    starts, stops = np.array([s for s, _ in iv]), np.array([e for _, e in iv])
    assert (b.pos >= starts[qi]).all() and (b.pos + span <= stops[qi] + 1).all()
    assert (lens > 0).all() and not np.isin(ops, [5, 6]).any()
    # idempotence, query by query on a sample
    for q in (0, 57, 150, 299):
        lo, hi = int(got.read_begin[q]), int(got.read_end[q])
        sub = synth.take_reads(b, np.arange(lo, hi))
        sub = synth.take_reads(sub, np.argsort(sub.pos, kind="stable"))      # trimmed starts need not be sorted
        as_rec = synth.RecordBatch(sub.pos, sub.seq_off, sub.cigar_off, (sub.flags.astype(np.uint16) & 1) * 16, sub.mapq, sub.seq, sub.qual, sub.cigar)
        again = trimmer.get_reads(as_rec, [iv[q]], False, 0, 0).to_host()
        same(again, sub, q)
