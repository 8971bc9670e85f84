"""GPU: the BGZF / DEFLATE inflate kernel and the device-side BAM fetch.

* k_bgzf_inflate against zlib on every block type: stored (level 0), fixed Huffman (Z_FIXED), dynamic Huffman at levels 1/6/9,
  several deflate blocks per stream (Z_FULL_FLUSH), overlapping LZ77 copies (runs), codes longer than the 10-bit lookup table,
  empty input — output byte-identical; corrupted streams are rejected with a status, not a crash.
* pb_bam_fetch_device against pb_bam_fetch (host zlib) on synthetic BAM files: identical records for many query windows,
  records that span BGZF blocks, long-CIGAR (CG tag) records, an empty window."""
import os
import zlib
import numpy as np
import pytest

from pepper_b200 import synth, synth_files

pytestmark = pytest.mark.gpu


def _deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    if not flush_every:
        return co.compress(data) + co.flush()
    out = b""
    for i in range(0, len(data), flush_every):
        out += co.compress(data[i:i + flush_every]) + co.flush(zlib.Z_FULL_FLUSH)
    return out + co.flush()


def test_inflate_matches_zlib_on_every_block_type():
    from pepper_b200.bamio import inflate_blocks
    rng = np.random.default_rng(5)
    text = bytes(rng.choice(np.frombuffer(b"ACGTACGTNNacgt\n\t0123456789", np.uint8), 60000))
    bamlike = bytes(rng.integers(0, 16, 30000).astype(np.uint8)) + bytes(rng.integers(1, 41, 30000).astype(np.uint8))
    skew = bytes(np.minimum(255, rng.geometric(0.02, 65000)).astype(np.uint8))           # long Huffman codes for rare symbols
    cases = [
        ("empty", b"", 6, zlib.Z_DEFAULT_STRATEGY, 0),
        ("stored", text[:50000], 0, zlib.Z_DEFAULT_STRATEGY, 0),
        ("fixed", text[:3000], 6, zlib.Z_FIXED, 0),
        ("dyn1", text, 1, zlib.Z_DEFAULT_STRATEGY, 0),
        ("dyn6", text, 6, zlib.Z_DEFAULT_STRATEGY, 0),
        ("dyn9", bamlike, 9, zlib.Z_DEFAULT_STRATEGY, 0),
        ("runs", b"A" * 40000 + b"AC" * 5000 + b"ACG" * 4000, 6, zlib.Z_DEFAULT_STRATEGY, 0),
        ("multi", text, 6, zlib.Z_DEFAULT_STRATEGY, 7000),
        ("skew", skew, 9, zlib.Z_DEFAULT_STRATEGY, 0),
        ("random", bytes(rng.integers(0, 256, 65536).astype(np.uint8)), 6, zlib.Z_DEFAULT_STRATEGY, 0),
        ("huffman_only", text[:20000], 6, zlib.Z_HUFFMAN_ONLY, 0),
        ("one_byte", b"x", 6, zlib.Z_DEFAULT_STRATEGY, 0),
    ]
    streams = [_deflate(d, lv, st, fl) for _, d, lv, st, fl in cases]
    outs, status = inflate_blocks(streams, [len(c[1]) for c in cases])
    for (name, d, *_), o, s in zip(cases, outs, status):
        assert s == 0, (name, int(s))
        assert o == d, name
    # 300 blocks at once (more warps than one wave of CTAs)
    many = [bytes(rng.integers(0, 8, int(rng.integers(1, 60000))).astype(np.uint8)) for _ in range(300)]
    outs, status = inflate_blocks([_deflate(m, 1 + (i % 9)) for i, m in enumerate(many)], [len(m) for m in many])
    assert not status.any() and all(o == m for o, m in zip(outs, many))


def test_inflate_rejects_corrupt_streams():
    from pepper_b200.bamio import inflate_blocks
    rng = np.random.default_rng(6)
    data = bytes(rng.integers(0, 20, 20000).astype(np.uint8))
    good = _deflate(data)
    bad = [good[:len(good) // 2],                               # truncated
           bytes([good[0] | 0x06]) + good[1:],                  # reserved block type
           good[:40] + bytes(rng.integers(0, 256, len(good) - 40).astype(np.uint8)),     # garbage after the header
           good]
    sizes = [len(data)] * 3 + [len(data) - 5]                   # the last one: wrong expected size
    outs, status = inflate_blocks(bad + [good], sizes + [len(data)])
    assert all(int(s) != 0 for s in status[:4]) and status[4] == 0 and outs[4] == data


@pytest.fixture(scope="module")
def bam_file(tmp_path_factory):
    d = tmp_path_factory.mktemp("bamdev")
    L = 120_000
    rec, genome = synth.simulate_contig_records(L, 25, synth.ONT, 61)
    rec2, _ = synth.simulate_contig_records(30_000, 10, synth.HIFI, 62)
    path = str(d / "t.bam")
    synth_files.write_bam(path, [("ctgA", L), ("ctgB", 30_000)], {0: rec, 1: rec2}, block_payload=20000, level=6, long_cigar_over=300)
    return path, L


def _same(a, b):
    for f in ("pos", "seq_off", "cigar_off", "flag", "mapq", "seq", "qual", "cigar"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("share", [0.0, 0.07, 0.33, 1.0])
def test_fetch_device_equals_host_fetch(bam_file, share):
    """`share` of the BGZF blocks (runs of 16) is inflated by the host pool beside the kernel (pb_bam_set_host_share): the records do
    not depend on it."""
    from pepper_b200.bamio import BamReader
    path, L = bam_file
    r = BamReader(path, 4)
    r.set_host_share(share)
    windows = [("ctgA", 0, L), ("ctgA", 40_000, 41_000), ("ctgA", 16_384, 32_768), ("ctgA", 99_000, 200_000), ("ctgA", 5, 6),
               ("ctgB", 0, 30_000), ("ctgB", 10_000, 20_000), ("ctgB", 29_990, 40_000)]
    for contig, a, b in windows:
        want = r.fetch(contig, a, b).to_batch()
        got = r.fetch_device(contig, a, b).to_batch()
        assert got.n_records == want.n_records, (contig, a, b)
        _same(got, want)
    assert r.fetch("ctgA", 0, L).n_records > 200
    t = r.fetch_device_timings()
    assert t["inflate_ms"] > 0
    on_host, on_device = r.inflate_split()
    assert (on_host == 0) == (share == 0.0) and on_host + on_device > 20
    if share == 1.0:
        assert on_host > on_device            # (empty blocks — the EOF marker — stay with the kernel)
    with pytest.raises(Exception):
        r.set_host_share(1.5)
    r.close()


def test_get_reads_from_device_fetch(bam_file):
    """file -> GPU inflate -> batched get_reads: the trimmed reads equal those of the host-inflate path."""
    from pepper_b200.bamio import BamReader
    from pepper_b200.reads import ReadTrimmer
    path, L = bam_file
    r = BamReader(path, 4)
    tr = ReadTrimmer(0)
    q = [(1000, 9000), (8000, 30000), (60000, 61200)]
    a = tr.get_reads(r.fetch("ctgA", 0, L), q, False, 0, 0).to_host()
    b = tr.get_reads(r.fetch_device("ctgA", 0, L), q, False, 0, 0).to_host()
    for f in ("pos", "seq_off", "cigar_off", "flags", "mapq", "seq", "qual", "cigar"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    tr.close(); r.close()
