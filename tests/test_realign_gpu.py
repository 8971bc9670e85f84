"""GPU parity for row f1: the realignment kernels (pepper_b200/csrc/realign.cu, through the C-ABI) against the plain-C
restatement of ReadAligner::align_reads_to_reference / SSW (oracle/port_realign.c), bit-exact positions and CIGARs."""
import numpy as np
import pytest

from pepper_b200 import synth

pytestmark = pytest.mark.gpu


def oracle_realign(oracle, reads, regions):
    pos, cig, off = [], [], [0]
    for r in range(regions.n_regions):
        row = regions.table[r]
        ref = regions.ref[int(row[4]):int(row[4] + row[5])].tobytes().decode()
        p, pe, co, c = oracle.realign(reads, int(row[6]), int(row[7]), int(row[0]), int(row[1]) + 20, ref, impl="port")
        assert p.shape[0] == int(row[7] - row[6])          # nothing dropped: reads start inside their region
        pos.append(p)
        cig.append(c)
        off.extend((co[1:] + off[-1]).tolist())
        off[-1] = off[-1]
    return np.concatenate(pos), np.array(off, dtype=np.int64), np.concatenate(cig)


def workload(n_regions, coverage, platform, seed):
    from pepper_b200.realign import realign_regions
    reads, regions = synth.make_polish_workload(n_regions, coverage, platform, seed=seed)
    genome = synth.make_reference(n_regions * 1000 + 1, seed)
    return reads, realign_regions(regions, genome)


@pytest.mark.parametrize("platform,seed", [(synth.ONT, 5), (synth.HIFI, 6)])
def test_realign_matches_oracle(oracle_built, platform, seed):
    from pepper_b200.realign import Realigner
    reads, regions = workload(4, 25, platform, seed)
    want_pos, want_off, want_cig = oracle_realign(oracle_built, reads, regions)
    ra = Realigner(0)
    got = ra.realign(reads, regions)
    st = ra.stats()
    assert st["realigned"] > 0.9 * reads.n_reads
    assert np.array_equal(got.pos, want_pos)
    assert np.array_equal(got.cigar_off, want_off)
    assert np.array_equal(got.cigar, want_cig)
    assert not np.array_equal(got.cigar_off, reads.cigar_off)          # the CIGARs really changed
    ra.close()


def test_realign_edge_reads(oracle_built):
    """Short reads (byte mode, score < 255), reads the aligner rejects (score <= 1), N bases, long indels, band growth."""
    from pepper_b200.realign import Realigner
    rng = np.random.default_rng(3)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 1400))
    recs = []

    def add(pos, seq):
        recs.append(dict(pos=pos, seq=seq, cigar=[(0, len(seq))]))
    add(0, ref[0:40])                                            # byte mode
    add(10, ref[10:72])                                          # 62 matches: 248 < 249 stays byte
    add(20, ref[20:83])                                          # 63 matches: 252 -> word
    add(30, "N" * 30)                                            # score 0 -> unchanged
    add(40, ref[40:300] + ref[340:700])                          # 40-base deletion -> band doubling
    add(50, ref[50:400] + "ACGTTGCA" * 6 + ref[400:800])         # 48-base insertion
    add(60, ref[60:500].replace("A", "N", 5))                    # N in the read
    add(70, "T" + ref[75:90])                                    # tiny
    add(80, ref[700:900])                                        # placed far from its true origin: soft clips / begin shift
    add(1300, ref[1300:1400] + "ACGTACGTACGTACGTACGTAAAA")       # runs past the reference end
    add(1399, "G")                                               # one base against a one-base reference
    recs.sort(key=lambda r: r["pos"])
    reads = synth.make_batch(recs)
    tab = np.array([[0, 1380, 0, 1380, 0, 1400, 0, reads.n_reads]], dtype=np.int64)
    regions = synth.RegionTable(tab, np.frombuffer(ref.encode(), dtype=np.uint8))
    want_pos, want_off, want_cig = oracle_realign(oracle_built, reads, regions)
    ra = Realigner(0)
    got = ra.realign(reads, regions)
    assert np.array_equal(got.pos, want_pos)
    assert np.array_equal(got.cigar_off, want_off)
    assert np.array_equal(got.cigar, want_cig)
    ra.close()


def test_realign_rejects_reads_before_region():
    from pepper_b200.realign import Realigner
    from pepper_b200._lib import PepperB200Error
    reads = synth.make_batch([dict(pos=5, seq="ACGTACGT", cigar=[(0, 8)])])
    tab = np.array([[10, 100, 10, 100, 0, 50, 0, 1]], dtype=np.int64)
    regions = synth.RegionTable(tab, np.frombuffer(b"ACGT" * 13, dtype=np.uint8)[:50].copy())
    ra = Realigner(0)
    with pytest.raises(PepperB200Error):
        ra.realign(reads, regions)
    ra.close()


def test_realign_long_reads_fallback_kernel(oracle_built):
    """Reads longer than 1344 bases take the int32 local-memory kernel."""
    from pepper_b200.realign import Realigner
    rng = np.random.default_rng(8)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 3300))
    q1 = ref[100:1500] + "ACGTAC" + ref[1500:2900]
    q2 = ref[0:700] + ref[720:2200]
    reads = synth.make_batch([dict(pos=0, seq=q2, cigar=[(0, len(q2))]), dict(pos=100, seq=q1, cigar=[(0, len(q1))]),
                              dict(pos=200, seq=ref[200:900], cigar=[(0, 700)])])
    tab = np.array([[0, 3280, 0, 3280, 0, 3300, 0, 3]], dtype=np.int64)
    regions = synth.RegionTable(tab, np.frombuffer(ref.encode(), dtype=np.uint8))
    want_pos, want_off, want_cig = oracle_realign(oracle_built, reads, regions)
    ra = Realigner(0)
    got = ra.realign(reads, regions)
    assert np.array_equal(got.pos, want_pos) and np.array_equal(got.cigar_off, want_off) and np.array_equal(got.cigar, want_cig)
    ra.close()


def test_realign_read_with_soft_clip_over_10kb(oracle_built):
    """a read that ends inside a ~1.2 kb polish region but carries a soft clip of > 10 kb (get_reads keeps trailing
    soft clips, bam_handler.cpp:240-266) used to abort the whole batch (32 x 320 = 10,240-base limit).  The global-slab kernel has
    no length limit, like the reference's SSW: the 13 kb read aligns, the others are unaffected."""
    from pepper_b200.realign import Realigner
    rng = np.random.default_rng(18)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 1300))
    junk = "".join("ACGT"[i] for i in rng.integers(0, 4, 12500))
    long_q = ref[300:1100] + junk                                 # 800 aligned bases + a 12.5 kb tail that matches nothing
    recs = [dict(pos=100, seq=ref[100:900], cigar=[(0, 800)]),
            dict(pos=300, seq=long_q, cigar=[(0, 800), (4, len(junk))]),
            dict(pos=400, seq=ref[400:1000].replace("C", "G", 3), cigar=[(0, 600)])]
    reads = synth.make_batch(recs)
    tab = np.array([[0, 1280, 0, 1280, 0, 1300, 0, 3]], dtype=np.int64)
    regions = synth.RegionTable(tab, np.frombuffer(ref.encode(), dtype=np.uint8))
    want_pos, want_off, want_cig = oracle_realign(oracle_built, reads, regions)
    ra = Realigner(0)
    got = ra.realign(reads, regions)
    assert np.array_equal(got.pos, want_pos) and np.array_equal(got.cigar_off, want_off) and np.array_equal(got.cigar, want_cig)
    a, b = int(got.cigar_off[1]), int(got.cigar_off[2])
    assert int(got.cigar[b - 1]) & 15 == 4 and int(got.cigar[b - 1]) >> 4 >= 12000      # the tail comes back as one soft clip
    ra.close()


def test_realign_properties_large_batch():
    """Properties at a size the oracle is not run at: sequences untouched, CIGAR read-consumption equals the read length, positions
    only move right, the realigned span stays inside the region reference, and the result is deterministic."""
    from pepper_b200.realign import Realigner
    reads, regions = workload(60, 40, synth.ONT, 17)
    ra = Realigner(0)
    a = ra.realign(reads, regions)
    st = ra.stats()
    b = ra.realign(reads, regions)
    assert np.array_equal(a.pos, b.pos) and np.array_equal(a.cigar, b.cigar) and np.array_equal(a.cigar_off, b.cigar_off)
    assert st["aligned"] == reads.n_reads and st["realigned"] > 0.95 * reads.n_reads
    ops, lens = a.cigar & 15, (a.cigar >> 4).astype(np.int64)
    assert np.isin(ops, [0, 1, 2, 4]).all()
    read_cons = np.where(np.isin(ops, [0, 1, 4]), lens, 0)
    ref_cons = np.where(np.isin(ops, [0, 2]), lens, 0)
    csum, rsum = np.concatenate([[0], np.cumsum(read_cons)]), np.concatenate([[0], np.cumsum(ref_cons)])
    assert np.array_equal(csum[a.cigar_off[1:]] - csum[a.cigar_off[:-1]], np.diff(reads.seq_off))
    assert (a.pos >= reads.pos).all()
    span = rsum[a.cigar_off[1:]] - rsum[a.cigar_off[:-1]]
    region_of = np.repeat(np.arange(regions.n_regions), regions.table[:, 7] - regions.table[:, 6])
    ref_end = regions.table[region_of, 0] + regions.table[region_of, 5]
    assert (a.pos + span <= ref_end).all()
    ra.close()


def test_realign_length_class_boundaries(oracle_built):
    """Read lengths at the register-class boundaries of the packed kernel (512/768/1024/1344 bases) and tiny reads."""
    from pepper_b200.realign import Realigner
    rng = np.random.default_rng(12)
    ref = "".join("ACGT"[i] for i in rng.integers(0, 4, 1500))
    recs = []
    for k, n in enumerate([1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513, 767, 768, 769, 1023, 1024, 1025,
                           1343, 1344, 1345]):
        start = k
        body = list(ref[start:start + n])
        for j in range(7, len(body), 53):                      # a few substitutions so that the CIGARs are not trivial
            body[j] = "ACGT"[("ACGT".index(body[j]) + 1) % 4]
        if n > 200:
            del body[100:103]                                  # and a deletion / an insertion
            body[150:150] = list("TTG")
        seq = "".join(body)[:n] if len(body) >= n else "".join(body) + ref[start + n:start + 2 * n - len(body)]
        recs.append(dict(pos=start, seq=seq[:n] if len(seq) >= n else seq, cigar=[(0, min(n, len(seq)))]))
    reads = synth.make_batch(recs)
    tab = np.array([[0, 1480, 0, 1480, 0, 1500, 0, reads.n_reads]], dtype=np.int64)
    regions = synth.RegionTable(tab, np.frombuffer(ref.encode(), dtype=np.uint8))
    want_pos, want_off, want_cig = oracle_realign(oracle_built, reads, regions)
    ra = Realigner(0)
    got = ra.realign(reads, regions)
    assert np.array_equal(got.pos, want_pos) and np.array_equal(got.cigar_off, want_off) and np.array_equal(got.cigar, want_cig)
    ra.close()
