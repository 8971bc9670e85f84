"""GPU: the file -> reads path (rows f4 + a2) through the reference-named handler classes: BAM_handler(path).get_reads
and FASTA_handler(path) of pepper_b200.build.PEPPER_VARIANT against the oracle's get_reads on the records that were
written to the file, and the whole chain file -> get_reads -> RegionalSummaryGenerator against the oracle chain."""
import numpy as np
import pytest

from pepper_b200 import synth, synth_files

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("files_gpu")
    start = 1000
    rec, genome = synth.simulate_contig_records(30000, 25, synth.ONT, 14, contig_start=start)
    full = np.concatenate([synth.make_reference(start, 99), genome])
    bam, fa = str(d / "s.bam"), str(d / "s.fa")
    synth_files.write_bam(bam, [("ctg", full.shape[0])], {0: rec})
    synth_files.write_fasta(fa, [("ctg", full)])
    return dict(bam=bam, fa=fa, rec=rec, genome=full)


def test_bam_handler_get_reads(oracle_built, files):
    from pepper_b200.build import PEPPER_VARIANT
    h = PEPPER_VARIANT.BAM_handler(files["bam"])
    assert h.get_chromosome_sequence_names() == ["ctg"]
    for (s, e, supp, mq, bq) in [(5000, 6201, False, 0, 0), (900, 1500, True, 0, 10), (20000, 26001, False, 5, 7), (31000, 40000, False, 0, 0)]:
        reads = h.get_reads("ctg", s, e, supp, mq, bq)
        want, pos_end, n_bad = oracle_built.get_reads(files["rec"], s, e, supp, mq, bq, impl="port")
        assert len(reads) == want.n_reads
        codes = want.codes()
        for i, r in enumerate(reads):
            so, se = int(want.seq_off[i]), int(want.seq_off[i + 1])
            assert r.pos == want.pos[i] and r.pos_end == pos_end[i]
            assert r.sequence == "".join(synth.NT16[c] for c in codes[so:se])
            assert r.base_qualities == want.qual[so:se].tolist()
            assert [(c.operation, c.length) for c in r.cigar_tuples] == \
                [(int(w & 15), int(w >> 4)) for w in want.cigar[want.cigar_off[i]:want.cigar_off[i + 1]]]
            assert r.flags.is_reverse == bool(want.flags[i] & 1) and r.mapping_quality == want.mapq[i]
            assert len(r.bad_indicies) == n_bad[i] and r.bad_indicies[-1] == len(r.sequence) + 1


def test_files_to_candidates_like_alignment_summarizer(oracle_built, files):
    """AlignmentSummarizer.py:181-236 with the handler classes swapped in."""
    from pepper_b200.build import PEPPER_VARIANT
    bam, fasta = PEPPER_VARIANT.BAM_handler(files["bam"]), PEPPER_VARIANT.FASTA_handler(files["fa"])
    p = synth.ont_params()
    s, e = 8000, 14000
    region_start, region_end = s - 100, e + 100
    all_reads = bam.get_reads("ctg", region_start, region_end, False, 0, int(p["min_snp_baseq"]))
    ref_seq = fasta.get_reference_sequence("ctg", region_start, region_end + 1)
    assert ref_seq == files["genome"][region_start:region_end + 1].tobytes().decode()
    gen = PEPPER_VARIANT.RegionalSummaryGenerator("ctg", region_start, region_end, ref_seq)
    gen.generate_max_insert_summary(all_reads)
    cands = gen.generate_summary(all_reads, p["min_snp_baseq"], p["min_indel_baseq"], p["snp_freq_threshold"], p["insert_freq_threshold"],
                                 p["delete_freq_threshold"], p["min_coverage_threshold"], p["snp_candidate_freq_threshold"],
                                 p["indel_candidate_freq_threshold"], p["candidate_support_threshold"], bool(p["skip_indels"]),
                                 s, e, 32, 26, False)
    want_reads, _, _ = oracle_built.get_reads(files["rec"], region_start, region_end, False, 0, int(p["min_snp_baseq"]), impl="port")
    tab = np.array([[region_start, region_end, s, e, 0, len(ref_seq), 0, want_reads.n_reads]], dtype=np.int64)
    o = oracle_built.variant_encode(want_reads, synth.RegionTable(tab, np.frombuffer(ref_seq.encode(), dtype=np.uint8)), p, "port")
    assert len(cands) == len(o["keys"]) > 5
    assert [c.position for c in cands] == o["positions"].tolist()
    assert [c.candidates[0] for c in cands] == o["keys"]


def test_polish_chain_file_to_consensus(oracle_built, files):
    """BAM file -> fetch -> get_reads -> realign -> polish encoder -> GRU -> bases, all on the device, against the same
    caller fed with the ORACLE's get_reads + realign output (and the oracle polish encoder on those reads)."""
    import torch
    from pepper_b200 import weights
    from pepper_b200.bamio import BamReader
    from pepper_b200.reads import ReadTrimmer
    from pepper_b200.realign import Realigner, realign_regions
    from pepper_b200.pipeline import PolishCaller, DeviceReads, FetchedReads
    from pepper_b200.polish import PolishEncoder
    genome = files["genome"]
    iv, rows = [], []
    for p in range(4000, 9000, 1000):
        rs, re_ = p - 100, p + 1100
        iv.append((rs, re_))
        rows.append([rs, re_, p, p + 1000, 0, 0, 0, 0])
    regions = realign_regions(synth.RegionTable(np.array(rows, dtype=np.int64), np.zeros(1, np.uint8)), genome)
    view = BamReader(files["bam"]).fetch("ctg", iv[0][0], iv[-1][1])
    tr, ra = ReadTrimmer(0), Realigner(0)
    got = tr.get_reads(view, iv, False, 0, 0)
    fetched = FetchedReads(got, regions)
    fetched.struct = ra.realign_device(fetched)
    # oracle chain
    batches, counts = [], []
    for (s, e) in iv:
        b, _, _ = oracle_built.get_reads(files["rec"], s, e, False, 0, 0, impl="port")
        batches.append(b)
        counts.append(b.n_reads)
    reads = synth.concat_batches(batches)
    tab = regions.table.copy()
    tab[:, 7] = np.cumsum(counts)
    tab[:, 6] = tab[:, 7] - counts
    oregions = synth.RegionTable(tab, regions.ref)
    pos, off, cig = [], [0], []
    for r in range(len(iv)):
        row = tab[r]
        ref = regions.ref[int(row[4]):int(row[4] + row[5])].tobytes().decode()
        p_, _, co, c = oracle_built.realign(reads, int(row[6]), int(row[7]), int(row[0]), int(row[1]) + 20, ref, impl="port")
        pos.append(p_)
        cig.append(c)
        off.extend((co[1:] + off[-1]).tolist())
    realigned = synth.ReadBatch(np.concatenate(pos), reads.seq_off, np.array(off, dtype=np.int64), reads.flags, reads.mapq, reads.seq,
                                reads.qual, np.concatenate(cig))
    # encoder parity on the realigned reads (host API vs oracle encoder)
    enc = PolishEncoder(0)
    s_gpu = enc.encode(realigned, oregions)
    s_or = oracle_built.polish_encode(realigned, oregions, "port")
    assert np.array_equal(s_gpu.image, s_or["image"]) and np.array_equal(s_gpu.pos, s_or["pos"])
    # whole chain on the device == caller fed with the oracle's reads
    pc = PolishCaller(weights.random_polish_state(0))
    dev = torch.device("cuda", 0)

    def outs(cap):
        return dict(bases=torch.empty((cap, 1000), dtype=torch.uint8, device=dev), phred=torch.empty((cap, 1000), dtype=torch.uint8, device=dev),
                    position=torch.empty((cap, 1000), dtype=torch.int64, device=dev), index=torch.empty((cap, 1000), dtype=torch.int32, device=dev),
                    image_region=torch.empty(cap, dtype=torch.int32, device=dev), chunk_id=torch.empty(cap, dtype=torch.int32, device=dev))
    a, b = outs(64), outs(64)
    na = pc.call_device(fetched, a)
    nb = pc.call_device(DeviceReads(realigned, oregions), b)
    assert na == nb and na >= len(iv)
    for k in a:
        assert torch.equal(a[k][:na], b[k][:nb]), k
    pc.close()
