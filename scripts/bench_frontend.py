#!/usr/bin/env python
"""Front-end rows (f4 file readers, a2 get_reads, f1 realignment) on one B200, polish tiling: synthetic contig -> BAM on disk ->
pb_bam_fetch -> batched get_reads -> realign -> polish encoder + GRU.  Prints one JSON line with per-stage times and the
CPU baselines (the unmodified reference functions compiled into oracle/_ref, one core, bounded sample)."""
import argparse, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contig", type=int, default=1_000_000)
    ap.add_argument("--coverage", type=float, default=40.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=6)
    a = ap.parse_args()
    import torch
    from pepper_b200 import synth, synth_files, weights
    from pepper_b200.bamio import BamReader
    from pepper_b200.reads import ReadTrimmer, DeviceRecords
    from pepper_b200.realign import Realigner, realign_regions
    from pepper_b200.pipeline import PolishCaller, FetchedReads
    t0 = time.time()
    rec, genome = synth.simulate_contig_records(a.contig, a.coverage, synth.ONT, 5)
    d = tempfile.mkdtemp()
    bam = os.path.join(d, "c.bam")
    synth_files.write_bam(bam, [("ctg", a.contig)], {0: rec})
    gen_s = time.time() - t0
    # polish tiling (pepper ImageGenerationUI.py:269-272): 1 kb chunks +-100
    iv, rows = [], []
    for p in range(0, a.contig, 1000):
        rs, re_ = max(0, p - 100), min(a.contig - 1, p + 1100)
        iv.append((rs, re_))
        rows.append([rs, re_, p, min(a.contig, p + 1000), 0, 0, 0, 0])
    regions = realign_regions(synth.RegionTable(np.array(rows, dtype=np.int64), np.zeros(1, np.uint8)), genome)
    out = {"workload": "polish front end, synthetic 40x ONT", "contig": a.contig, "records": rec.n_records, "bases": rec.n_bases,
           "bam_bytes": os.path.getsize(bam), "intervals": len(iv), "gen_seconds": round(gen_s, 1)}
    # f4: file -> records
    rd = BamReader(bam, threads=a.threads)
    best = 1e9
    for _ in range(a.steps):
        t0 = time.perf_counter()
        view = rd.fetch("ctg", 0, a.contig)
        best = min(best, time.perf_counter() - t0)
    comp, infl = rd.io_stats()
    out["bam_fetch_ms"] = best * 1e3
    out["bam_fetch_inflated_MBps"] = infl / a.steps / best / 1e6
    assert view.n_records == rec.n_records
    # a2: get_reads
    tr = ReadTrimmer(0)
    drec = DeviceRecords(rec)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(a.steps + 1):
        t0 = time.perf_counter()
        got = tr.get_reads(drec, iv, False, 0, 0)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    sizes = np.zeros(3, dtype=np.int64)
    tr.L.pb_get_reads_sizes(tr.h, sizes.ctypes.data)
    out["get_reads_ms"] = best * 1e3
    out["get_reads_out"] = {"reads": int(sizes[0]), "bases": int(sizes[1]), "cigar_ops": int(sizes[2])}
    out["get_reads_Gbases_per_s"] = float(sizes[1]) / best / 1e9
    # f1: realign
    fr = FetchedReads(got, regions)
    ra = Realigner(0)
    best = 1e9
    for _ in range(a.steps):
        t0 = time.perf_counter()
        new = ra.realign_device(fr)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    st = ra.stats()
    cells = 0.0
    rb, re_ = got.read_begin, got.read_end
    out["realign_ms"] = best * 1e3
    out["realign_stats"] = st
    out["realign_reads_per_s"] = float(sizes[0]) / best
    # cell updates: per read readLen x refLen (forward word pass) as the unit
    ref_len = regions.table[:, 5]
    mean_read = float(sizes[1]) / max(1, int(sizes[0]))
    out["realign_GCUPS_fwd_equiv"] = float(sizes[1]) * float(ref_len.mean()) * 0.5 / best / 1e9
    # downstream: encoder + GRU on the realigned reads
    fr.struct = new
    pc = PolishCaller(weights.random_polish_state(0))
    dev = torch.device("cuda", 0)
    cap = 3 * len(iv) + 16
    o = dict(bases=torch.empty((cap, 1000), dtype=torch.uint8, device=dev), phred=torch.empty((cap, 1000), dtype=torch.uint8, device=dev),
             position=torch.empty((cap, 1000), dtype=torch.int64, device=dev), index=torch.empty((cap, 1000), dtype=torch.int32, device=dev),
             image_region=torch.empty(cap, dtype=torch.int32, device=dev), chunk_id=torch.empty(cap, dtype=torch.int32, device=dev))
    best = 1e9
    for _ in range(a.steps):
        t0 = time.perf_counter()
        n_img = pc.call_device(fr, o)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    out["encode_infer_ms"] = best * 1e3
    out["images"] = n_img
    out["frontend_plus_polish_bases_per_s"] = a.contig / ((out["bam_fetch_ms"] + out["get_reads_ms"] + out["realign_ms"] + out["encode_infer_ms"]) / 1e3)
    # CPU baselines: the unmodified reference functions (oracle/_ref), one core, a few intervals
    try:
        from oracle import oracle
        if oracle.have_ref_getreads() and oracle.have_ref_realign():
            k = min(a.cpu_sample, len(iv))
            pick = np.linspace(0, len(iv) - 1, k).astype(int)
            t_get = t_re = 0.0
            n_reads = 0
            sub_rec = rec
            for i in pick:
                s, e = iv[i]
                t0 = time.perf_counter()
                b, _, _ = oracle.get_reads(sub_rec, s, e, False, 0, 0, impl="ref")
                t_get += time.perf_counter() - t0
                row = regions.table[i]
                ref = regions.ref[int(row[4]):int(row[4] + row[5])].tobytes().decode()
                t0 = time.perf_counter()
                oracle.realign(b, 0, b.n_reads, s, e + 20, ref, impl="ref")
                t_re += time.perf_counter() - t0
                n_reads += b.n_reads
            out["cpu_reference_1core"] = {"intervals": int(k), "get_reads_ms_per_interval": t_get / k * 1e3,
                                          "realign_ms_per_interval": t_re / k * 1e3, "realign_reads_per_s": n_reads / t_re,
                                          "note": "get_reads timing includes building the in-memory record table per call"}
    except Exception as ex:      # noqa: BLE001
        out["cpu_reference_1core"] = {"error": str(ex)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
