#!/usr/bin/env python
"""Variant path from files on one B200: synthetic contig -> BAM + FASTA on disk -> pb_bam_fetch (host inflate) -> batched
get_reads on the GPU -> variant encoder -> LSTM.  Prints one JSON line with per-stage times (best of --steps)."""
import argparse, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=32)            # x 100 kb
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    import torch
    from pepper_b200 import synth, synth_files, weights
    from pepper_b200.frontend import VariantFromFiles, variant_intervals
    L = a.regions * 100_000 + 200
    t0 = time.time()
    rec, genome = synth.simulate_contig_records(L, a.coverage, synth.ONT, 7)
    d = tempfile.mkdtemp()
    bam, fa = os.path.join(d, "v.bam"), os.path.join(d, "v.fa")
    synth_files.write_bam(bam, [("ctg", L)], {0: rec})
    synth_files.write_fasta(fa, [("ctg", genome)])
    gen_s = time.time() - t0
    iv = variant_intervals(100, L - 100, 100_000)
    params = synth.ont_params()
    vf = VariantFromFiles(bam, fa, weights.random_variant_state(0), threads=a.threads)
    genomic = sum(e - s for s, e in iv)
    # stage times
    best = dict(fetch=1e9, get_reads=1e9, total=1e9)
    n_cand = 0
    for _ in range(a.steps + 1):
        t0 = time.perf_counter()
        view = vf.bam.fetch("ctg", 0, L)
        t1 = time.perf_counter()
        q = [(max(0, s - 100), e + 100) for s, e in iv]
        got = vf.trimmer.get_reads(view, q, False, 0, int(params["min_snp_baseq"]), max_reads=5000)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        best["fetch"] = min(best["fetch"], t1 - t0)
        best["get_reads"] = min(best["get_reads"], t2 - t1)
        t0 = time.perf_counter()
        calls, table = vf.call("ctg", iv, params, capacity=int(genomic // 40), want_images=False)
        torch.cuda.synchronize()
        best["total"] = min(best["total"], time.perf_counter() - t0)
        n_cand = len(calls)
    # streaming: batches of 32 regions, the next batch's BAM span is inflated while the GPU works on the current one
    stream_s = None
    if len(iv) >= 64:
        for _ in range(2):
            t0 = time.perf_counter()
            nb = sum(len(c) for c, _ in vf.call_batches("ctg", iv, params, batch=32, capacity=int(32 * 100_000 // 40), want_images=False))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            stream_s = dt if stream_s is None else min(stream_s, dt)
    t = vf.caller.timings()
    out = {"metric": "genomic bases/sec (make_images+inference) from BAM + FASTA files", "value": genomic / best["total"], "unit": "bases/s", "n_gpus": 1,
           "config": {"workload": "pepper_variant from files, synthetic ONT 30x", "regions": len(iv), "genomic_bases": genomic, "records": rec.n_records,
                      "aligned_bases": rec.n_bases, "bam_bytes": os.path.getsize(bam), "candidates": n_cand, "gen_seconds": round(gen_s, 1)},
           "ms": {"whole_call": best["total"] * 1e3, "bam_fetch": best["fetch"] * 1e3, "get_reads_incl_upload": best["get_reads"] * 1e3,
                  "encoder_kernels": t["encode_ms"], "network_kernels": t["network_ms"]},
           "bam_GBps": os.path.getsize(bam) / best["fetch"] / 1e9}
    if stream_s is not None:
        out["streaming"] = {"value": genomic / stream_s, "unit": "bases/s", "ms": stream_s * 1e3, "batch_regions": 32}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
