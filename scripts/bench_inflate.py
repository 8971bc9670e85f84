#!/usr/bin/env python
"""Stand-alone timing of the GPU BAM fetch (pb_bam_fetch_device): BGZF inflate + record chains + parse + SoA scatter for one
batch of 100 kb regions of a synthetic ONT 30x BAM, nothing else running on the GPU.  Prints one JSON line."""
import argparse, json, os, sys, tempfile, time, shutil
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--level", type=int, default=1)
    a = ap.parse_args()
    from pepper_b200 import synth, synth_files
    from pepper_b200.bamio import BamReader
    span = 8 * 100_000
    rec, genome = synth.simulate_contig_records(span, 30.0, synth.ONT, 9)
    d = tempfile.mkdtemp(prefix="pb_inflate_")
    try:
        bam = os.path.join(d, "s.bam")
        times = max(1, -(-a.regions * 100_000 // span))
        L = synth_files.write_bam_tiled(bam, "c", rec, span, times, level=a.level)
        r = BamReader(bam, 0)
        end = min(L, a.regions * 100_000)
        best = None
        for _ in range(a.steps + 1):
            t0 = time.perf_counter()
            v = r.fetch_device("c", 0, end)
            dt = time.perf_counter() - t0
            t = r.fetch_device_timings()
            if best is None or dt < best[0]:
                best = (dt, t, v.n_records)
        comp, infl = r.io_stats()
        per = comp / (a.steps + 1), infl / (a.steps + 1)
        t0 = time.perf_counter(); hv = r.fetch("c", 0, end); host_s = time.perf_counter() - t0
        print(json.dumps({"regions": a.regions, "records": best[2], "compressed_MB": per[0] / 1e6, "inflated_MB": per[1] / 1e6, "wall_ms": best[0] * 1e3,
                          "device_ms": best[1], "inflate_out_GBps": per[1] / (best[1]["inflate_ms"] / 1e3) / 1e9,
                          "host_zlib_fetch_ms": host_s * 1e3, "host_records": hv.n_records, "bam_bytes": os.path.getsize(bam)}))
        r.close()
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
