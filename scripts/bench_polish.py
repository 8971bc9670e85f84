#!/usr/bin/env python
"""Secondary workload (BASELINE configs[2]): pepper polish make_images + call_consensus, synthetic draft + 40x ONT reads,
1 x B200.  Prints one JSON line in the same shape as bench.py (device-resident value + e2e through PolishCaller.call)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=5000)       # 5 Mb draft in 1 kb regions
    ap.add_argument("--block", type=int, default=250)
    ap.add_argument("--coverage", type=float, default=40.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    import torch
    from pepper_b200 import synth, weights
    from pepper_b200.pipeline import PolishCaller, DeviceReads
    t0 = time.time()
    reads, regions = synth.make_polish_workload(min(a.block, a.regions), a.coverage, synth.ONT, seed=3)
    reads, regions = synth.tile_workload(reads, regions, -(-a.regions // min(a.block, a.regions)))
    gen_s = time.time() - t0
    genomic = int((regions.col("cand_end") - regions.col("cand_start")).sum())
    pc = PolishCaller(weights.random_polish_state(0))
    d = DeviceReads(reads, regions)
    dev = torch.device("cuda", 0)
    cap = 3 * regions.n_regions + 16
    out = dict(bases=torch.empty((cap, 1000), dtype=torch.uint8, device=dev), phred=torch.empty((cap, 1000), dtype=torch.uint8, device=dev),
               position=torch.empty((cap, 1000), dtype=torch.int64, device=dev), index=torch.empty((cap, 1000), dtype=torch.int32, device=dev),
               image_region=torch.empty(cap, dtype=torch.int32, device=dev), chunk_id=torch.empty(cap, dtype=torch.int32, device=dev))
    n_img = 0
    for _ in range(a.warmup):
        n_img = pc.call_device(d, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    enc, net = [], []
    for _ in range(a.steps):
        n_img = pc.call_device(d, out)
        t = pc.timings(); enc.append(t["encode_ms"]); net.append(t["network_ms"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # end to end through the host-buffer API: page-locked reads -> H2D -> kernels -> D2H of bases / phred / positions
    from pepper_b200.abi import HostReads
    hr = HostReads(reads, pin=True)
    calls = pc.call_prepared(hr, regions, reuse_buffers=True)            # warm-up (allocates the pinned result buffers)
    t0 = time.perf_counter()
    for _ in range(2):
        calls = pc.call_prepared(hr, regions, reuse_buffers=True)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / 2
    print(json.dumps({"metric": "genomic bases/sec (make_images+call_consensus)", "value": genomic / (ms / 1e3), "unit": "bases/s", "n_gpus": 1,
                      "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
                      "config": {"workload": "pepper polish, synthetic draft + 40x ONT (BASELINE configs[2])", "regions": regions.n_regions,
                                 "genomic_bases": genomic, "aligned_bases": reads.n_bases, "images": n_img, "gen_seconds": round(gen_s, 1)},
                      "phase_ms": {"encoder": float(np.mean(enc)), "network": float(np.mean(net))},
                      "network_tflops": n_img * 1.53e9 / (float(np.mean(net)) / 1e3) / 1e12,
                      "e2e": {"value": genomic / (e2e_ms / 1e3), "unit": "bases/s", "ms_per_step": e2e_ms, "images": int(calls.bases.shape[0])},
                      "launches": pc.net.launches()}))


if __name__ == "__main__":
    main()
