#!/usr/bin/env python
"""bench.py — genomic bases/sec through make_images + inference (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps K --warmup W            our CUDA path (libpepper_b200 through the C-ABI)
  python bench.py --impl reference ...                     the reference's CPU algorithm on the host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...    one rank per GPU, regions sharded by rank, one NCCL
                                                           all-gather of the per-candidate predictions

A "step" is one pass of the hot path (variant pileup-summary encoder -> bi-LSTM/MLP network) over one batch of
synthetic regions: configs[1] of BASELINE.json, "pepper_variant make_images + run_inference, synthetic ONT R9.4.1
30x", tiled in 100 kb intervals with 100 bp halos like pepper_variant ImageGenerationUI.py:307-316.

Printed JSON (one line, rank 0): `value` = whole-job genomic bases/s with the reads already resident in HBM;
`e2e` = the same metric through the public host-buffer API (pinned host reads -> H2D -> kernels -> D2H of the
prediction records); `roofline` for the dominant kernel (by time: the fused GEMM, tensor bound) and
`roofline_encoder` for the HBM-bound pileup kernel; `cpu_baseline` = the oracle (oracle/) timed on one host core
over a bounded sample.  Only the cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_CAND = 161.4e6          # 92.1 M dense + 69.2 M recurrent per candidate
TENSOR_FLOP_PER_CAND = 161.4e6   # in this design every one of them runs as a batched GEMM


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu: int):
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def encoder_algorithmic_bytes(reads, regions, n_cand: int) -> int:
    """per region sum over reads of (36 + 4 n_cigar + ceil(l_seq/2) + l_seq) + L_ref
    + N_cand * (33*26 + 4 + 1 + 1 + 62)."""
    lseq = np.diff(reads.seq_off)
    ncig = np.diff(reads.cigar_off)
    rd = int((36 + 4 * ncig + (lseq + 1) // 2 + lseq).sum())
    ref = int(regions.col("ref_len").sum())
    return rd + ref + n_cand * (33 * 26 + 4 + 1 + 1 + 62)


def build_workload(args, rank: int):
    from pepper_b200 import synth
    t0 = time.time()
    block_regions = min(args.block, args.regions)
    reads, regions = synth.make_variant_workload(block_regions, args.region_size, args.coverage, synth.ONT, seed=args.seed + rank)
    times = (args.regions + block_regions - 1) // block_regions
    reads, regions = synth.tile_workload(reads, regions, times)
    if regions.n_regions > args.regions:       # trim to the requested number of regions
        keep = args.regions
        nr = int(regions.table[keep - 1, 7])
        nb = int(reads.seq_off[nr]); nc = int(reads.cigar_off[nr])
        codes = reads.codes()[:nb]
        reads = synth.ReadBatch(reads.pos[:nr], reads.seq_off[:nr + 1], reads.cigar_off[:nr + 1], reads.flags[:nr], reads.mapq[:nr],
                                synth.pack_codes(codes), reads.qual[:nb], reads.cigar[:nc])
        rl = int(regions.table[keep - 1, 4] + regions.table[keep - 1, 5])
        regions = synth.RegionTable(regions.table[:keep].copy(), regions.ref[:rl])
    return reads, regions, time.time() - t0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pepper_b200 import synth, weights, _lib
    from pepper_b200.abi import HostReads
    from pepper_b200.pipeline import VariantCaller, DeviceReads
    from pepper_b200.dist import gather_predictions

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    _lib.require_gpu()
    peaks = load_peaks()
    params = synth.ont_params()

    reads, regions, gen_s = build_workload(args, rank)
    genomic_bases = regions.genomic_bases()
    caller = VariantCaller(weights.random_variant_state(0), device=local)
    dreads = DeviceReads(reads, regions, device=local)
    dev = torch.device("cuda", local)
    cap = max(4096, genomic_bases // 24)
    out = dict(images=torch.empty((cap, 33, 26), dtype=torch.int8, device=dev), positions=torch.empty(cap, dtype=torch.int64, device=dev),
               depths=torch.empty(cap, dtype=torch.uint8, device=dev), freqs=torch.empty(cap, dtype=torch.uint8, device=dev),
               keys=torch.empty((cap, 64), dtype=torch.uint8, device=dev), region_of=torch.empty(cap, dtype=torch.int32, device=dev),
               probs=torch.empty((cap, 3), dtype=torch.float32, device=dev))

    def gather(n_cand):
        """north_star: one NCCL all-gather of the per-region predictions (ragged counts, padded to the max)."""
        if world > 1:
            gather_predictions(out["probs"], n_cand, world)

    def step_device():
        n = caller.call_device(dreads, params, out)
        gather(n)
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg: W warm-up + K timed steps, CUDA events, max over ranks
    n_cand = 0
    for _ in range(args.warmup):
        n_cand = step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc_ms, net_ms, count_ms = [], [], []
    enc_parts = {}
    ev0.record()
    for _ in range(args.steps):
        n_cand = step_device()
        t = caller.timings()
        enc_ms.append(t["encode_ms"]); net_ms.append(t["network_ms"]); count_ms.append(t["enc_count"])
        for k, v in t.items():
            if k.startswith("enc_"):
                enc_parts.setdefault(k, []).append(v)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / args.steps
    value = world * genomic_bases / (ms_per_step / 1e3)
    enc_launches = 0
    import ctypes as C
    nl = C.c_int64(0)
    _lib.lib().pb_variant_encoder_launches.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    _lib.lib().pb_variant_encoder_launches(caller.enc.h, C.byref(nl))
    enc_launches = int(nl.value)
    gpu_launches = (enc_launches + caller.net.launches()) * args.steps

    # ---- end-to-end leg: pinned host reads -> H2D -> kernels -> D2H of the prediction records, every step
    hr = HostReads(reads, pin=True)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    calls = caller.call_prepared(hr, regions, params, capacity=n_cand + 16, reuse_buffers=True)     # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(e2e_steps):
        calls = caller.call_prepared(hr, regions, params, capacity=n_cand + 16, reuse_buffers=True)
        if world > 1:
            gather(len(calls))
    e1.record()
    barrier()
    e2e_t = caller.timings()                                   # device time of the encoder / network kernels inside the last call
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    h2d = hr.nbytes + regions.table.nbytes + regions.ref.nbytes
    d2h = len(calls) * (8 + 1 + 1 + 64 + 4 + 12)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    alg_bytes = encoder_algorithmic_bytes(reads, regions, n_cand)
    count_s = float(np.mean(count_ms)) / 1e3
    net_s = float(np.mean(net_ms)) / 1e3
    roof_enc = dict(bound="hbm", kernel="k_tile_count", achieved=alg_bytes / count_s / 1e9, peak=peaks["hbm_gbs"], unit="GB/s",
                    frac=alg_bytes / count_s / 1e9 / peaks["hbm_gbs"], traffic=int(alg_bytes * 1.30), peak_source=peaks["source"],
                    traffic_note="dram__bytes_read+write of k_tile_count = 1.30 x algorithmic bytes in the ncu --set full capture "
                                 "(77.2 MB vs 59.4 MB at 8 regions, profiles/README.md); scaled to this launch",
                    algorithmic_bytes_per_launch=alg_bytes, launch_ms=count_s * 1e3,
                    note="algorithmic bytes of the whole encoder (SURVEY 8d) over the pileup-count kernel's time")
    tf = n_cand * FLOP_PER_CAND / net_s / 1e12
    roof_net = dict(bound="tensor", kernel="k_lstm_layer + k_tc_gemm_p (tcgen05 LSTM layers / MLP GEMMs, all launches of the step; 3 bf16 products per algorithmic FLOP)", achieved=tf,
                    peak=peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"], unit="TFLOP/s",
                    frac=tf / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]),
                    traffic=int(n_cand * (237.36e6 / 3840 + 767.39e6 / 3840 + 678.66e6 / 9472)),
                    traffic_note="dram__bytes_read+write of the ncu --set full captures, per candidate: encoder LSTM layer (k_lstm_layer) 237.4 MB and "
                                 "decoder LSTM layer 767.4 MB per launch over 3,840 candidates, linear_1 678.7 MB per launch over 9,472 "
                                 "(profiles/r1_prof_lstm_*_raw.csv, r1_prof_tcp_lin1_final_raw.csv), scaled to the candidates of this step",
                    peak_source=peaks["source"] + ", sustained bf16",
                    flops_per_step=n_cand * FLOP_PER_CAND, step_ms=net_s * 1e3,
                    executed_bf16_tflops=3 * tf, executed_frac=3 * tf / (peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]),
                    note="achieved = fp32-equivalent algorithmic FLOPs (161.4 MFLOP per candidate) / network time; every FLOP is "
                         "executed as three bf16 tensor-core products (hi/lo split), so the tensor pipe runs at executed_frac; "
                         "ncu: decoder LSTM layer 92.8 % tensor-pipe active, encoder LSTM layer 59.1 %, linear_1 94.5 % (profiles/README.md); the step runs under the board power cap (clocks.reasons)")

    line = {
        "metric": "genomic bases/sec (make_images+inference)", "value": value, "unit": "bases/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 counts / int8 images (encoder); networks f32-equivalent (bf16 hi/lo split x3 on tcgen05, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": "pepper_variant make_images + run_inference, synthetic ONT R9.4.1 30x (BASELINE configs[1])",
                   "regions_per_gpu": regions.n_regions, "region_size": args.region_size, "coverage": args.coverage,
                   "genomic_bases_per_gpu": genomic_bases, "aligned_bases_per_gpu": reads.n_bases, "reads_per_gpu": reads.n_reads,
                   "candidates_per_gpu": n_cand, "parallelism": f"regions sharded over {world} GPU(s), 1 all-gather of predictions",
                   "l2": "inputs larger than L2 (%.0f MB of reads per step)" % (dreads.nbytes / 1e6),
                   "weights": "seeded random (no trained checkpoint offline)", "generated_block_regions": min(args.block, args.regions),
                   "gen_seconds": round(gen_s, 1)},
        "clocks": clocks,
        "e2e": {"value": world * genomic_bases / (e2e_ms / 1e3), "unit": "bases/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": e2e_steps,
                "kernel_ms": {"encoder": e2e_t["encode_ms"], "network": e2e_t["network_ms"]},
                "api": "pepper_b200.pipeline.VariantCaller.call -> pb_variant_call_host (pinned host buffers)"},
        "gpu_launches": int(gpu_launches),
        "phase_ms": {"encoder": float(np.mean(enc_ms)), "network": float(np.mean(net_ms)), "encoder_count_kernel": float(np.mean(count_ms)),
                     "encoder_phases": {k: float(np.mean(v)) for k, v in enc_parts.items()}},
        "roofline": roof_net,
        "roofline_encoder": roof_enc,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, threads=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/): the reference's algorithm on the host cores, organised the way
# the reference organises its CPU run — P single-threaded worker processes over regions for make_images
# (pepper_variant ImageGenerationUI.py:326) and over candidate slices for inference with 1 intra-op thread per
# caller (pepper_variant predict_distributed_cpu.py:47-57).
# ---------------------------------------------------------------------------------------------------------------
_W = {}


def _worker_init():
    import torch
    torch.set_num_threads(1)
    from oracle import oracle, nets
    oracle.lib("port")
    _W["impl"] = "ref" if oracle.have_ref() else "port"
    if _W["impl"] == "ref":
        oracle.lib("ref_variant")
    _W["state"] = nets.make_variant_weights(0)
    _W["oracle"], _W["nets"] = oracle, nets


def _worker_encode(task):
    from pepper_b200 import synth
    sub, tab = task
    c = _W["oracle"].variant_encode(sub, tab, synth.ont_params(), _W["impl"])
    return _W["oracle"].images_to_int8(c["images"])


def _worker_net(images):
    return _W["nets"].variant_predict(_W["state"], images, batch=512, threads=1)


def _worker_ready(_):
    return _W["impl"]


class CpuReference:
    def __init__(self, args, procs: int):
        from pepper_b200 import synth
        self.procs = procs
        self.nreg = 1 if procs == 1 else max(2, min(32, procs // 8))
        reads, regions = synth.make_variant_workload(self.nreg, args.region_size, args.coverage, synth.ONT, seed=args.seed)
        self.tasks = [synth.region_batch(reads, regions, r) for r in range(self.nreg)]
        self.genomic_bases = regions.genomic_bases()
        if procs == 1:
            _worker_init()
            self.pool = None
            self.impl = _W["impl"]
        else:
            import multiprocessing as mp
            self.pool = mp.get_context("spawn").Pool(procs, initializer=_worker_init)
            self.impl = self.pool.map(_worker_ready, range(procs))[0]

    def step(self):
        t0 = time.perf_counter()
        if self.pool is None:
            parts = [_worker_encode(t) for t in self.tasks]
        else:
            parts = self.pool.map(_worker_encode, self.tasks, chunksize=1)
        images = np.concatenate(parts)
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        if self.pool is None:
            _worker_net(images)
        else:
            per = max(64, -(-images.shape[0] // self.procs))
            self.pool.map(_worker_net, [images[i:i + per] for i in range(0, images.shape[0], per)], chunksize=1)
        t_net = time.perf_counter() - t0
        return t_enc, t_net, images.shape[0]

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()

    def describe(self, t_enc, t_net, n_cand):
        enc = "reference C++ compiled into oracle/_ref" if self.impl == "ref" else "oracle/port_encoders.c"
        return (f"{self.nreg} region(s) x 100 kb of the same workload ({n_cand} candidates) per step over {self.procs} single-threaded "
                f"worker process(es): encoder = {enc} {t_enc:.2f}s, network = oracle/nets.py (PyTorch CPU, bit-identical to the "
                f"reference nn.Module) {t_net:.2f}s")


def cpu_baseline(args, threads: int):
    ref = CpuReference(args, threads)
    t_enc, t_net, n_cand = ref.step()
    ref.close()
    return {"value": ref.genomic_bases / (t_enc + t_net), "unit": "bases/s", "cores": threads,
            "kind": "reference" if ref.impl == "ref" else "port", "sample": ref.describe(t_enc, t_net, n_cand),
            "encoder_s": t_enc, "network_s": t_net}


def run_reference(args):
    """The reference's CPU implementation of the path on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    ref = CpuReference(args, procs)
    tot_t, last = 0.0, None
    for i in range(args.warmup + args.steps):
        t_enc, t_net, n_cand = ref.step()
        if i >= args.warmup:
            tot_t += t_enc + t_net
        last = (t_enc, t_net, n_cand)
    ref.close()
    value = ref.genomic_bases * args.steps / tot_t
    line = {"impl": "reference", "metric": "genomic bases/sec (make_images+inference)", "value": value, "unit": "bases/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/f64 (encoder), f32 (network)", "data": "synthetic",
            "config": {"workload": "pepper_variant make_images + run_inference, synthetic ONT R9.4.1 30x (BASELINE configs[1])",
                       "region_size": args.region_size, "coverage": args.coverage, "sample_regions_per_step": ref.nreg},
            "cpu_baseline": {"value": value, "unit": "bases/s", "cores": procs, "kind": "reference" if ref.impl == "ref" else "port",
                             "sample": ref.describe(*last)},
            "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--regions", type=int, default=int(os.environ.get("PB_BENCH_REGIONS", "645")),
                    help="100 kb regions per GPU per step (chr20 = 645)")
    ap.add_argument("--region-size", type=int, default=100000)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--block", type=int, default=8, help="regions generated from scratch; tiled up to --regions")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
