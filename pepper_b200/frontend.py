"""File -> predictions in one call: the per-process loops of the reference's make_images + inference steps
(`pepper_variant ImageGenerationUI.generate_image_and_save_to_file` :191-251 + `AlignmentSummarizer.create_summary`
:181-236 + `predict_distributed_gpu.predict` :58-70;  `pepper ImageGenerationUI.image_generator` :191-236 +
`AlignmentSummarizer.create_summary` :296-356 + `predict_distributed_gpu.predict` :63-105) with every stage after the BGZF
inflate on the GPU: `pb_bam_fetch` -> batched `get_reads` (+ reservoir sampling) -> [SSW realignment] -> pileup encoder
-> recurrent network.  Nothing here is a CLI: intervals in, arrays out (what the reference writes to its HDF5 stores).
"""
from __future__ import annotations

import numpy as np

from .bamio import BamReader, FastaReader
from ._lib import PepperB200Error
from .abi import PB_ERR_CAPACITY
from .pipeline import FetchedReads, PolishCaller, VariantCaller, PolishCalls, VariantCalls
from .reads import ReadTrimmer
from .realign import Realigner, ALIGNMENT_SAFE_BASES
from .synth import RegionTable

REGION_SAFE_BASES = 100              # ConsensCandidateFinder.REGION_SAFE_BASES (pepper_variant Options.py:2)
MIN_IMAGE_OVERLAP = 100              # ImageSizeOptions.MIN_IMAGE_OVERLAP (pepper Options.py:10)
VARIANT_MAX_READS = 5000             # AlingerOptions.MAX_READS_IN_REGION (pepper_variant Options.py:98)
# BAM fetches in flight in the streaming calls, and readers in rotation (FETCH_WORKERS being filled + the one whose records are
# being trimmed).  Measured on the chr20-scale files leg: 2 workers / 3 readers = 988 ms per step against 862 ms with 1 / 2 — the GPU
# is the bottleneck of that path, and two inflate kernels beside the network only slow the network down.
FETCH_WORKERS = 1
FETCH_READERS = 2
POLISH_MAX_READS = 1500              # AlingerOptions.MAX_READS_IN_REGION (pepper Options.py:28)


def polish_intervals(interval_start: int, interval_end: int, max_size: int = 1000) -> list[tuple[int, int]]:
    """pepper ImageGenerationUI.py:269-272."""
    return [(max(interval_start, pos - MIN_IMAGE_OVERLAP), min(interval_end, pos + max_size + MIN_IMAGE_OVERLAP))
            for pos in range(interval_start, interval_end, max_size)]


def variant_intervals(interval_start: int, interval_end: int, region_size: int = 100_000) -> list[tuple[int, int]]:
    """pepper_variant ImageGenerationUI.py:307-316."""
    return [(max(interval_start, pos), min(interval_end, pos + region_size)) for pos in range(interval_start, interval_end, region_size)]


class _FromFiles:
    """`gpu_inflate` (default): BGZF inflate, record walk and parse run on the GPU (pb_bam_fetch_device; only compressed blocks
    cross PCIe); False = the host thread-pool zlib path (pb_bam_fetch)."""

    def __init__(self, bam_path: str, fasta_path: str, device: int = 0, threads: int = 0, gpu_inflate: bool = True,
                 host_share: float | None = None):
        self.gpu_inflate = gpu_inflate
        self.host_share = host_share          # None = the reader's default (pb_bam_set_host_share)
        self.threads = threads
        self.bam = self._open_bam(bam_path)
        self.fasta = FastaReader(fasta_path)
        self.trimmer = ReadTrimmer(device)
        self.device = device
        self._extra = []                      # further readers of the same file (one per fetch in flight)

    def _open_bam(self, path: str) -> BamReader:
        r = BamReader(path, self.threads)
        if self.host_share is not None:
            r.set_host_share(self.host_share)
        return r

    def _readers(self, n: int) -> list:
        """`n` readers of the BAM (each owns the host / device buffers of one fetch in flight)."""
        while 1 + len(self._extra) < n:
            self._extra.append(self._open_bam(self.bam.path))
        return [self.bam] + self._extra[:n - 1]

    @property
    def _bam2(self):
        return self._extra[0] if self._extra else None

    def inflate_split(self) -> tuple[int, int]:
        """(BGZF blocks inflated by the host pools, by the kernel) over all readers."""
        parts = [r.inflate_split() for r in [self.bam] + self._extra if r is not None]
        return sum(a for a, _ in parts), sum(b for _, b in parts)

    def close(self):
        """Releases the file readers (the extra readers of the streaming calls included) and the trimmer."""
        for r in [getattr(self, "bam", None)] + list(getattr(self, "_extra", [])):
            if r is not None and hasattr(r, "close"):
                r.close()
        self.bam, self._extra = None, []
        if getattr(self, "trimmer", None) is not None:
            self.trimmer.close()
            self.trimmer = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _fetch(self, reader, contig: str, beg: int, end: int):
        return reader.fetch_device(contig, beg, end, self.device) if self.gpu_inflate else reader.fetch(contig, beg, end)

    @staticmethod
    def _group_span(g: list[tuple[int, int]]) -> tuple[int, int]:
        """The BAM span one batch of intervals needs (each interval widened by REGION_SAFE_BASES, AlignmentSummarizer.py:181-182)."""
        return max(0, min(s for s, _ in g) - REGION_SAFE_BASES), max(e for _, e in g) + REGION_SAFE_BASES

    def _side_stream(self):
        import torch
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)        # get_reads of batch k+1 runs here, beside the network of batch k
        return self._side

    def _trim_variant(self, view, contig: str, g: list[tuple[int, int]], min_snp_baseq: int, include_supplementary: bool = False,
                      min_mapq: int = 0, max_reads: int = VARIANT_MAX_READS, downsample_rate: float = 1.0, prof: dict | None = None):
        """Reference strings + batched get_reads of one batch of variant intervals on the side stream (synchronous for the host):
        fetched records -> the FetchedReads the encoder consumes."""
        import time
        t1 = time.perf_counter()
        rows, queries, spans = [], [], []
        for (a, b) in g:
            rs, re_ = max(0, a - REGION_SAFE_BASES), b + REGION_SAFE_BASES
            queries.append((rs, re_)); rows.append([rs, re_, a, b, 0, 0, 0, 0]); spans.append((rs, re_ + 1))
        regions = self._ref_table(contig, rows, spans)
        t2 = time.perf_counter()
        side = self._side_stream()
        got = self.trimmer.get_reads(view, queries, include_supplementary, min_mapq, min_snp_baseq,
                                     max_reads=max_reads, downsample_rate=downsample_rate, stream=side.cuda_stream)
        fetched = FetchedReads(got, regions, self.device, stream=side)
        if prof is not None:
            prof["ref_table"] += t2 - t1; prof["get_reads"] += time.perf_counter() - t2
        return fetched

    def _ref_table(self, contig: str, rows: list[list[int]], spans: list[tuple[int, int]]) -> RegionTable:
        """Region table + reference strings: ONE faidx fetch of the covering span, the (overlapping) per-region strings are
        offsets into it (get_reference_sequence clamps at the contig end; so do the lengths here)."""
        lo, hi = min(a for a, _ in spans), max(b for _, b in spans)
        ref = self.fasta.fetch_array(contig, lo, hi)
        clen = self.fasta.get_chromosome_sequence_length(contig)
        for row, (a, b) in zip(rows, spans):
            row[4], row[5] = a - lo, max(0, min(b, clen) - a)
        return RegionTable(np.array(rows, dtype=np.int64).reshape(-1, 8), ref if ref.shape[0] else np.zeros(1, np.uint8))


class VariantFromFiles(_FromFiles):
    """call_variant's make_images + run_inference for a list of intervals of one contig."""

    def __init__(self, bam_path: str, fasta_path: str, state: dict, device: int = 0, threads: int = 0, gpu_inflate: bool = True,
                 host_share: float | None = None):
        super().__init__(bam_path, fasta_path, device, threads, gpu_inflate, host_share)
        self.caller = VariantCaller(state, device)

    def call(self, contig: str, intervals: list[tuple[int, int]], params: dict, include_supplementary: bool = False,
             min_mapq: int = 0, downsample_rate: float = 1.0, max_reads: int = VARIANT_MAX_READS,
             capacity: int | None = None, want_images: bool = True, _view=None) -> tuple[VariantCalls, RegionTable]:
        import torch
        if not intervals:
            raise ValueError("no intervals")
        rows, queries, spans = [], [], []
        for (s, e) in intervals:
            rs, re_ = max(0, s - REGION_SAFE_BASES), e + REGION_SAFE_BASES          # AlignmentSummarizer.py:181-182
            queries.append((rs, re_))                                               # get_reads(chrom, region_start, region_end, ...)
            rows.append([rs, re_, s, e, 0, 0, 0, 0])
            spans.append((rs, re_ + 1))                                             # get_reference_sequence(.., region_end + 1)
        regions = self._ref_table(contig, rows, spans)
        view = _view if _view is not None else self._fetch(self.bam, contig, min(q[0] for q in queries), max(q[1] for q in queries))
        got = self.trimmer.get_reads(view, queries, include_supplementary, min_mapq, int(params["min_snp_baseq"]),
                                     max_reads=max_reads, downsample_rate=downsample_rate)
        fetched = FetchedReads(got, regions, self.device)
        dev = torch.device("cuda", self.device)
        cap = capacity or max(1024, int(sum(e - s + 1 for s, e in intervals)) // 16)
        while True:
            out = dict(images=torch.empty((cap, 33, 26), dtype=torch.int8, device=dev), positions=torch.empty(cap, dtype=torch.int64, device=dev),
                       depths=torch.empty(cap, dtype=torch.uint8, device=dev), freqs=torch.empty(cap, dtype=torch.uint8, device=dev),
                       keys=torch.empty((cap, 64), dtype=torch.uint8, device=dev), region_of=torch.empty(cap, dtype=torch.int32, device=dev),
                       probs=torch.empty((cap, 3), dtype=torch.float32, device=dev))
            try:
                n = self.caller.call_device(fetched, params, out)
                break
            except PepperB200Error as ex:                  # candidate capacity too small: retry with twice the room
                if ex.rc != PB_ERR_CAPACITY:
                    raise
                cap *= 2
        h = {k: v[:n].cpu().numpy() for k, v in out.items() if want_images or k != "images"}
        return VariantCalls(h["positions"], h["depths"], h["freqs"], h["keys"], h["region_of"], h["probs"], h.get("images")), fetched_table(fetched, regions)


    def find_candidate_tuples(self, contig: str, intervals: list[tuple[int, int]], params: dict, options: dict | None = None, **kw):
        """make_images + run_inference + the per-record candidate selection of find_candidates (CandidateFinder.py:356-530):
        returns (margin_records, deepvariant_records) in the tuple layouts the reference's VcfWriter consumes."""
        from .candidates import find_candidates, ONT_OPTIONS
        calls, table = self.call(contig, intervals, params, want_images=False, **kw)
        return find_candidates(contig, calls.positions, calls.region_of, calls.depths, calls.freqs, calls.keys_raw, calls.probs, table,
                               options or ONT_OPTIONS)

    def find_candidates(self, contig: str, intervals: list[tuple[int, int]], params: dict, options: dict | None = None,
                        vcf_options: dict | None = None, **kw) -> list[dict]:
        """The whole `pepper_variant find_candidates` step for one contig: make_images + run_inference, the per-record selection
        (CUDA kernel), the (contig, position) merge with (ref, alt) de-duplication (CandidateFinder.py:547-581) and the per-site VCF
        record assembly (VcfWriter.py:48-218).  Returns one dict per VCF record (pysam `new_record` keywords + `files`)."""
        from .candidates import ONT_OPTIONS
        from .vcf import find_site_records
        calls, table = self.call(contig, intervals, params, want_images=False, **kw)
        return find_site_records(contig, calls.positions, calls.region_of, calls.depths, calls.freqs, calls.keys_raw, calls.probs, table,
                                 options or ONT_OPTIONS, vcf_options)

    def call_batches(self, contig: str, intervals: list[tuple[int, int]], params: dict, batch: int = 32, **kw):
        """Streaming form: yields (VariantCalls, RegionTable) per batch of `batch` intervals while a helper thread inflates the next
        batch's BAM span with a second reader (pb_bam_fetch runs outside the GIL), so the host inflate overlaps the GPU work."""
        from concurrent.futures import ThreadPoolExecutor
        readers = self._readers(2)
        groups = [intervals[i:i + batch] for i in range(0, len(intervals), batch)]

        def span(g):
            return max(0, min(s for s, _ in g) - REGION_SAFE_BASES), max(e for _, e in g) + REGION_SAFE_BASES

        def prefetch(k):
            return self._fetch(readers[k & 1], contig, *span(groups[k]))
        # a Future re-raises the worker's exception (bad contig, corrupt BGZF block) in the consumer
        with ThreadPoolExecutor(max_workers=1) as pool:
            fut = pool.submit(prefetch, 0)
            for k, g in enumerate(groups):
                view = fut.result()
                if k + 1 < len(groups):
                    fut = pool.submit(prefetch, k + 1)
                yield self.call(contig, g, params, _view=view, **kw)


    def call_stream(self, contig: str, intervals: list[tuple[int, int]], params: dict, batch: int = 32, capacity: int | None = None,
                    include_supplementary: bool = False, min_mapq: int = 0, downsample_rate: float = 1.0,
                    max_reads: int = VARIANT_MAX_READS, want_images: bool = False) -> VariantCalls:
        """The whole interval list as ONE streaming session (pipeline.VariantStream): batch k+1's BGZF blocks are read, copied and
        inflated (helper thread, second reader, its own CUDA stream) while batch k's get_reads / encoder / network kernels run;
        candidates accumulate on the device and the network runs over whole 9,472-candidate chunks across batch boundaries.
        `region_of` counts intervals from the start of the list."""
        from concurrent.futures import ThreadPoolExecutor
        if not intervals:
            raise ValueError("no intervals")
        readers = self._readers(FETCH_READERS)
        groups = [intervals[i:i + batch] for i in range(0, len(intervals), batch)]

        def prefetch(k):
            return self._fetch(readers[k % FETCH_READERS], contig, *self._group_span(groups[k]))
        import time
        cap = capacity or max(4096, int(sum(e - s + 1 for s, e in intervals)) // 30)
        prof = dict(wait_fetch=0.0, ref_table=0.0, get_reads=0.0, stage=0.0, run=0.0, sync=0.0, end_fetch=0.0, fetch_thread=0.0)

        def timed_prefetch(k):
            t0 = time.perf_counter()
            v = prefetch(k)
            prof["fetch_thread"] += time.perf_counter() - t0
            return v
        def trim(view, g):
            return self._trim_variant(view, contig, g, int(params["min_snp_baseq"]), include_supplementary, min_mapq, max_reads,
                                      downsample_rate, prof)
        while True:
            s = self.caller.stream(params, cap)
            try:
                with ThreadPoolExecutor(max_workers=FETCH_WORKERS) as pool:
                    # FETCH_WORKERS fetches run at once (the pread of one batch beside the inflate of another), each on its own
                    # reader; batch j + FETCH_READERS is requested on the reader of batch j as soon as j's records are trimmed
                    futs = {k: pool.submit(timed_prefetch, k) for k in range(min(FETCH_READERS, len(groups)))}

                    def take(k):
                        t0 = time.perf_counter()
                        view = futs.pop(k).result()
                        prof["wait_fetch"] += time.perf_counter() - t0
                        fetched = trim(view, groups[k])
                        if k + FETCH_READERS < len(groups):
                            futs[k + FETCH_READERS] = pool.submit(timed_prefetch, k + FETCH_READERS)
                        return fetched
                    nxt = take(0)
                    done = 0
                    for k, g in enumerate(groups):
                        fetched = nxt
                        t3 = time.perf_counter()
                        s.stage_device(fetched, 0, len(g), done)
                        t4 = time.perf_counter()
                        s.run(flush=False)                 # encoder done (host-synchronous), network of this batch queued
                        t5 = time.perf_counter()
                        if k + 1 < len(groups):            # while that network runs: the next batch's records -> trimmed reads
                            nxt = take(k + 1)
                        # no sync: run(k+1) builds its tables while network(k) runs; its encoder synchronises the stream before the
                        # trimmer's buffers (consumed by encoder(k+1)) are overwritten by the trim of batch k+2
                        prof["stage"] += t4 - t3; prof["run"] += t5 - t4
                        done += len(g)
                t0 = time.perf_counter()
                n = s.end()
                out = s.fetch(n, want_images=want_images)
                prof["end_fetch"] = time.perf_counter() - t0
                self.last_profile = {k: round(v * 1e3, 2) for k, v in prof.items()}      # host wall time (ms) per stage of the last call
                return out
            except PepperB200Error as ex:
                if ex.rc != PB_ERR_CAPACITY:
                    raise
                cap *= 2


class VariantFileSource(_FromFiles):
    """The read source of dist.DistributedVariantCaller for a job given as FILES: every rank opens the same coordinate-sorted BAM
    (+ .bai) and FASTA and turns the interval groups it claims into trimmed device reads (GPU inflate -> record parse -> batched
    get_reads), so that make_images + run_inference of one contig shard over the ranks from the alignment file itself — no rank
    reads, inflates or copies a block outside the groups it runs.  `request(key, g0, g1)` starts the fetch of intervals[g0:g1] on a
    helper thread (FETCH_WORKERS fetches at once, FETCH_READERS readers in rotation, so the records of the group being trimmed stay
    valid while later spans are read and inflated — the caller keeps at most FETCH_READERS requests outstanding and takes them in
    request order); `take(key, g0, g1)` waits for it and returns the FetchedReads of that group."""

    def __init__(self, bam_path: str, fasta_path: str, contig: str, intervals: list[tuple[int, int]], min_snp_baseq: int,
                 device: int = 0, threads: int = 0, gpu_inflate: bool = True, include_supplementary: bool = False, min_mapq: int = 0,
                 downsample_rate: float = 1.0, max_reads: int = VARIANT_MAX_READS, host_share: float | None = None):
        from concurrent.futures import ThreadPoolExecutor
        super().__init__(bam_path, fasta_path, device, threads, gpu_inflate, host_share)
        self._ring = self._readers(FETCH_READERS)
        self.contig, self.intervals = contig, list(intervals)
        self._filters = (int(min_snp_baseq), include_supplementary, min_mapq, max_reads, downsample_rate)
        self._pool = ThreadPoolExecutor(max_workers=FETCH_WORKERS)
        self._futs, self._issued = {}, 0

    @property
    def n_regions(self) -> int:
        return len(self.intervals)

    def interval_work(self) -> np.ndarray:
        """Interval lengths: the work estimate the static schedule balances when the read counts are not known before the fetch."""
        return np.array([e - s + 1 for s, e in self.intervals], dtype=np.int64)

    def request(self, key, g0: int, g1: int) -> None:
        reader = self._ring[self._issued % FETCH_READERS]
        self._issued += 1
        self._futs[key] = self._pool.submit(self._fetch, reader, self.contig, *self._group_span(self.intervals[g0:g1]))

    def take(self, key, g0: int, g1: int) -> FetchedReads:
        view = self._futs.pop(key).result()            # a Future re-raises the worker's exception (bad contig, corrupt block) here
        return self._trim_variant(view, self.contig, self.intervals[g0:g1], *self._filters)

    def close(self):
        if getattr(self, "_pool", None) is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        super().close()


class PolishFromFiles(_FromFiles):
    """polish's make_images (with read realignment) + call_consensus for a list of regions of one contig."""

    def __init__(self, bam_path: str, fasta_path: str, state: dict, device: int = 0, threads: int = 0, gpu_inflate: bool = True,
                 host_share: float | None = None):
        super().__init__(bam_path, fasta_path, device, threads, gpu_inflate, host_share)
        self.caller = PolishCaller(state, device)
        self.realigner = Realigner(device)

    def call(self, contig: str, regions_se: list[tuple[int, int]], realign: bool = True, max_reads: int = POLISH_MAX_READS,
             capacity: int | None = None) -> tuple[PolishCalls, RegionTable]:
        import torch
        if not regions_se:
            raise ValueError("no regions")
        rows, spans = [], []
        for (rs, re_) in regions_se:
            rows.append([rs, re_, rs, re_, 0, 0, 0, 0])
            spans.append((rs, re_ + ALIGNMENT_SAFE_BASES))                         # AlignmentSummarizer.py:164-170
        regions = self._ref_table(contig, rows, spans)
        view = self._fetch(self.bam, contig, min(r[0] for r in regions_se), max(r[1] for r in regions_se))
        got = self.trimmer.get_reads(view, regions_se, False, 0, 0, max_reads=max_reads, downsample_rate=1.0)   # :300-325
        fetched = FetchedReads(got, regions, self.device)
        if realign:
            fetched.struct = self.realigner.realign_device(fetched)                # :328-332
        dev = torch.device("cuda", self.device)
        cap = capacity or 3 * (sum(e - s + 1 for s, e in regions_se) // 950 + len(regions_se)) + 8
        while True:
            out = dict(bases=torch.empty((cap, 1000), dtype=torch.uint8, device=dev), phred=torch.empty((cap, 1000), dtype=torch.uint8, device=dev),
                       position=torch.empty((cap, 1000), dtype=torch.int64, device=dev), index=torch.empty((cap, 1000), dtype=torch.int32, device=dev),
                       image_region=torch.empty(cap, dtype=torch.int32, device=dev), chunk_id=torch.empty(cap, dtype=torch.int32, device=dev))
            try:
                n = self.caller.call_device(fetched, out)
                break
            except PepperB200Error as ex:
                if ex.rc != PB_ERR_CAPACITY:
                    raise
                cap *= 2
        h = {k: v[:n].cpu().numpy() for k, v in out.items()}
        return PolishCalls(h["bases"], h["phred"], h["position"], h["index"], h["image_region"], h["chunk_id"]), fetched_table(fetched, regions)


def _polish_contig(self, contig: str, start: int = 0, end: int | None = None, batch: int = 2000, realign: bool = True,
                   return_calls: bool = False):
    """The whole polish path of one contig span: tiling (ImageGenerationUI.py:262-272: the span is [start, length - 1]),
    make_images + call_consensus in batches of `batch` regions, then the stitch (Stitch.py:36-128) -> polished sequence."""
    from .polish import stitch
    if end is None:
        end = self.fasta.get_chromosome_sequence_length(contig) - 1
    regs = polish_intervals(start, end)
    parts = []
    for i in range(0, len(regs), batch):
        calls, _ = self.call(contig, regs[i:i + batch], realign=realign)
        parts.append((calls, i))
    cat = lambda f: np.concatenate([getattr(c, f) for c, _ in parts])       # noqa: E731
    image_region = np.concatenate([c.image_region + off for c, off in parts]).astype(np.int32)
    allc = PolishCalls(cat("bases"), cat("phred"), cat("position"), cat("index"), image_region, cat("chunk_id"))
    seq = stitch(allc.bases, allc.position, allc.index, allc.image_region, allc.chunk_id, np.array([r[0] for r in regs], dtype=np.int64))
    return (seq, allc, regs) if return_calls else seq


PolishFromFiles.polish_contig = _polish_contig


def fetched_table(fetched: FetchedReads, regions: RegionTable) -> RegionTable:
    """The region table with the read ranges get_reads produced (after down-sampling)."""
    return RegionTable(fetched.table.copy(), regions.ref)


# ---------------------------------------------------------------------------------------------- prediction stores
def write_variant_predictions(store, contig: str, calls: VariantCalls, batch_size: int = 512, first_batch: int = 0) -> int:
    """What run_inference leaves on disk: predictions/batch_<n>/... in batches of `batch_size` candidates
    (pepper_variant predict_distributed_gpu.py:58-70 -> DataStorePredict.write_prediction).  Returns the next batch number."""
    keys = calls.keys
    b = first_batch
    for lo in range(0, len(calls), batch_size):
        hi = min(len(calls), lo + batch_size)
        store.write_prediction(b, [contig] * (hi - lo), calls.positions[lo:hi], calls.depths[lo:hi], keys[lo:hi], calls.freqs[lo:hi],
                               calls.probs[lo:hi])
        b += 1
    return b


def write_polish_predictions(store, contig: str, calls: PolishCalls, regions_se: list[tuple[int, int]]) -> None:
    """What call_consensus leaves on disk: one predictions/<contig>/<contig>-<start>-<end>/<chunk> group per image
    (pepper predict_distributed_gpu.py:107-109 -> DataStorePredict.write_prediction)."""
    for i in range(calls.bases.shape[0]):
        rs, re_ = regions_se[int(calls.image_region[i])]
        store.write_prediction(contig, rs, re_, int(calls.chunk_id[i]), calls.position[i], calls.index[i], calls.bases[i], calls.phred[i])

