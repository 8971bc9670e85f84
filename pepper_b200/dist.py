"""Multi-GPU product path of the variant hot path (replaces the reference's file-chunk round robin over
processes / DataParallel, pepper_variant RunInference.py:70-72,104-106, ImageGenerationUI.py:307-316,
predict_distributed_gpu.py:41).

Regions (100 kb intervals) are independent units.  A job is cut into region GROUPS; every rank runs a streaming session
(`pipeline.VariantStream`) over the groups it owns and the network's head kernel writes each candidate's 84-byte prediction
record straight into this rank's slice of ONE gather buffer; a single fixed-capacity all-gather (no host round trip for the
ragged counts, which travel in a second tiny all-gather) then gives every rank the whole job's records.

Two schedules:
  * static  — contiguous blocks of groups balanced by aligned-base count (`shard_regions`): rank-major order == genomic order;
  * dynamic — ranks claim the next group from an atomic counter in the rendezvous store (`GroupClaimer`), so a GPU that runs
              slower under the board power cap simply takes fewer groups; genomic order is restored from the region ids the
              records carry (`order_records`).
Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests)."""
from __future__ import annotations

import numpy as np

from .abi import PRED_RECORD

RECORD_BYTES = PRED_RECORD.itemsize


def shard_regions(work: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous blocks [begin, end) of units per rank, balanced by `work` (e.g. aligned bases per region / group).
    Contiguity keeps rank-major order == genomic order after the gather."""
    n = int(work.shape[0])
    if n == 0:
        return [(0, 0)] * world
    csum = np.concatenate([[0], np.cumsum(work.astype(np.float64))])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        # choose the nearer of k-1 / k to the target
        if k > cuts[-1] and abs(csum[k - 1] - target) <= abs(csum[min(k, n)] - target):
            k -= 1
        cuts.append(max(k, cuts[-1]))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def plan_groups(n_regions: int, group: int = 32) -> list[tuple[int, int]]:
    """Region groups [g0, g1) of a job: the hand-out unit (about 50 k candidates = 5 network chunks at ONT 30x)."""
    return [(g, min(n_regions, g + group)) for g in range(0, n_regions, group)]


def plan_groups_tapered(n_regions: int, group: int, world: int, min_group: int = 4) -> list[tuple[int, int]]:
    """Guided self-scheduling for the dynamic schedule: full `group`-sized groups, then shrinking ones (remaining / (2 world), at
    least `min_group`) so that the tail imbalance between ranks is a few regions, not one full group."""
    out, g = [], 0
    while g < n_regions:
        rem = n_regions - g
        size = max(min_group, min(group, -(-rem // (2 * max(1, world)))))
        out.append((g, min(n_regions, g + size)))
        g += size
    return out


def group_work(seq_off: np.ndarray, table: np.ndarray, groups: list[tuple[int, int]]) -> np.ndarray:
    """Aligned bases per group (the balance criterion of SURVEY 8e) from the reads' sequence offsets."""
    return np.array([int(seq_off[int(table[g1 - 1, 7])] - seq_off[int(table[g0, 6])]) for g0, g1 in groups], dtype=np.int64)


class GroupClaimer:
    """Hands out group indices.  `dynamic`: an atomic counter in the process group's rendezvous store (TCPStore.add is atomic;
    ~20 us per claim against ~20 ms of GPU work per group).  `static`: this rank's contiguous block."""

    def __init__(self, n_groups: int, rank: int, world: int, schedule: str = "dynamic", work: np.ndarray | None = None,
                 store=None, key: str = "pb_claim"):
        self.n, self.rank, self.world, self.schedule = n_groups, rank, world, schedule
        if schedule == "static" or world == 1:
            b, e = shard_regions(work if work is not None else np.ones(n_groups), world)[rank]
            self._it = iter(range(b, e))
            self.store = None
        elif schedule == "dynamic":
            if store is None:
                import torch.distributed as dist
                store = dist.distributed_c10d._get_default_store()
            self.store, self.key = store, key
            self._it = None
        else:
            raise ValueError("schedule must be 'static' or 'dynamic'")

    def next(self):
        if self._it is not None:
            return next(self._it, None)
        g = int(self.store.add(self.key, 1)) - 1
        return g if g < self.n else None

    def exhausted(self) -> bool:
        """Dynamic schedule: has every group been claimed?  (a rank that stopped claiming for lack of room checks this)"""
        if self._it is not None:
            return True
        return int(self.store.add(self.key, 0)) >= self.n


class GatherBuffer:
    """[world][capacity] prediction records + [world] counts.  On NCCL the record tensor is allocated from NCCL's own
    allocator (ncclMemAlloc through torch.cuda.MemPool) when this torch build offers it, so the collective works on registered
    user buffers; otherwise a plain tensor."""

    def __init__(self, capacity: int, world: int, rank: int, device=None, max_groups: int = 0):
        import torch
        self.capacity, self.world, self.rank = int(capacity), world, rank
        self.max_groups = int(max_groups)
        self.registered = False
        nbytes = world * self.capacity * RECORD_BYTES
        dev = device if device is not None else torch.device("cpu")
        self._dev = dev
        self.records = None
        if getattr(dev, "type", "cpu") == "cuda" and world > 1:
            try:
                import torch.distributed as dist
                backend = dist.distributed_c10d._get_default_group()._get_backend(dev)
                pool = torch.cuda.MemPool(backend.mem_allocator)
                with torch.cuda.use_mem_pool(pool):
                    self.records = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                backend.register_mem_pool(pool)
                self._pool = pool
                self.registered = True
            except Exception:
                self.records = None
        if self.records is None:
            self.records = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        # per rank: [n_valid, (group id, candidates) x max_groups] — the ragged counts and, for the dynamic schedule, the order
        # in which the rank ran its groups (group id -1 = unused slot)
        self._w = 1 + 2 * self.max_groups
        self.meta = torch.zeros((world, self._w), dtype=torch.int64, device=dev)
        self._mine = torch.zeros(self._w, dtype=torch.int64, device=dev)
        self._mine_host = torch.zeros(self._w, dtype=torch.int64).pin_memory() if getattr(dev, "type", "cpu") == "cuda" else torch.zeros(self._w, dtype=torch.int64)

    @property
    def counts(self):
        return self.meta[:, 0]

    @property
    def slice_bytes(self) -> int:
        return self.capacity * RECORD_BYTES

    def my_slice(self):
        return self.records[self.rank * self.slice_bytes:(self.rank + 1) * self.slice_bytes]

    def my_ptr(self) -> int:
        return self.records.data_ptr() + self.rank * self.slice_bytes

    def set_mine(self, n_valid: int, segments=None):
        """This rank's row of the meta table: candidate count and the (group id, candidates) list in processing order."""
        if n_valid > self.capacity:
            raise ValueError("n_valid %d > capacity %d" % (n_valid, self.capacity))
        h = self._mine_host
        h.zero_()
        h[0] = n_valid
        if self.max_groups:
            h[1::2] = -1
            for i, (g, c) in enumerate(segments or []):
                h[1 + 2 * i], h[2 + 2 * i] = g, c
        self._mine.copy_(h, non_blocking=True)

    def gather_meta(self, group=None):
        import torch.distributed as dist
        if self.world == 1:
            self.meta[0].copy_(self._mine)
        else:
            dist.all_gather_into_tensor(self.meta.view(-1), self._mine, group=group)

    def gather_records(self, group=None):
        """In place: this rank's input is its own slice of the output."""
        import torch.distributed as dist
        if self.world > 1:
            dist.all_gather_into_tensor(self.records, self.my_slice(), group=group)

    def gather(self, n_valid: int, segments=None, group=None):
        """Both collectives are enqueued back to back; nothing is read on the host here."""
        self.set_mine(n_valid, segments)
        self.gather_meta(group)
        self.gather_records(group)

    def to_host(self, order: bool = True) -> np.ndarray:
        """The valid records of every rank on the host (one small host sync for the meta table, then asynchronous D2H copies
        straight into their FINAL place in one page-locked buffer — rank-major, or group by group in genomic order when the
        ranks reported their group segments — so no host-side re-ordering pass is needed).  The returned array is a view of
        that buffer (valid until the next to_host)."""
        import torch
        meta = self.meta.cpu().numpy()
        counts = [int(meta[r, 0]) for r in range(self.world)]
        total = sum(counts)
        segs = []                                              # (group id, source record index in the gather buffer, count)
        for r in range(self.world):
            off = r * self.capacity
            for i in range(self.max_groups):
                g, k = int(meta[r, 1 + 2 * i]), int(meta[r, 2 + 2 * i])
                if g < 0:
                    break
                segs.append((g, off, k))
                off += k
        by_segment = order and bool(segs) and sum(k for _, _, k in segs) == total
        if by_segment:
            segs.sort()
            # merge neighbours that are contiguous in the source (a static schedule collapses to one copy per rank)
            plan = []
            for g, o, k in segs:
                if plan and plan[-1][0] + plan[-1][1] == o:
                    plan[-1][1] += k
                else:
                    plan.append([o, k])
        else:
            plan = [[r * self.capacity, counts[r]] for r in range(self.world)]
        cuda = self.records.is_cuda
        need = max(total, 1) * RECORD_BYTES
        if getattr(self, "_host", None) is None or self._host.numel() < need:
            self._host = torch.empty(need + need // 8, dtype=torch.uint8)
            if cuda:
                self._host = self._host.pin_memory()
        dst = 0
        for o, k in plan:
            if k:
                self._host[dst * RECORD_BYTES:(dst + k) * RECORD_BYTES].copy_(self.records[o * RECORD_BYTES:(o + k) * RECORD_BYTES], non_blocking=True)
                dst += k
        if cuda:
            torch.cuda.current_stream(self.records.device).synchronize()
        rec = self._host.numpy()[:total * RECORD_BYTES].view(PRED_RECORD)
        return order_records(rec) if (order and not by_segment) else rec

    def release(self):
        """Drops the NCCL-registered pool BEFORE the process group goes away (a registered MemPool destroyed after
        destroy_process_group aborts the process)."""
        self.records = None
        self.meta = None
        pool = getattr(self, "_pool", None)
        if pool is not None:
            try:
                import torch.distributed as dist
                backend = dist.distributed_c10d._get_default_group()._get_backend(self._dev)
                backend.deregister_mem_pool(pool)
            except Exception:
                pass
            self._pool = None
            del pool


def order_records(rec: np.ndarray) -> np.ndarray:
    """Genomic order = region id order (within a region the encoder's order is kept: stable sort)."""
    if rec.shape[0] == 0 or np.all(np.diff(rec["region"]) >= 0):
        return rec
    return rec[np.argsort(rec["region"], kind="stable")]


def records_from_calls(calls) -> np.ndarray:
    """A single-GPU VariantCalls as prediction records (the 1-rank answer the N-rank gather must reproduce)."""
    rec = np.zeros(len(calls), PRED_RECORD)
    rec["probs"] = calls.probs
    rec["position"] = calls.positions.astype(np.int32)
    rec["region"] = calls.region_of
    rec["depth"] = calls.depths
    rec["freq"] = calls.freqs
    rec["key"] = np.ascontiguousarray(calls.keys_raw[:, :62]).view("S62")[:, 0]
    return rec


class DistributedVariantCaller:
    """make_images + run_inference of ONE job (a region list every rank can read) over all ranks of the default process group.

        dvc = DistributedVariantCaller(state, local_device, capacity)
        dvc.run(source, regions, params)        # source: abi.HostReads (pinned host buffers), pipeline.DeviceReads, or
                                                # frontend.VariantFileSource (BAM + FASTA every rank opens; regions=None)
        records = dvc.buffer.to_host()          # every rank holds the whole job's records after the gather
    """

    def __init__(self, state: dict, device: int, capacity: int, schedule: str = "dynamic", group_regions: int = 32):
        import torch
        import torch.distributed as dist
        from .pipeline import VariantCaller
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.caller = VariantCaller(state, device)
        self.device = torch.device("cuda", device)
        self.schedule, self.group_regions = schedule, group_regions
        self.buffer = None
        self._capacity = capacity
        self._calls = 0
        self._ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        self.phase_ms = {}
        self.groups_done = 0

    def close(self):
        """Call before destroy_process_group()."""
        if self.buffer is not None:
            self.buffer.release()
            self.buffer = None
        self.caller.close()

    def run(self, source, regions, params: dict, seq_off: np.ndarray | None = None, replicas: int = 1) -> int:
        """Returns this rank's candidate count; the gathered records are in self.buffer (device).
        `replicas` > 1 (benchmarks only): the job is `replicas` copies of the region list laid end to end — job group J is local
        group J mod G with region ids shifted by (J div G) * n_regions — so that a weak-scaling run can hand out N x the
        per-GPU block without holding N copies of it."""
        import torch
        from collections import deque
        from .abi import HostReads, regions_array
        from .frontend import VariantFileSource
        from .pipeline import DeviceReads
        from_files = isinstance(source, VariantFileSource)
        n_reg = source.n_regions if from_files else regions.n_regions
        plain = plan_groups(n_reg, self.group_regions)
        dynamic = self.schedule == "dynamic" and self.world > 1
        last = plan_groups_tapered(n_reg, self.group_regions, self.world) if dynamic else plain
        # the job: `replicas` copies of the region list; only the last copy is cut into shrinking groups
        job = [(rep, g0, g1) for rep in range(replicas - 1) for (g0, g1) in plain] + [(replicas - 1, g0, g1) for (g0, g1) in last]
        work = None
        if seq_off is not None:
            w_plain, w_last = group_work(seq_off, regions.table, plain), group_work(seq_off, regions.table, last)
            work = np.concatenate([np.tile(w_plain, replicas - 1), w_last])
        n_job = len(job)
        if self.buffer is None or self.buffer.max_groups < n_job:
            self.buffer = GatherBuffer(self._capacity, self.world, self.rank, self.device, max_groups=n_job)
        self._calls += 1
        claimer = GroupClaimer(n_job, self.rank, self.world, self.schedule, work, key="pb_claim_%d" % self._calls)
        s = self.caller.stream(params, self.buffer.capacity, d_records=self.buffer.my_ptr())
        depth = 1                                 # groups claimed ahead of the one being run
        if from_files:
            from .frontend import FETCH_READERS
            depth = FETCH_READERS                 # fetches (pread + H2D + inflate) outstanding beside the group being run: one per reader of
                                                  # the source's rotation (the reader of the group just taken is free again)

            def stage(j):
                rep, g0, g1 = job[j]
                s.stage_device(source.take(j, g0, g1), 0, g1 - g0, rep * n_reg + g0)
        elif isinstance(source, DeviceReads):
            def stage(j):
                rep, g0, g1 = job[j]
                s.stage_device(source, g0, g1, rep * n_reg + g0)
        else:
            assert isinstance(source, HostReads)
            regs, keep = regions_array(regions)
            ref = np.ascontiguousarray(regions.ref, dtype=np.uint8)

            def stage(j):
                rep, g0, g1 = job[j]
                s.stage_host(source, regs, g0, g1, ref, rep * n_reg + g0)
        segments, seen = [], 0
        # a rank stops claiming when its slice of the gather buffer could not hold the groups in flight plus the one it would
        # claim: under the dynamic schedule a rank that started early may otherwise take more than its capacity; what it leaves is
        # claimed by the others
        worst = max(1, self.buffer.capacity // max(4, 2 * n_job // max(1, self.world)))
        ahead = deque()

        def claim():
            while len(ahead) < depth and self.buffer.capacity - seen >= (len(ahead) + 2.2) * worst:
                j = claimer.next()
                if j is None:
                    return
                if from_files:
                    source.request(j, job[j][1], job[j][2])
                ahead.append(j)
        claim()
        cur = ahead.popleft() if ahead else None
        if cur is not None:
            stage(cur)                        # (file source: takes this group's records before its reader is asked for another span)
        while cur is not None:
            claim()                           # claimed ahead: the next group's copies (or file fetch) overlap this group's kernels
            tot = s.run(flush=False)
            segments.append((cur, tot - seen))
            worst = max(worst, tot - seen)
            seen = tot
            cur = ahead.popleft() if ahead else None
            if cur is not None:
                stage(cur)                    # no sync: the next run() prepares its tables while this group's network runs
        n = s.end()
        self.groups_done = len(segments)
        t = self.caller.timings()
        self.stats = s.stats()
        self.buffer.set_mine(n, segments)
        self._ev[0].record()
        self.buffer.gather_meta()                 # tiny: completes when the slowest rank arrives -> the wait, not the transfer
        self._ev[1].record()
        self.buffer.gather_records()
        self._ev[2].record()
        torch.cuda.synchronize(self.device)
        self.phase_ms = dict(encoder_ms=t["encode_ms"], network_ms=t["network_ms"], wait_ms=self._ev[0].elapsed_time(self._ev[1]),
                             gather_ms=self._ev[1].elapsed_time(self._ev[2]), groups=len(segments))
        if self.schedule == "dynamic" and self.world > 1:
            done = int((self.buffer.meta[:, 1::2] >= 0).sum().item())
            if done != n_job:
                raise RuntimeError("only %d of %d groups were run: every rank's record capacity (%d) is exhausted — raise `capacity`"
                                   % (done, n_job, self.buffer.capacity))
        return n
