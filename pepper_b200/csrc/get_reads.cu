// Region read fetch + trim on the device.
//
// Replaces the per-record body of BAM_handler::get_reads (pepper/modules/src/dataio/bam_handler.cpp:115-451; the
// pepper_variant copy is identical) for a batch of region queries over one coordinate-sorted contig resident in HBM.
//
// The reference walks every CIGAR base by base.  Here the walk is closed-form on per-op prefix arrays:
//   rpos_k  = reference position at the start of op k        (op_ref, int32 relative to the record's pos)
//   ridx_k  = read index at the start of op k                (op_rd)
//   cnt_k   = number of ops before k that can yield a tuple  (op_cnt: op in {M,I,D,N,S,=,X}, len > 0)
// and the facts (derived from the loop at bam_handler.cpp:176-303):
//   * the loop stops at the first op with rpos_k > stop (:186), so the ops it visits are a prefix;
//   * the anchor (pos_start != -1) is set by the first M/=/X op that has a base in [start, stop], i.e. the first such
//     op k0 with rpos_k0 + len > start and rpos_k0 <= stop; everything before it contributes nothing;
//   * every op after k0 that is visited (rpos_k <= stop) is kept: I/S whole, M/=/X and D/N cut at `stop` inclusive;
//     only the last visited op k1 can be cut; H/P/B and zero-length ops never yield a tuple;
//   * the kept bases are therefore ONE contiguous slice of the record's sequence, the kept tuples one contiguous
//     run of ops with the first and last lengths adjusted.
// A query returns a record iff the htslib 1.9 iterator would (overlap of [pos, pos + rlen) with [start, stop)), it
// passes the flag / MAPQ filters (:139-150) and keeps at least one base (:432).
#include "common.cuh"
#include <vector>
#include <algorithm>

using namespace pb;

struct pb_read_trimmer {
    int device = 0;
    // per record / per op
    DevBuf rec_end, rec_ok, op_ref, op_rd, op_cnt, scal;       // scal: [0] max span, [1] unsorted flag, [2..] totals
    // per interval
    DevBuf iv, iv_lo, cand_cnt, cand_off, kept_cnt;
    // per (interval, record) pair
    DevBuf p_rec, p_keep, p_k0, p_k1, p_b0, p_nb, p_nc, p_pos, keep_off, kept_pair;
    // emit
    DevBuf sel, out_pair, out_nb, out_nc;
    DevBuf o_pos, o_seq_off, o_cigar_off, o_flags, o_mapq, o_seq, o_qual, o_cigar;
    // staging of host records (pb_get_reads_plan_host)
    DevBuf h_pos, h_seq_off, h_cigar_off, h_flag, h_mapq, h_seq, h_qual, h_cigar;
    pb_records_t rec{};                   // device pointers of the planned batch
    std::vector<pb_interval_t> intervals;
    std::vector<int64_t> kept_off;        // [n_intervals + 1]
    int64_t n_pairs = 0, n_kept = 0;
    int64_t out_reads = 0, out_bases = 0, out_cigar = 0;
    bool planned = false;
};

namespace {

__device__ __forceinline__ bool op_is_match(int op) { return op == 0 || op == 7 || op == 8; }
__device__ __forceinline__ bool op_ref_consuming(int op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
__device__ __forceinline__ bool op_read_consuming(int op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }
__device__ __forceinline__ bool op_has_case(int op) { return op <= 4 || op == 7 || op == 8; }     // the switch at :189 (H: empty case)

// warp per record: exclusive prefixes of reference / read consumption and of tuple-eligible ops; record end; filters
__global__ void __launch_bounds__(256) k_rec_prefix(pb_records_t R, int include_supp, int min_mapq, int32_t *__restrict__ op_ref,
                                                    int32_t *__restrict__ op_rd, int32_t *__restrict__ op_cnt, int64_t *__restrict__ rec_end,
                                                    uint8_t *__restrict__ rec_ok, int64_t *__restrict__ scal) {
    const int64_t r = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R.n_records) return;
    const int64_t co = R.cigar_off[r], n = R.cigar_off[r + 1] - co;
    int ref = 0, rd = 0, cnt = 0;
    for (int64_t b = 0; b < n; b += 32) {
        const int64_t k = b + lane;
        int op = 15, len = 0;
        if (k < n) { const uint32_t c = R.cigar[co + k]; op = (int) (c & 15); len = (int) (c >> 4); }
        const int dr = op_ref_consuming(op) ? len : 0, dq = op_read_consuming(op) ? len : 0;
        const int de = (op_has_case(op) && op != 5 && len > 0) ? 1 : 0;
        int ir = dr, iq = dq, ie = de;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ur = __shfl_up_sync(0xffffffffu, ir, d), uq = __shfl_up_sync(0xffffffffu, iq, d), ue = __shfl_up_sync(0xffffffffu, ie, d);
            if (lane >= d) { ir += ur; iq += uq; ie += ue; }
        }
        if (k < n) { op_ref[co + k] = ref + ir - dr; op_rd[co + k] = rd + iq - dq; op_cnt[co + k] = cnt + ie - de; }
        ref += __shfl_sync(0xffffffffu, ir, 31); rd += __shfl_sync(0xffffffffu, iq, 31); cnt += __shfl_sync(0xffffffffu, ie, 31);
    }
    if (lane == 0) {
        const int64_t span = n ? ref : 1;                        // htslib bam_readrec: end = pos + (n_cigar ? rlen : 1)
        rec_end[r] = R.pos[r] + span;
        const int flag = R.flag[r];
        bool ok = !(flag & (512 | 1024 | 256 | 4));              // qc-fail, duplicate, secondary, unmapped (:139-142)
        if (!include_supp && (flag & 2048)) ok = false;          // :143
        if ((int) R.mapq[r] < min_mapq) ok = false;              // :148
        rec_ok[r] = ok ? 1 : 0;
        atomicMax((unsigned long long *) &scal[0], (unsigned long long) span);
        if (r > 0 && R.pos[r] < R.pos[r - 1]) scal[1] = 1;
    }
}

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t *__restrict__ a, int64_t n, int64_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

// thread per interval: candidate record range [lo, hi): pos < stop and pos > start - max_span
__global__ void k_iv_range(const int64_t *__restrict__ pos, int64_t n_rec, const pb_interval_t *__restrict__ iv, int64_t n_iv,
                           const int64_t *__restrict__ scal, int64_t *__restrict__ iv_lo, int32_t *__restrict__ cand_cnt) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_iv) return;
    const int64_t lo = lower_bound_i64(pos, n_rec, iv[i].start - scal[0] + 1), hi = lower_bound_i64(pos, n_rec, iv[i].stop);
    iv_lo[i] = lo;
    cand_cnt[i] = (int32_t) (hi > lo ? hi - lo : 0);
}

// thread per (interval, candidate record): overlap + filters + closed-form trim
__global__ void __launch_bounds__(256) k_pair_eval(pb_records_t R, const pb_interval_t *__restrict__ iv, int64_t n_iv,
                                                   const int64_t *__restrict__ iv_lo, const int64_t *__restrict__ cand_off, int64_t n_pairs,
                                                   const int32_t *__restrict__ op_ref, const int32_t *__restrict__ op_rd,
                                                   const int32_t *__restrict__ op_cnt, const int64_t *__restrict__ rec_end,
                                                   const uint8_t *__restrict__ rec_ok, int32_t *__restrict__ p_rec, int32_t *__restrict__ p_keep,
                                                   int32_t *__restrict__ p_k0, int32_t *__restrict__ p_k1, int32_t *__restrict__ p_b0,
                                                   int32_t *__restrict__ p_nb, int32_t *__restrict__ p_nc, int64_t *__restrict__ p_pos) {
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int64_t lo = 0, hi = n_iv;                                   // last interval with cand_off <= p
    while (hi - lo > 1) { const int64_t m = (lo + hi) >> 1; if (cand_off[m] <= p) lo = m; else hi = m; }
    const int64_t i = lo, r = iv_lo[i] + (p - cand_off[i]);
    const int64_t start = iv[i].start, stop = iv[i].stop;
    p_rec[p] = (int32_t) r;
    int keep = 0;
    if (rec_ok[r] && rec_end[r] > start) {
        const int64_t pos = R.pos[r], co = R.cigar_off[r];
        const int n = (int) (R.cigar_off[r + 1] - co);
        const int64_t s_rel = start - pos, e_rel = stop - pos;   // ops are tested in record-relative coordinates
        // first op whose reference end exceeds start (ends are non-decreasing in k)
        int a = 0, b = n;
        while (a < b) {
            const int m = (a + b) >> 1;
            const uint32_t c = R.cigar[co + m];
            const int64_t end = (int64_t) op_ref[co + m] + (op_ref_consuming((int) (c & 15)) ? (int64_t) (c >> 4) : 0);
            if (end > s_rel) b = m; else a = m + 1;
        }
        int k0 = -1;
        for (int k = a; k < n; k++) {
            const int64_t rp = op_ref[co + k];
            if (rp > e_rel) break;                               // :186
            const uint32_t c = R.cigar[co + k];
            if (op_is_match((int) (c & 15)) && (c >> 4) > 0 && rp + (int64_t) (c >> 4) > s_rel) { k0 = k; break; }
        }
        if (k0 >= 0) {
            // last visited op: last k with rpos_k <= stop
            int a1 = k0, b1 = n;
            while (b1 - a1 > 1) { const int m = (a1 + b1) >> 1; if ((int64_t) op_ref[co + m] <= e_rel) a1 = m; else b1 = m; }
            const int k1 = a1;
            const uint32_t c0 = R.cigar[co + k0], c1 = R.cigar[co + k1];
            const int64_t rp0 = op_ref[co + k0], rp1 = op_ref[co + k1];
            const int64_t skip = s_rel > rp0 ? s_rel - rp0 : 0;
            const int64_t b0 = (int64_t) op_rd[co + k0] + skip;
            const int op1 = (int) (c1 & 15);
            const int64_t len1 = c1 >> 4;
            int64_t bend = op_rd[co + k1];
            if (op_is_match(op1)) bend += min(len1, e_rel - rp1 + 1);
            else if (op1 == 1 || op1 == 4) bend += len1;
            const bool elig1 = op_has_case(op1) && op1 != 5 && len1 > 0;
            keep = 1;
            p_k0[p] = k0; p_k1[p] = k1; p_b0[p] = (int32_t) b0; p_nb[p] = (int32_t) (bend - b0);
            p_nc[p] = op_cnt[co + k1] - op_cnt[co + k0] + (elig1 ? 1 : 0);
            p_pos[p] = pos + max(rp0, s_rel);
            (void) c0;
        }
    }
    p_keep[p] = keep;
    if (!keep) { p_nb[p] = 0; p_nc[p] = 0; }
}

__global__ void k_compact(const int32_t *__restrict__ p_keep, const int64_t *__restrict__ keep_off, int64_t n_pairs,
                          int32_t *__restrict__ kept_pair) {
    const int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pairs && p_keep[p]) kept_pair[keep_off[p]] = (int32_t) p;
}

__global__ void k_kept_per_interval(const int64_t *__restrict__ cand_off, const int64_t *__restrict__ keep_off, int64_t n_iv,
                                    int64_t *__restrict__ kept_cnt) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_iv) kept_cnt[i] = keep_off[cand_off[i + 1]] - keep_off[cand_off[i]];
}

// out read j <- kept pair sel[j] (or j); gather its sizes for the offset scans
__global__ void k_out_gather(const int64_t *__restrict__ sel, const int32_t *__restrict__ kept_pair, const int32_t *__restrict__ p_nb,
                             const int32_t *__restrict__ p_nc, int64_t n_out, int32_t *__restrict__ out_pair,
                             int32_t *__restrict__ out_nb, int32_t *__restrict__ out_nc) {
    const int64_t j = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    const int32_t p = kept_pair[sel ? sel[j] : j];
    out_pair[j] = p; out_nb[j] = p_nb[p]; out_nc[j] = p_nc[p];
}

// warp per output read: copy the base slice (re-packing nibbles), the qualities and the adjusted CIGAR run
__global__ void __launch_bounds__(256) k_emit(pb_records_t R, const pb_interval_t *__restrict__ iv, int64_t n_iv,
                                              const int64_t *__restrict__ cand_off, const int32_t *__restrict__ out_pair, int64_t n_out,
                                              const int32_t *__restrict__ p_rec, const int32_t *__restrict__ p_k0, const int32_t *__restrict__ p_k1,
                                              const int32_t *__restrict__ p_b0, const int64_t *__restrict__ p_pos,
                                              const int32_t *__restrict__ op_ref, const int64_t *__restrict__ o_seq_off,
                                              const int64_t *__restrict__ o_cigar_off, int64_t *__restrict__ o_pos, uint8_t *__restrict__ o_flags,
                                              uint8_t *__restrict__ o_mapq, uint8_t *__restrict__ o_seq, uint8_t *__restrict__ o_qual,
                                              uint32_t *__restrict__ o_cigar) {
    const int64_t j = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= n_out) return;
    const int32_t p = out_pair[j];
    const int64_t r = p_rec[p];
    const int64_t so = R.seq_off[r] + p_b0[p];                   // first kept base (nibble index in the input)
    const int64_t oo = o_seq_off[j], nb = o_seq_off[j + 1] - oo;
    // qualities
    for (int64_t i = lane; i < nb; i += 32) o_qual[oo + i] = R.qual[so + i];
    // sequence: output bytes [oo/2 rounded up .. (oo+nb)/2) are owned by this read; a byte shared with the previous read is
    // written by THIS read (high nibble = the previous read's last base); a trailing half byte is written only by the last read
    const int64_t byte_lo = (oo + 1) >> 1, byte_hi = (oo + nb) >> 1;
    for (int64_t B = byte_lo + lane; B < byte_hi; B += 32) {
        const int64_t i = 2 * B - oo;
        o_seq[B] = (uint8_t) (seq_code_at(R.seq, so + i) << 4 | seq_code_at(R.seq, so + i + 1));
    }
    if (lane == 0) {
        if (oo & 1) {
            const int32_t pp = out_pair[j - 1];
            const int64_t pso = R.seq_off[p_rec[pp]] + p_b0[pp] + (oo - o_seq_off[j - 1]) - 1;   // previous read's last base
            o_seq[oo >> 1] = (uint8_t) (seq_code_at(R.seq, pso) << 4 | seq_code_at(R.seq, so));
        }
        if (j == n_out - 1 && ((oo + nb) & 1)) o_seq[(oo + nb) >> 1] = (uint8_t) (seq_code_at(R.seq, so + nb - 1) << 4);
        const int flag = R.flag[r];
        o_pos[j] = p_pos[p];
        o_flags[j] = (flag & 16) ? 1 : 0;                        // type_read_flags.is_reverse (:88)
        o_mapq[j] = R.mapq[r];
    }
    // CIGAR run k0..k1 with the first / last lengths adjusted; ops that yield no tuple are squeezed out
    int64_t lo = 0, hi = n_iv;
    while (hi - lo > 1) { const int64_t m = (lo + hi) >> 1; if (cand_off[m] <= p) lo = m; else hi = m; }
    const int64_t pos = R.pos[r], co = R.cigar_off[r];
    const int64_t s_rel = iv[lo].start - pos, e_rel = iv[lo].stop - pos;
    const int k0 = p_k0[p], k1 = p_k1[p];
    int64_t w = o_cigar_off[j];
    for (int kb = k0; kb <= k1; kb += 32) {
        const int k = kb + lane;
        uint32_t word = 0;
        bool emit = false;
        if (k <= k1) {
            const uint32_t c = R.cigar[co + k];
            const int op = (int) (c & 15);
            int64_t len = c >> 4;
            const int64_t rp = op_ref[co + k];
            if (op_ref_consuming(op)) {                          // M/=/X/D/N: cut at start (only k0 can begin before it) and at stop
                const int64_t a = max(rp, s_rel), b = min(rp + len - 1, e_rel);
                len = (k == k0 || k == k1) ? b - a + 1 : len;
            }
            emit = op_has_case(op) && op != 5 && len > 0;
            word = (uint32_t) (len << 4) | (uint32_t) op;
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        if (emit) o_cigar[w + __popc(m & ((1u << lane) - 1))] = word;
        w += __popc(m);
    }
}

int scan32(const DevBuf &in, DevBuf &out, int64_t n, int64_t *d_total, cudaStream_t st) {
    PB_TRY(out.reserve(sizeof(int64_t) * (n + 1)));
    k_scan_excl<<<1, 1024, 0, st>>>(in.as<int32_t>(), out.as<int64_t>(), n, d_total);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

}  // namespace

extern "C" int pb_read_trimmer_create(pb_read_trimmer_t **out, int device) {
    if (!out) { set_error("null out"); return PB_ERR_ARG; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= device) {
        set_error("no CUDA device %d (found %d): libpepper_b200 has no CPU fallback", device, n);
        return PB_ERR_CUDA;
    }
    PB_CUDA(cudaSetDevice(device));
    auto *t = new pb_read_trimmer();
    t->device = device;
    *out = t;
    return PB_OK;
}

extern "C" int pb_read_trimmer_destroy(pb_read_trimmer_t *t) {
    if (!t) return PB_OK;
    DevBuf *bufs[] = {&t->rec_end, &t->rec_ok, &t->op_ref, &t->op_rd, &t->op_cnt, &t->scal, &t->iv, &t->iv_lo, &t->cand_cnt, &t->cand_off,
                      &t->kept_cnt, &t->p_rec, &t->p_keep, &t->p_k0, &t->p_k1, &t->p_b0, &t->p_nb, &t->p_nc, &t->p_pos, &t->keep_off,
                      &t->kept_pair, &t->sel, &t->out_pair, &t->out_nb, &t->out_nc, &t->o_pos, &t->o_seq_off, &t->o_cigar_off, &t->o_flags,
                      &t->o_mapq, &t->o_seq, &t->o_qual, &t->o_cigar, &t->h_pos, &t->h_seq_off, &t->h_cigar_off, &t->h_flag, &t->h_mapq,
                      &t->h_seq, &t->h_qual, &t->h_cigar};
    for (auto *b : bufs) b->release();
    delete t;
    return PB_OK;
}

extern "C" int pb_get_reads_plan_device(pb_read_trimmer_t *t, const pb_records_t *dr, const pb_interval_t *h_iv, int64_t n_iv,
                                        const pb_get_reads_options_t *opt, int64_t *h_reads_per_interval, void *stream_) {
    if (!t || !dr || (!h_iv && n_iv) || !opt) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(t->device));
    t->planned = false;
    t->rec = *dr;
    t->intervals.assign(h_iv, h_iv + n_iv);
    t->kept_off.assign(n_iv + 1, 0);
    t->n_pairs = t->n_kept = 0;
    const int64_t n = dr->n_records;
    if (n == 0 || n_iv == 0) {
        for (int64_t i = 0; i < n_iv; i++) if (h_reads_per_interval) h_reads_per_interval[i] = 0;
        t->planned = true;
        return PB_OK;
    }
    // number of cigar ops: last offset lives on the device
    int64_t n_ops = 0;
    PB_CUDA(cudaMemcpyAsync(&n_ops, dr->cigar_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    PB_TRY(t->rec_end.reserve(sizeof(int64_t) * n));
    PB_TRY(t->rec_ok.reserve(n));
    PB_TRY(t->op_ref.reserve(sizeof(int32_t) * (n_ops + 1)));
    PB_TRY(t->op_rd.reserve(sizeof(int32_t) * (n_ops + 1)));
    PB_TRY(t->op_cnt.reserve(sizeof(int32_t) * (n_ops + 1)));
    PB_TRY(t->scal.reserve(sizeof(int64_t) * 8));
    PB_CUDA(cudaMemsetAsync(t->scal.p, 0, sizeof(int64_t) * 8, st));
    k_rec_prefix<<<(unsigned) ceil_div(n * 32, 256), 256, 0, st>>>(*dr, opt->include_supplementary, opt->min_mapq, t->op_ref.as<int32_t>(),
                                                                   t->op_rd.as<int32_t>(), t->op_cnt.as<int32_t>(), t->rec_end.as<int64_t>(),
                                                                   t->rec_ok.as<uint8_t>(), t->scal.as<int64_t>());
    PB_CUDA(cudaGetLastError());
    PB_TRY(upload(t->iv, h_iv, sizeof(pb_interval_t) * n_iv, st));
    PB_TRY(t->iv_lo.reserve(sizeof(int64_t) * n_iv));
    PB_TRY(t->cand_cnt.reserve(sizeof(int32_t) * n_iv));
    k_iv_range<<<(unsigned) ceil_div(n_iv, 256), 256, 0, st>>>(dr->pos, n, t->iv.as<pb_interval_t>(), n_iv, t->scal.as<int64_t>(),
                                                               t->iv_lo.as<int64_t>(), t->cand_cnt.as<int32_t>());
    PB_CUDA(cudaGetLastError());
    PB_TRY(scan32(t->cand_cnt, t->cand_off, n_iv, t->scal.as<int64_t>() + 2, st));
    int64_t h_scal[4];
    PB_CUDA(cudaMemcpyAsync(h_scal, t->scal.p, sizeof(h_scal), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    if (h_scal[1]) { set_error("records are not sorted by position (get_reads needs a coordinate-sorted contig)"); return PB_ERR_ARG; }
    const int64_t P = h_scal[2];
    if (P > 0x7fffffffLL) { set_error("too many (interval, record) pairs: %lld", (long long) P); return PB_ERR_ARG; }
    t->n_pairs = P;
    if (P > 0) {
        DevBuf *i32[] = {&t->p_rec, &t->p_keep, &t->p_k0, &t->p_k1, &t->p_b0, &t->p_nb, &t->p_nc, &t->kept_pair};
        for (auto *b : i32) PB_TRY(b->reserve(sizeof(int32_t) * P));
        PB_TRY(t->p_pos.reserve(sizeof(int64_t) * P));
        k_pair_eval<<<(unsigned) ceil_div(P, 256), 256, 0, st>>>(*dr, t->iv.as<pb_interval_t>(), n_iv, t->iv_lo.as<int64_t>(),
                                                                 t->cand_off.as<int64_t>(), P, t->op_ref.as<int32_t>(), t->op_rd.as<int32_t>(),
                                                                 t->op_cnt.as<int32_t>(), t->rec_end.as<int64_t>(), t->rec_ok.as<uint8_t>(),
                                                                 t->p_rec.as<int32_t>(), t->p_keep.as<int32_t>(), t->p_k0.as<int32_t>(),
                                                                 t->p_k1.as<int32_t>(), t->p_b0.as<int32_t>(), t->p_nb.as<int32_t>(),
                                                                 t->p_nc.as<int32_t>(), t->p_pos.as<int64_t>());
        PB_CUDA(cudaGetLastError());
        PB_TRY(scan32(t->p_keep, t->keep_off, P, t->scal.as<int64_t>() + 3, st));
        k_compact<<<(unsigned) ceil_div(P, 256), 256, 0, st>>>(t->p_keep.as<int32_t>(), t->keep_off.as<int64_t>(), P, t->kept_pair.as<int32_t>());
        PB_CUDA(cudaGetLastError());
        PB_TRY(t->kept_cnt.reserve(sizeof(int64_t) * n_iv));
        k_kept_per_interval<<<(unsigned) ceil_div(n_iv, 256), 256, 0, st>>>(t->cand_off.as<int64_t>(), t->keep_off.as<int64_t>(), n_iv,
                                                                           t->kept_cnt.as<int64_t>());
        PB_CUDA(cudaGetLastError());
        std::vector<int64_t> cnt(n_iv);
        PB_CUDA(cudaMemcpyAsync(cnt.data(), t->kept_cnt.p, sizeof(int64_t) * n_iv, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        for (int64_t i = 0; i < n_iv; i++) t->kept_off[i + 1] = t->kept_off[i] + cnt[i];
        t->n_kept = t->kept_off[n_iv];
    }
    if (h_reads_per_interval)
        for (int64_t i = 0; i < n_iv; i++) h_reads_per_interval[i] = t->kept_off[i + 1] - t->kept_off[i];
    t->planned = true;
    return PB_OK;
}

extern "C" int pb_get_reads_plan_host(pb_read_trimmer_t *t, const pb_records_t *h, const pb_interval_t *h_iv, int64_t n_iv,
                                      const pb_get_reads_options_t *opt, int64_t *h_reads_per_interval, void *stream_) {
    if (!t || !h) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(t->device));
    const int64_t n = h->n_records;
    const int64_t nb = n ? h->seq_off[n] : 0, nc = n ? h->cigar_off[n] : 0;
    PB_TRY(upload(t->h_pos, h->pos, sizeof(int64_t) * n, st));
    PB_TRY(upload(t->h_seq_off, h->seq_off, sizeof(int64_t) * (n + 1), st));
    PB_TRY(upload(t->h_cigar_off, h->cigar_off, sizeof(int64_t) * (n + 1), st));
    PB_TRY(upload(t->h_flag, h->flag, sizeof(uint16_t) * n, st));
    PB_TRY(upload(t->h_mapq, h->mapq, n, st));
    PB_TRY(upload(t->h_seq, h->seq, (size_t) ((nb + 1) / 2), st));
    PB_TRY(upload(t->h_qual, h->qual, (size_t) nb, st));
    PB_TRY(upload(t->h_cigar, h->cigar, sizeof(uint32_t) * nc, st));
    pb_records_t d;
    d.n_records = n;
    d.pos = t->h_pos.as<int64_t>(); d.seq_off = t->h_seq_off.as<int64_t>(); d.cigar_off = t->h_cigar_off.as<int64_t>();
    d.flag = t->h_flag.as<uint16_t>(); d.mapq = t->h_mapq.as<uint8_t>(); d.seq = t->h_seq.as<uint8_t>();
    d.qual = t->h_qual.as<uint8_t>(); d.cigar = t->h_cigar.as<uint32_t>();
    return pb_get_reads_plan_device(t, &d, h_iv, n_iv, opt, h_reads_per_interval, stream_);
}

extern "C" int pb_get_reads_emit_device(pb_read_trimmer_t *t, const int64_t *h_select_off, const int32_t *h_select, pb_reads_t *out,
                                        int64_t *h_read_begin, int64_t *h_read_end, void *stream_) {
    if (!t || !out) { set_error("null argument"); return PB_ERR_ARG; }
    if (!t->planned) { set_error("pb_get_reads_emit_device without a plan"); return PB_ERR_STATE; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(t->device));
    const int64_t n_iv = (int64_t) t->intervals.size();
    int64_t n_out = t->n_kept;
    const int64_t *d_sel = nullptr;
    std::vector<int64_t> sel;
    if (h_select_off) {
        if (!h_select && h_select_off[n_iv]) { set_error("h_select is null"); return PB_ERR_ARG; }
        n_out = h_select_off[n_iv];
        sel.resize(n_out);
        for (int64_t i = 0; i < n_iv; i++) {
            const int64_t cnt = t->kept_off[i + 1] - t->kept_off[i];
            for (int64_t s = h_select_off[i]; s < h_select_off[i + 1]; s++) {
                if (h_select[s] < 0 || h_select[s] >= cnt) { set_error("selection index %d out of range for interval %lld (%lld reads)", h_select[s], (long long) i, (long long) cnt); return PB_ERR_ARG; }
                sel[s] = t->kept_off[i] + h_select[s];
            }
            if (h_read_begin) h_read_begin[i] = h_select_off[i];
            if (h_read_end) h_read_end[i] = h_select_off[i + 1];
        }
        PB_TRY(upload(t->sel, sel.data(), sizeof(int64_t) * n_out, st));
        d_sel = t->sel.as<int64_t>();
    } else {
        for (int64_t i = 0; i < n_iv; i++) {
            if (h_read_begin) h_read_begin[i] = t->kept_off[i];
            if (h_read_end) h_read_end[i] = t->kept_off[i + 1];
        }
    }
    t->out_reads = n_out; t->out_bases = 0; t->out_cigar = 0;
    PB_TRY(t->o_seq_off.reserve(sizeof(int64_t) * (n_out + 1)));
    PB_TRY(t->o_cigar_off.reserve(sizeof(int64_t) * (n_out + 1)));
    PB_TRY(t->o_pos.reserve(sizeof(int64_t) * (n_out + 1)));
    PB_TRY(t->o_flags.reserve(n_out + 1));
    PB_TRY(t->o_mapq.reserve(n_out + 1));
    if (n_out > 0) {
        PB_TRY(t->out_pair.reserve(sizeof(int32_t) * n_out));
        PB_TRY(t->out_nb.reserve(sizeof(int32_t) * n_out));
        PB_TRY(t->out_nc.reserve(sizeof(int32_t) * n_out));
        k_out_gather<<<(unsigned) ceil_div(n_out, 256), 256, 0, st>>>(d_sel, t->kept_pair.as<int32_t>(), t->p_nb.as<int32_t>(), t->p_nc.as<int32_t>(),
                                                                      n_out, t->out_pair.as<int32_t>(), t->out_nb.as<int32_t>(), t->out_nc.as<int32_t>());
        PB_CUDA(cudaGetLastError());
        PB_TRY(scan32(t->out_nb, t->o_seq_off, n_out, t->scal.as<int64_t>() + 4, st));
        PB_TRY(scan32(t->out_nc, t->o_cigar_off, n_out, t->scal.as<int64_t>() + 5, st));
        int64_t tot[2];
        PB_CUDA(cudaMemcpyAsync(tot, t->scal.as<int64_t>() + 4, sizeof(tot), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        t->out_bases = tot[0]; t->out_cigar = tot[1];
    } else {
        PB_CUDA(cudaMemsetAsync(t->o_seq_off.p, 0, sizeof(int64_t), st));
        PB_CUDA(cudaMemsetAsync(t->o_cigar_off.p, 0, sizeof(int64_t), st));
    }
    PB_TRY(t->o_seq.reserve((size_t) (t->out_bases + 1) / 2 + 16));
    PB_TRY(t->o_qual.reserve((size_t) t->out_bases + 16));
    PB_TRY(t->o_cigar.reserve(sizeof(uint32_t) * (t->out_cigar + 4)));
    if (n_out > 0) {
        k_emit<<<(unsigned) ceil_div(n_out * 32, 256), 256, 0, st>>>(t->rec, t->iv.as<pb_interval_t>(), n_iv, t->cand_off.as<int64_t>(),
                                                                     t->out_pair.as<int32_t>(), n_out, t->p_rec.as<int32_t>(), t->p_k0.as<int32_t>(),
                                                                     t->p_k1.as<int32_t>(), t->p_b0.as<int32_t>(), t->p_pos.as<int64_t>(),
                                                                     t->op_ref.as<int32_t>(), t->o_seq_off.as<int64_t>(), t->o_cigar_off.as<int64_t>(),
                                                                     t->o_pos.as<int64_t>(), t->o_flags.as<uint8_t>(), t->o_mapq.as<uint8_t>(),
                                                                     t->o_seq.as<uint8_t>(), t->o_qual.as<uint8_t>(), t->o_cigar.as<uint32_t>());
        PB_CUDA(cudaGetLastError());
    }
    out->n_reads = n_out;
    out->pos = t->o_pos.as<int64_t>(); out->seq_off = t->o_seq_off.as<int64_t>(); out->cigar_off = t->o_cigar_off.as<int64_t>();
    out->flags = t->o_flags.as<uint8_t>(); out->mapq = t->o_mapq.as<uint8_t>(); out->seq = t->o_seq.as<uint8_t>();
    out->qual = t->o_qual.as<uint8_t>(); out->cigar = t->o_cigar.as<uint32_t>();
    return PB_OK;
}

extern "C" int pb_get_reads_sizes(pb_read_trimmer_t *t, int64_t *sizes) {
    if (!t || !sizes) { set_error("null argument"); return PB_ERR_ARG; }
    sizes[0] = t->out_reads; sizes[1] = t->out_bases; sizes[2] = t->out_cigar;
    return PB_OK;
}

extern "C" int pb_get_reads_fetch(pb_read_trimmer_t *t, int64_t *pos, int64_t *seq_off, int64_t *cigar_off, uint8_t *flags, uint8_t *mapq,
                                  uint8_t *seq, uint8_t *qual, uint32_t *cigar, void *stream_) {
    if (!t) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(t->device));
    const int64_t n = t->out_reads;
    if (pos && n) PB_CUDA(cudaMemcpyAsync(pos, t->o_pos.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
    if (seq_off) PB_CUDA(cudaMemcpyAsync(seq_off, t->o_seq_off.p, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, st));
    if (cigar_off) PB_CUDA(cudaMemcpyAsync(cigar_off, t->o_cigar_off.p, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, st));
    if (flags && n) PB_CUDA(cudaMemcpyAsync(flags, t->o_flags.p, n, cudaMemcpyDeviceToHost, st));
    if (mapq && n) PB_CUDA(cudaMemcpyAsync(mapq, t->o_mapq.p, n, cudaMemcpyDeviceToHost, st));
    if (seq && t->out_bases) PB_CUDA(cudaMemcpyAsync(seq, t->o_seq.p, (size_t) (t->out_bases + 1) / 2, cudaMemcpyDeviceToHost, st));
    if (qual && t->out_bases) PB_CUDA(cudaMemcpyAsync(qual, t->o_qual.p, (size_t) t->out_bases, cudaMemcpyDeviceToHost, st));
    if (cigar && t->out_cigar) PB_CUDA(cudaMemcpyAsync(cigar, t->o_cigar.p, sizeof(uint32_t) * t->out_cigar, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}
