// Polish pileup-summary encoder for sm_100a.
//
// Re-design of SummaryGenerator::generate_summary / iterate_over_read / generate_image
// (pepper/modules/src/pileup_summary/summary_generator.cpp:47-121, 274-306, 370-393).  The reference
// accumulates fp64 counts in std::map<pair<long long,int>,double> keyed per (position, feature) and
// per (position, insert index, feature).  Here:
//
//   k_polish_prefix   one warp per read: (ref, read) prefix sums per CIGAR op (polish rules: N and P
//                     behave like D, summary_generator.cpp:99-114)
//   k_polish_count    one CTA per 512-position tile: shared-memory counters for the 10 base features,
//                     coverage and the longest insert per position; one write per position
//   k_polish_columns  per tile: position -> output column (prefix of 1 + longest_insert), writes
//                     genomic_pos (:381-388)
//   k_polish_inserts  one warp per read: insert bases -> int32 counters indexed by output column
//   k_polish_image    per column: (uint8)(int)((count / max(1.0, coverage)) * 254) in fp64 (:281,:295)
#include "handles.cuh"
#include <vector>
#include <algorithm>

namespace pb {

constexpr int PTILE = 512;
constexpr int PC_THREADS = 256;
constexpr int PLIST_CAP = 1024;

struct PReads {
    const int64_t *pos, *seq_off, *cigar_off;
    const uint8_t *flags, *mapq, *seq, *qual;
    const uint32_t *cigar;
    int64_t n_reads;
};

__global__ void k_polish_prefix(PReads R, int32_t *__restrict__ op_ref, int32_t *__restrict__ op_rd,
                                int32_t *__restrict__ read_reflen) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R.n_reads) return;
    const int64_t c0 = R.cigar_off[r], c1 = R.cigar_off[r + 1];
    int cref = 0, crd = 0;
    for (int64_t c = c0; c < c1; c += 32) {
        int dref = 0, drd = 0;
        if (c + lane < c1) {
            const uint32_t w = __ldg(R.cigar + c + lane);
            const int op = w & 15, len = (int) (w >> 4);
            switch (op) {
                case 0: case 7: case 8: dref = len; drd = len; break;
                case 1: case 4: drd = len; break;
                case 2: case 3: case 6: dref = len; break;
                default: break;
            }
        }
        int sref = dref, srd = drd;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, sref, d), b = __shfl_up_sync(0xffffffffu, srd, d);
            if (lane >= d) { sref += a; srd += b; }
        }
        if (c + lane < c1) { op_ref[c + lane] = cref + sref - dref; op_rd[c + lane] = crd + srd - drd; }
        cref += __shfl_sync(0xffffffffu, sref, 31);
        crd += __shfl_sync(0xffffffffu, srd, 31);
    }
    if (lane == 0) read_reflen[r] = cref;
}

// summary_generator.cpp:16-32: reverse A0 C1 G2 T3 other 8; forward A4 C5 G6 T7 other 9
__device__ __forceinline__ int polish_feature(int code, int rev) {
    int b;
    switch (code) { case 1: b = 0; break; case 2: b = 1; break; case 4: b = 2; break; case 8: b = 3; break; default: b = -1; }
    if (b < 0) return rev ? 8 : 9;
    return rev ? b : 4 + b;
}

struct PCountArgs {
    PReads R;
    const pb_region_t *regions;
    const int32_t *op_ref, *op_rd, *read_reflen;
    const int32_t *tile_region, *tile_x0;
    const int64_t *region_goff;
    int32_t *basecnt;     // [G][10]
    int32_t *cov;         // [G]
    int32_t *longest;     // [G]
    int32_t *tile_ncols;  // [n_tiles]
};

__global__ void __launch_bounds__(PC_THREADS, 2) k_polish_count(PCountArgs A) {
    __shared__ int32_t cnt[12 * PTILE];          // 0..9 features, 10 coverage, 11 longest insert
    __shared__ int s_cols;
    __shared__ int s_list[PLIST_CAP];
    __shared__ int s_nlist, s_next;
    const int t = blockIdx.x;
    const int reg = A.tile_region[t];
    const pb_region_t rg = A.regions[reg];
    const int64_t x0 = A.tile_x0[t];
    const int64_t L1 = rg.ref_end - rg.ref_start + 1;
    const int npos = (int) min((int64_t) PTILE, L1 - x0);
    const int64_t lo = rg.ref_start + x0, hi = lo + npos - 1;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 12 * PTILE; i += PC_THREADS) cnt[i] = 0;
    if (tid == 0) { s_cols = 0; s_nlist = 0; s_next = 0; }
    __syncthreads();
    const PReads &R = A.R;

    for (int64_t blk = rg.read_begin; blk < rg.read_end; blk += PLIST_CAP) {
        const int64_t blk_end = min(rg.read_end, blk + (int64_t) PLIST_CAP);
        for (int64_t rmine = blk + tid; rmine < blk_end; rmine += PC_THREADS) {
            if (__ldg(R.mapq + rmine) > 0) {                                     // :375
                const int64_t p0 = __ldg(R.pos + rmine);
                const int64_t p1 = p0 + __ldg(A.read_reflen + rmine);
                if ((p0 <= hi + 1) && (p1 >= lo - 1)) s_list[atomicAdd(&s_nlist, 1)] = (int) (rmine - blk);
            }
        }
        __syncthreads();
        const int nlist = s_nlist;
        while (true) {
            int item = 0;
            if (lane == 0) item = atomicAdd(&s_next, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= nlist) break;
            const int64_t r = blk + s_list[item];
            {
            const int64_t rpos = __ldg(R.pos + r);
            const int64_t so = R.seq_off[r];
            const int64_t lseq = R.seq_off[r + 1] - so;
            const int64_t c0 = R.cigar_off[r], c1 = R.cigar_off[r + 1];
            const int nops = (int) (c1 - c0);
            const int rev = __ldg(R.flags + r) & 1;
            const int64_t target = lo - rpos;
            int first = 0;
            if (target > 0) {
                int base = 0, n = nops;
                while (n > 1) {
                    const int stride = (n + 31) / 32;
                    const int idx = base + lane * stride;
                    const bool le = (lane * stride < n) && ((int64_t) __ldg(A.op_ref + c0 + idx) <= target);
                    const int k = __popc(__ballot_sync(0xffffffffu, le));
                    if (k == 0) { n = 0; break; }
                    const int nb = base + (k - 1) * stride;
                    n = min(stride, base + n - nb);
                    base = nb;
                }
                first = base;
            }
            for (int j0 = first; j0 < nops; j0 += 32) {
                const int j = j0 + lane;
                uint32_t w = 0; int pr = 0, pd = 0;
                if (j < nops) { w = __ldg(R.cigar + c0 + j); pr = __ldg(A.op_ref + c0 + j); pd = __ldg(A.op_rd + c0 + j); }
                const int op = (j < nops) ? (int) (w & 15) : 15;
                const int len = (int) (w >> 4);
                const int64_t a = rpos + pr;
                const bool live = (j < nops) && (a <= rg.ref_end);               // :54 break
                const int64_t a_last = __shfl_sync(0xffffffffu, a, 31);
                const bool is_m = (op == 0 || op == 7 || op == 8);
                const bool is_d = (op == 2 || op == 3 || op == 6);
                int64_t s0 = 0; int scnt = 0;
                if (live && (is_m || is_d)) {
                    const int64_t b0 = max(a, lo), b1 = min(a + len - 1, hi);
                    if (b1 >= b0) { s0 = b0; scnt = (int) (b1 - b0 + 1); }
                }
                if (live && is_d && a >= lo && a <= hi) {
                    // coverage[ref_position] += 1 once per deleted base inside the region (:105-110)
                    const int64_t b0 = max(a, rg.ref_start), b1 = min(a + len - 1, rg.ref_end);
                    if (b1 >= b0) atomicAdd(&cnt[10 * PTILE + (int) (a - lo)], (int) (b1 - b0 + 1));
                }
                if (live && op == 1) {
                    const int64_t p = a - 1;
                    if (p >= lo && p <= hi) {                                     // :83-84 (p within [ref_start, ref_end])
                        int64_t n = len;
                        if (pd + n > lseq) n = lseq - pd;
                        if (n < 0) n = 0;
                        atomicMax(&cnt[11 * PTILE + (int) (p - lo)], (int) n);
                    }
                }
                int incl = scnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                for (int k0 = 0; k0 < total; k0 += 32) {
                    const int idx = k0 + lane;
                    int l = 0;
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const int v = __shfl_sync(0xffffffffu, incl, l + step - 1);
                        if (v <= idx) l += step;
                    }
                    l = min(l, 31);
                    const int o_incl = __shfl_sync(0xffffffffu, incl, l);
                    const int o_cnt = __shfl_sync(0xffffffffu, scnt, l);
                    const int64_t o_s0 = __shfl_sync(0xffffffffu, s0, l);
                    const int64_t o_a = __shfl_sync(0xffffffffu, a, l);
                    const int o_pd = __shfl_sync(0xffffffffu, pd, l);
                    const int o_m = __shfl_sync(0xffffffffu, (int) is_m, l);
                    if (idx < total) {
                        const int64_t p = o_s0 + (idx - (o_incl - o_cnt));
                        const int x = (int) (p - lo);
                        if (o_m) {
                            const int code = seq_code_at(R.seq, so + o_pd + (p - o_a));
                            atomicAdd(&cnt[polish_feature(code, rev) * PTILE + x], 1);
                            atomicAdd(&cnt[10 * PTILE + x], 1);
                        } else {
                            atomicAdd(&cnt[(rev ? 8 : 9) * PTILE + x], 1);
                        }
                    }
                }
                if (a_last > hi + 1) break;
            }
            }
        }
        __syncthreads();
        if (tid == 0) { s_nlist = 0; s_next = 0; }
        __syncthreads();
    }
    const int64_t g0 = A.region_goff[reg] + x0;
    int cols = 0;
    for (int x = tid; x < npos; x += PC_THREADS) {
        const int64_t g = g0 + x;
#pragma unroll
        for (int f = 0; f < 10; f++) A.basecnt[g * 10 + f] = cnt[f * PTILE + x];
        A.cov[g] = cnt[10 * PTILE + x];
        A.longest[g] = cnt[11 * PTILE + x];
        cols += 1 + cnt[11 * PTILE + x];
    }
    atomicAdd(&s_cols, cols);
    __syncthreads();
    if (tid == 0) A.tile_ncols[t] = s_cols;
}

struct PColArgs {
    const pb_region_t *regions;
    const int32_t *tile_region, *tile_x0;
    const int64_t *region_goff;
    const int32_t *longest;
    const int64_t *tile_col_base;
    int64_t *col_of;       // [G] column of (pos, 0)
    int64_t *out_pos; int32_t *out_idx;
    int64_t *col_off;      // [n_regions+1]
    int64_t n_regions, total_cols, capacity;
};

__global__ void __launch_bounds__(PTILE) k_polish_columns(PColArgs A) {
    __shared__ int s_w[PTILE / 32];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int reg = A.tile_region[t];
    const pb_region_t rg = A.regions[reg];
    const int64_t x0 = A.tile_x0[t];
    const int64_t L1 = rg.ref_end - rg.ref_start + 1;
    const int npos = (int) min((int64_t) PTILE, L1 - x0);
    const int64_t g = A.region_goff[reg] + x0 + tid;
    const int mine = (tid < npos) ? 1 + A.longest[g] : 0;
    int inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; w++) base += s_w[w];
    if (tid < npos) {
        const int64_t c = A.tile_col_base[t] + base + inc - mine;
        A.col_of[g] = c;
        for (int k = 0; k < mine; k++)
            if (c + k < A.capacity) { A.out_pos[c + k] = rg.ref_start + x0 + tid; A.out_idx[c + k] = k; }
    }
    if (x0 == 0 && tid == 0) A.col_off[reg] = A.tile_col_base[t];
    if (t == 0 && tid == 0) A.col_off[A.n_regions] = A.total_cols;
}

struct PInsArgs {
    PReads R;
    const pb_region_t *regions;
    const int32_t *read_region;
    const int64_t *region_goff;
    const int32_t *op_ref, *op_rd;
    const int64_t *col_of;
    int32_t *inscnt;       // [total_cols][10], only insert columns are touched
};

__global__ void k_polish_inserts(PInsArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const PReads &R = A.R;
    if (r >= R.n_reads) return;
    if (R.mapq[r] == 0) return;
    const int reg = A.read_region[r];
    if (reg < 0) return;
    const pb_region_t rg = A.regions[reg];
    const int64_t so = R.seq_off[r], lseq = R.seq_off[r + 1] - so;
    const int64_t rpos = R.pos[r];
    const int64_t c0 = R.cigar_off[r], c1 = R.cigar_off[r + 1];
    const int rev = R.flags[r] & 1;
    for (int64_t cb = c0; cb < c1; cb += 32) {
        const int64_t c = cb + lane;
        uint32_t w = 0; int64_t a = 0; int pd = 0;
        bool is_ins = false;
        if (c < c1) {
            w = __ldg(R.cigar + c);
            a = rpos + __ldg(A.op_ref + c);
            pd = __ldg(A.op_rd + c);
            is_ins = ((w & 15) == 1) && (a <= rg.ref_end) && (a - 1 >= rg.ref_start) && (a - 1 <= rg.ref_end);
        }
        unsigned m = __ballot_sync(0xffffffffu, is_ins);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int len = (int) (__shfl_sync(0xffffffffu, w, src) >> 4);
            const int64_t aa = __shfl_sync(0xffffffffu, a, src);
            const int pdd = __shfl_sync(0xffffffffu, pd, src);
            int64_t n = len;
            if (pdd + n > lseq) n = lseq - pdd;
            const int64_t col0 = A.col_of[A.region_goff[reg] + (aa - 1 - rg.ref_start)] + 1;
            for (int64_t i = lane; i < n; i += 32) {
                const int code = seq_code_at(R.seq, so + pdd + i);
                atomicAdd(&A.inscnt[(col0 + i) * 10 + polish_feature(code, rev)], 1);
            }
        }
    }
}

struct PImgArgs {
    const int32_t *basecnt, *cov, *inscnt;
    const int64_t *col_of;
    const int32_t *longest;
    int64_t G;
    uint8_t *image;
    int64_t capacity;
};

// one thread per (position, feature): writes the base column and this position's insert columns
__global__ void k_polish_image(PImgArgs A) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.G * 10) return;
    const int64_t g = i / 10;
    const int f = (int) (i - g * 10);
    const double c = fmax(1.0, (double) A.cov[g]);
    const int64_t col = A.col_of[g];
    if (col >= A.capacity) return;
    {
        const double v = __dmul_rn(__ddiv_rn((double) A.basecnt[g * 10 + f], c), 254.0);
        A.image[col * 10 + f] = (uint8_t) (int32_t) v;       // double -> int32 truncation, low byte kept (gcc/x86-64)
    }
    const int n = A.longest[g];
    for (int k = 1; k <= n; k++) {
        if (col + k >= A.capacity) return;
        const double v = __dmul_rn(__ddiv_rn((double) A.inscnt[(col + k) * 10 + f], c), 254.0);
        A.image[(col + k) * 10 + f] = (uint8_t) (int32_t) v;
    }
}

}  // namespace pb

using namespace pb;


extern "C" int pb_polish_encoder_create(pb_polish_encoder_t **out, int device) {
    if (!out) { set_error("null out"); return PB_ERR_ARG; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= device) {
        set_error("no CUDA device %d (found %d): libpepper_b200 has no CPU fallback", device, n);
        return PB_ERR_CUDA;
    }
    PB_CUDA(cudaSetDevice(device));
    auto *e = new pb_polish_encoder();
    e->device = device;
    for (auto &ev : e->evt) PB_CUDA(cudaEventCreate(&ev));
    *out = e;
    return PB_OK;
}

extern "C" int pb_polish_encoder_destroy(pb_polish_encoder_t *e) {
    if (!e) return PB_OK;
    DevBuf *bufs[] = {&e->op_ref, &e->op_rd, &e->read_reflen, &e->read_region, &e->tile_region, &e->tile_x0, &e->region_goff,
                      &e->basecnt, &e->cov, &e->longest, &e->tile_ncols, &e->tile_col_base, &e->col_of, &e->inscnt, &e->scalars,
                      &e->h_pos, &e->h_seq_off, &e->h_cigar_off, &e->h_flags, &e->h_mapq, &e->h_seq, &e->h_qual, &e->h_cigar,
                      &e->h_regions, &e->o_image, &e->o_pos, &e->o_idx, &e->o_col_off, &e->p_image, &e->p_pos, &e->p_idx,
                      &e->p_col_off, &e->p_chunks, &e->p_imgs, &e->p_position, &e->p_index, &e->p_bases, &e->p_phred, &e->p_iregion,
                      &e->p_cid};
    for (auto *b : bufs) b->release();
    for (auto &ev : e->evt) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->pevt) if (ev) cudaEventDestroy(ev);
    delete e;
    return PB_OK;
}

extern "C" int pb_polish_encode_device(pb_polish_encoder_t *e, const pb_reads_t *dr, const pb_region_t *d_regions,
                                       int64_t n_regions, const pb_region_t *h_regions, int64_t capacity_cols,
                                       uint8_t *d_image, int64_t *d_pos, int32_t *d_idx, int64_t *d_col_off,
                                       int64_t *n_cols_out, void *stream_) {
    if (!e || !dr || !h_regions || !n_cols_out) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(e->device));
    *n_cols_out = 0;
    if (n_regions <= 0) return PB_OK;
    const int64_t n_reads = dr->n_reads;
    std::vector<int64_t> goff(n_regions + 1, 0);
    std::vector<int32_t> tile_region, tile_x0;
    for (int64_t r = 0; r < n_regions; r++) {
        const int64_t L1 = h_regions[r].ref_end - h_regions[r].ref_start + 1;
        if (L1 <= 0) { set_error("region %lld has ref_end < ref_start", (long long) r); return PB_ERR_ARG; }
        if (h_regions[r].read_begin < 0 || h_regions[r].read_end > n_reads || h_regions[r].read_begin > h_regions[r].read_end) {
            set_error("region %lld read range out of bounds", (long long) r); return PB_ERR_ARG;
        }
        goff[r + 1] = goff[r] + L1;
        for (int64_t x = 0; x < L1; x += PTILE) { tile_region.push_back((int32_t) r); tile_x0.push_back((int32_t) x); }
    }
    const int64_t G = goff[n_regions];
    const int64_t n_tiles = (int64_t) tile_region.size();
    std::vector<int32_t> read_region((size_t) std::max<int64_t>(n_reads, 1), -1);
    for (int64_t r = 0; r < n_regions; r++)
        for (int64_t i = h_regions[r].read_begin; i < h_regions[r].read_end; i++) read_region[i] = (int32_t) r;
    int64_t n_ops = 0;
    if (n_reads > 0) PB_CUDA(cudaMemcpyAsync(&n_ops, dr->cigar_off + n_reads, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));

    PB_TRY(e->op_ref.reserve(sizeof(int32_t) * (n_ops + 1)));
    PB_TRY(e->op_rd.reserve(sizeof(int32_t) * (n_ops + 1)));
    PB_TRY(e->read_reflen.reserve(sizeof(int32_t) * (n_reads + 1)));
    PB_TRY(e->read_region.reserve(sizeof(int32_t) * (n_reads + 1)));
    PB_TRY(e->tile_region.reserve(sizeof(int32_t) * n_tiles));
    PB_TRY(e->tile_x0.reserve(sizeof(int32_t) * n_tiles));
    PB_TRY(e->region_goff.reserve(sizeof(int64_t) * (n_regions + 1)));
    PB_TRY(e->basecnt.reserve(sizeof(int32_t) * 10 * G));
    PB_TRY(e->cov.reserve(sizeof(int32_t) * G));
    PB_TRY(e->longest.reserve(sizeof(int32_t) * G));
    PB_TRY(e->col_of.reserve(sizeof(int64_t) * G));
    PB_TRY(e->tile_ncols.reserve(sizeof(int32_t) * n_tiles));
    PB_TRY(e->tile_col_base.reserve(sizeof(int64_t) * (n_tiles + 1)));
    PB_TRY(e->scalars.reserve(sizeof(int64_t) * 4));
    PB_CUDA(cudaMemcpyAsync(e->tile_region.p, tile_region.data(), sizeof(int32_t) * n_tiles, cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemcpyAsync(e->tile_x0.p, tile_x0.data(), sizeof(int32_t) * n_tiles, cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemcpyAsync(e->region_goff.p, goff.data(), sizeof(int64_t) * (n_regions + 1), cudaMemcpyHostToDevice, st));
    if (n_reads > 0)
        PB_CUDA(cudaMemcpyAsync(e->read_region.p, read_region.data(), sizeof(int32_t) * n_reads, cudaMemcpyHostToDevice, st));

    PReads R{dr->pos, dr->seq_off, dr->cigar_off, dr->flags, dr->mapq, dr->seq, dr->qual, dr->cigar, n_reads};
    PB_CUDA(cudaEventRecord(e->evt[0], st));
    if (n_reads > 0)
        k_polish_prefix<<<(unsigned) ceil_div(n_reads, 8), 256, 0, st>>>(R, e->op_ref.as<int32_t>(), e->op_rd.as<int32_t>(),
                                                                       e->read_reflen.as<int32_t>());
    PB_CUDA(cudaEventRecord(e->evt[1], st));
    PCountArgs CA;
    CA.R = R; CA.regions = d_regions; CA.op_ref = e->op_ref.as<int32_t>(); CA.op_rd = e->op_rd.as<int32_t>();
    CA.read_reflen = e->read_reflen.as<int32_t>(); CA.tile_region = e->tile_region.as<int32_t>();
    CA.tile_x0 = e->tile_x0.as<int32_t>(); CA.region_goff = e->region_goff.as<int64_t>();
    CA.basecnt = e->basecnt.as<int32_t>(); CA.cov = e->cov.as<int32_t>(); CA.longest = e->longest.as<int32_t>();
    CA.tile_ncols = e->tile_ncols.as<int32_t>();
    k_polish_count<<<(unsigned) n_tiles, PC_THREADS, 0, st>>>(CA);
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaEventRecord(e->evt[2], st));
    int64_t *sc = e->scalars.as<int64_t>();
    k_scan_excl<<<1, 1024, 0, st>>>(CA.tile_ncols, e->tile_col_base.as<int64_t>(), n_tiles, sc);
    int64_t total_cols = 0;
    PB_CUDA(cudaMemcpyAsync(&total_cols, sc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    *n_cols_out = total_cols;
    if (total_cols > capacity_cols) {
        set_error("column capacity %lld < %lld needed", (long long) capacity_cols, (long long) total_cols);
        return PB_ERR_CAPACITY;
    }
    PB_TRY(e->inscnt.reserve(sizeof(int32_t) * 10 * (total_cols + 1)));
    PB_CUDA(cudaMemsetAsync(e->inscnt.p, 0, sizeof(int32_t) * 10 * total_cols, st));
    PColArgs PA;
    PA.regions = d_regions; PA.tile_region = CA.tile_region; PA.tile_x0 = CA.tile_x0; PA.region_goff = CA.region_goff;
    PA.longest = CA.longest; PA.tile_col_base = e->tile_col_base.as<int64_t>(); PA.col_of = e->col_of.as<int64_t>();
    PA.out_pos = d_pos; PA.out_idx = d_idx; PA.col_off = d_col_off; PA.n_regions = n_regions; PA.total_cols = total_cols;
    PA.capacity = capacity_cols;
    k_polish_columns<<<(unsigned) n_tiles, PTILE, 0, st>>>(PA);
    PInsArgs IA;
    IA.R = R; IA.regions = d_regions; IA.read_region = e->read_region.as<int32_t>(); IA.region_goff = CA.region_goff;
    IA.op_ref = CA.op_ref; IA.op_rd = CA.op_rd; IA.col_of = PA.col_of; IA.inscnt = e->inscnt.as<int32_t>();
    if (n_reads > 0) k_polish_inserts<<<(unsigned) ceil_div(n_reads, 8), 256, 0, st>>>(IA);
    PImgArgs MA;
    MA.basecnt = CA.basecnt; MA.cov = CA.cov; MA.inscnt = IA.inscnt; MA.col_of = PA.col_of; MA.longest = CA.longest;
    MA.G = G; MA.image = d_image; MA.capacity = capacity_cols;
    k_polish_image<<<(unsigned) ceil_div(G * 10, 256), 256, 0, st>>>(MA);
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaEventRecord(e->evt[3], st));
    PB_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 3; i++) cudaEventElapsedTime(&e->ms[i], e->evt[i], e->evt[i + 1]);
    return PB_OK;
}

extern "C" int pb_polish_encoder_timings(pb_polish_encoder_t *e, float *ms3) {
    if (!e || !ms3) return PB_ERR_ARG;
    for (int i = 0; i < 3; i++) ms3[i] = e->ms[i];
    return PB_OK;
}

extern "C" int pb_polish_encode_host(pb_polish_encoder_t *e, const pb_reads_t *h_reads, const pb_region_t *h_regions,
                                     int64_t n_regions, int64_t capacity_cols, uint8_t *h_image, int64_t *h_pos,
                                     int32_t *h_idx, int64_t *h_col_off, int64_t *n_cols_out, void *stream_) {
    if (!e || !h_reads || !h_regions || !n_cols_out) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(e->device));
    DevBuf *rb[8] = {&e->h_pos, &e->h_seq_off, &e->h_cigar_off, &e->h_flags, &e->h_mapq, &e->h_seq, &e->h_qual, &e->h_cigar};
    pb_reads_t d;
    PB_TRY(upload_reads(h_reads, rb, &d, st));
    PB_TRY(upload(e->h_regions, h_regions, sizeof(pb_region_t) * n_regions, st));
    const int64_t cap = std::max<int64_t>(capacity_cols, 1);
    PB_TRY(e->o_image.reserve((size_t) cap * 10));
    PB_TRY(e->o_pos.reserve(sizeof(int64_t) * cap));
    PB_TRY(e->o_idx.reserve(sizeof(int32_t) * cap));
    PB_TRY(e->o_col_off.reserve(sizeof(int64_t) * (n_regions + 1)));
    int rc = pb_polish_encode_device(e, &d, e->h_regions.as<pb_region_t>(), n_regions, h_regions, capacity_cols,
                                     e->o_image.as<uint8_t>(), e->o_pos.as<int64_t>(), e->o_idx.as<int32_t>(),
                                     e->o_col_off.as<int64_t>(), n_cols_out, stream_);
    if (rc != PB_OK) return rc;
    const int64_t n = *n_cols_out;
    if (n > 0) {
        PB_CUDA(cudaMemcpyAsync(h_image, e->o_image.p, (size_t) n * 10, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_pos, e->o_pos.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_idx, e->o_idx.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    }
    if (n_regions > 0)
        PB_CUDA(cudaMemcpyAsync(h_col_off, e->o_col_off.p, sizeof(int64_t) * (n_regions + 1), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}
