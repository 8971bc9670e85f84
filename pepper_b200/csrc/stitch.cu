// Polish stitch on the GPU: the reference builds a Python dict keyed by (position, index) per
// base (pepper/modules/python/Stitch.py:64-74) and sorts it; here every image column decides locally whether it survives
// (padding, region buffer, chunk-overlap winner, label 0), a scan gives its output offset, and the consensus string is
// written once.
#include "common.cuh"
#include <algorithm>

namespace pb {

constexpr int SCOLS = 1000;
constexpr int S_OVERLAP = 50;       // SEQ_OVERLAP, pepper Options.py
constexpr int S_BUFFER = 200;       // 2 * MIN_IMAGE_OVERLAP, Stitch.py:42

// str(a) < str(b) for non-negative decimal integers (chunk ids are sorted as strings, Stitch.py:50)
__device__ bool str_less(int a, int b) {
    char sa[12], sb[12];
    int la = 0, lb = 0;
    do { sa[la++] = (char) ('0' + a % 10); a /= 10; } while (a);
    do { sb[lb++] = (char) ('0' + b % 10); b /= 10; } while (b);
    for (int i = 0; i < la && i < lb; i++) {
        const char ca = sa[la - 1 - i], cb = sb[lb - 1 - i];
        if (ca != cb) return ca < cb;
    }
    return la < lb;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) k_stitch(const uint8_t *__restrict__ bases, const int64_t *__restrict__ position,
                                                const int32_t *__restrict__ index, const int32_t *__restrict__ image_region,
                                                const int32_t *__restrict__ chunk_id, const int64_t *__restrict__ region_starts,
                                                int64_t n_images, int32_t *__restrict__ counts, const int64_t *__restrict__ offsets,
                                                char *__restrict__ out, int64_t capacity) {
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    const int64_t im = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int reg = image_region[im], cid = chunk_id[im];
    const int64_t st = region_starts[reg];
    // which side of each 50-column overlap loses (the chunk whose id sorts earlier as a string is overwritten)
    const bool has_next = (im + 1 < n_images) && image_region[im + 1] == reg;
    const bool has_prev = (im > 0) && image_region[im - 1] == reg;
    const bool tail_lost = has_next && str_less(cid, chunk_id[im + 1]);
    const bool head_lost = has_prev && !str_less(chunk_id[im - 1], cid);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    int total = 0;
    for (int c0 = 0; c0 < SCOLS; c0 += 256) {
        const int c = c0 + tid;
        bool keep = false;
        uint8_t b = 0;
        if (c < SCOLS) {
            const int64_t pos = position[im * SCOLS + c];
            const int idx = index[im * SCOLS + c];
            b = bases[im * SCOLS + c];
            keep = pos >= 0 && idx >= 0 && !(st > 0 && pos <= st + S_BUFFER) && b != 0 && b <= 4;
            if (tail_lost && c >= SCOLS - S_OVERLAP) keep = false;
            if (head_lost && c < S_OVERLAP) keep = false;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int wcount = __popc(m), wprefix = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[warp] = wcount;
        __syncthreads();
        int base = s_carry;
        for (int w = 0; w < warp; w++) base += s_warp[w];
        if (WRITE && keep) {
            const int64_t o = offsets[im] + base + wprefix;
            if (o < capacity) out[o] = "ACGT"[b - 1];
        }
        __syncthreads();
        if (tid == 0) { int s = 0; for (int w = 0; w < 8; w++) s += s_warp[w]; s_carry += s; }
        __syncthreads();
        total = s_carry;
    }
    if (!WRITE && tid == 0) counts[im] = total;
}

}  // namespace pb

using namespace pb;

extern "C" int pb_polish_stitch_device(const uint8_t *d_bases, const int64_t *d_position, const int32_t *d_index,
                                       const int32_t *d_image_region, const int32_t *d_chunk_id, const int64_t *d_region_starts,
                                       int64_t n_images, char *d_out, int64_t capacity, int64_t *n_out, void *stream_) {
    if (!n_out) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    *n_out = 0;
    if (n_images <= 0) return PB_OK;
    DevBuf counts, offsets;
    PB_TRY(counts.reserve(sizeof(int32_t) * n_images));
    PB_TRY(offsets.reserve(sizeof(int64_t) * (n_images + 2)));
    k_stitch<false><<<(unsigned) n_images, 256, 0, st>>>(d_bases, d_position, d_index, d_image_region, d_chunk_id, d_region_starts, n_images,
                                                        counts.as<int32_t>(), nullptr, nullptr, 0);
    k_scan_excl<<<1, 1024, 0, st>>>(counts.as<int32_t>(), offsets.as<int64_t>(), n_images, offsets.as<int64_t>() + n_images + 1);
    int64_t total = 0;
    PB_CUDA(cudaMemcpyAsync(&total, offsets.as<int64_t>() + n_images, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    *n_out = total;
    int rc = PB_OK;
    if (total > capacity) {
        set_error("consensus capacity %lld < %lld needed", (long long) capacity, (long long) total);
        rc = PB_ERR_CAPACITY;
    } else if (total > 0) {
        k_stitch<true><<<(unsigned) n_images, 256, 0, st>>>(d_bases, d_position, d_index, d_image_region, d_chunk_id, d_region_starts, n_images,
                                                           nullptr, offsets.as<int64_t>(), d_out, capacity);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("stitch: %s", cudaGetErrorString(e)); rc = PB_ERR_CUDA; }
    }
    counts.release(); offsets.release();
    return rc;
}

extern "C" int pb_polish_stitch_host(const uint8_t *h_bases, const int64_t *h_position, const int32_t *h_index,
                                     const int32_t *h_image_region, const int32_t *h_chunk_id, const int64_t *h_region_starts,
                                     int64_t n_regions, int64_t n_images, char *h_out, int64_t capacity, int64_t *n_out, void *stream_) {
    if (!n_out) { set_error("null argument"); return PB_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { set_error("no CUDA device: libpepper_b200 has no CPU fallback"); return PB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t) stream_;
    *n_out = 0;
    if (n_images <= 0) return PB_OK;
    DevBuf b, p, i, r, c, s, o;
    PB_TRY(upload(b, h_bases, (size_t) n_images * SCOLS, st));
    PB_TRY(upload(p, h_position, sizeof(int64_t) * n_images * SCOLS, st));
    PB_TRY(upload(i, h_index, sizeof(int32_t) * n_images * SCOLS, st));
    PB_TRY(upload(r, h_image_region, sizeof(int32_t) * n_images, st));
    PB_TRY(upload(c, h_chunk_id, sizeof(int32_t) * n_images, st));
    PB_TRY(upload(s, h_region_starts, sizeof(int64_t) * n_regions, st));
    PB_TRY(o.reserve((size_t) std::max<int64_t>(capacity, 1)));
    int rc = pb_polish_stitch_device(b.as<uint8_t>(), p.as<int64_t>(), i.as<int32_t>(), r.as<int32_t>(), c.as<int32_t>(), s.as<int64_t>(),
                                     n_images, o.as<char>(), capacity, n_out, stream_);
    if (rc == PB_OK && *n_out > 0) {
        cudaError_t e = cudaMemcpyAsync(h_out, o.p, (size_t) *n_out, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("stitch copy back: %s", cudaGetErrorString(e)); rc = PB_ERR_CUDA; }
    }
    DevBuf *bufs[] = {&b, &p, &i, &r, &c, &s, &o};
    for (auto *x : bufs) x->release();
    return rc;
}
