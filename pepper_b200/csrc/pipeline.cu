// Fused make_images + inference entry points (the path `pepper_variant call_variant` steps 1+2 and
// `pepper polish` steps 1+2 take, without the intermediate image HDF5 files): encoder output stays in HBM and
// feeds the network directly.
#include "handles.cuh"
#include <algorithm>
#include <vector>
#include <stdint.h>

using namespace pb;

namespace pb {

// AlignmentSummarizer.chunk_images (pepper AlignmentSummarizer.py:19-56) on the device:
// image k = columns [start, start+nvalid) of its region, zero padded to 1000, positions padded with (-1,-1)
struct ChunkRow { int64_t col_start; int32_t nvalid; int32_t region; int32_t chunk_id; int32_t pad; };

__global__ void k_polish_chunk(const ChunkRow *__restrict__ rows, const uint8_t *__restrict__ image, const int64_t *__restrict__ pos,
                               const int32_t *__restrict__ idx, uint8_t *__restrict__ imgs, int64_t *__restrict__ position,
                               int32_t *__restrict__ index, int32_t *__restrict__ iregion, int32_t *__restrict__ cid) {
    const int64_t k = blockIdx.x;
    const ChunkRow r = rows[k];
    for (int i = threadIdx.x; i < PB_POLISH_SEQ_LEN * PB_POLISH_FEATURES; i += blockDim.x) {
        const int c = i / PB_POLISH_FEATURES;
        imgs[k * PB_POLISH_SEQ_LEN * PB_POLISH_FEATURES + i] = (c < r.nvalid) ? image[(r.col_start + c) * PB_POLISH_FEATURES + (i - c * PB_POLISH_FEATURES)] : 0;
    }
    for (int c = threadIdx.x; c < PB_POLISH_SEQ_LEN; c += blockDim.x) {
        position[k * PB_POLISH_SEQ_LEN + c] = (c < r.nvalid) ? pos[r.col_start + c] : -1;
        index[k * PB_POLISH_SEQ_LEN + c] = (c < r.nvalid) ? idx[r.col_start + c] : -1;
    }
    if (threadIdx.x == 0) { iregion[k] = r.region; cid[k] = r.chunk_id; }
}

}  // namespace pb

static int ensure_events(cudaEvent_t *ev, int n) {
    for (int i = 0; i < n; i++) if (!ev[i]) PB_CUDA(cudaEventCreate(&ev[i]));
    return PB_OK;
}

extern "C" int pb_variant_call_device(pb_variant_encoder_t *enc, pb_variant_net_t *net, const pb_reads_t *d_reads,
                                      const pb_region_t *d_regions, int64_t n_regions, const pb_region_t *h_regions,
                                      const char *d_ref, int64_t ref_bytes, const pb_variant_params_t *params,
                                      int64_t capacity, int8_t *d_images, int64_t *d_positions, uint8_t *d_depths,
                                      uint8_t *d_freqs, char *d_keys, int32_t *d_region_of, float *d_probs,
                                      int64_t *n_out, void *stream_) {
    if (!enc || !net) { set_error("null handle"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_TRY(ensure_events(enc->pevt, 3));
    PB_CUDA(cudaEventRecord(enc->pevt[0], st));
    PB_TRY(pb_variant_encode_device(enc, d_reads, d_regions, n_regions, h_regions, d_ref, ref_bytes, params, capacity,
                                    d_images, d_positions, d_depths, d_freqs, d_keys, d_region_of, nullptr, n_out, stream_));
    PB_CUDA(cudaEventRecord(enc->pevt[1], st));
    PB_TRY(pb_variant_net_forward_device(net, d_images, *n_out, d_probs, nullptr, stream_));
    PB_CUDA(cudaEventRecord(enc->pevt[2], st));
    PB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&enc->pms[0], enc->pevt[0], enc->pevt[1]);
    cudaEventElapsedTime(&enc->pms[1], enc->pevt[1], enc->pevt[2]);
    return PB_OK;
}

// One group of regions [g0, g1) staged into buffer set `b` on the copy stream.  Read / reference arrays keep their
// ABSOLUTE offsets: the device base pointers are shifted back by the slice start ("virtual base"), so seq_off /
// cigar_off / ref_off need no rebasing; only the per-read arrays are re-indexed from the group's first read.
struct GroupView { pb_reads_t d; const pb_region_t *d_regions; const char *d_ref; std::vector<pb_region_t> h_regions; };

static int stage_group(pb_variant_encoder_t *e, int b, const pb_reads_t *h, const pb_region_t *h_regions, int64_t g0, int64_t g1,
                       const char *h_ref, GroupView &V) {
    cudaStream_t cs = e->copy_stream;
    DevBuf *B = e->g_buf[b];
    const int64_t r0 = h_regions[g0].read_begin, r1 = h_regions[g1 - 1].read_end;
    const int64_t n = r1 - r0;
    const int64_t nb0 = h->seq_off[r0], nb1 = h->seq_off[r1], c0 = h->cigar_off[r0], c1 = h->cigar_off[r1];
    const int64_t sb0 = nb0 >> 1, sb1 = (nb1 + 1) >> 1;
    PB_TRY(upload(B[0], h->pos + r0, sizeof(int64_t) * n, cs));
    PB_TRY(upload(B[1], h->seq_off + r0, sizeof(int64_t) * (n + 1), cs));
    PB_TRY(upload(B[2], h->cigar_off + r0, sizeof(int64_t) * (n + 1), cs));
    PB_TRY(upload(B[3], h->flags + r0, n, cs));
    PB_TRY(upload(B[4], h->mapq + r0, n, cs));
    PB_TRY(upload(B[5], h->seq + sb0, (size_t) (sb1 - sb0), cs));
    PB_TRY(upload(B[6], h->qual + nb0, (size_t) (nb1 - nb0), cs));
    PB_TRY(upload(B[7], h->cigar + c0, sizeof(uint32_t) * (c1 - c0), cs));
    V.h_regions.assign(h_regions + g0, h_regions + g1);
    int64_t ref0 = INT64_MAX, ref1 = 0;
    for (auto &rg : V.h_regions) {
        rg.read_begin -= r0; rg.read_end -= r0;
        ref0 = std::min(ref0, rg.ref_off); ref1 = std::max(ref1, rg.ref_off + rg.ref_len);
    }
    if (ref1 < ref0) { ref0 = 0; ref1 = 0; }
    PB_TRY(upload(B[8], V.h_regions.data(), sizeof(pb_region_t) * V.h_regions.size(), cs));
    PB_TRY(upload(B[9], h_ref + ref0, (size_t) (ref1 - ref0), cs));
    PB_CUDA(cudaEventRecord(e->copied[b], cs));
    V.d.n_reads = n;
    V.d.pos = B[0].as<int64_t>(); V.d.seq_off = B[1].as<int64_t>(); V.d.cigar_off = B[2].as<int64_t>();
    V.d.flags = B[3].as<uint8_t>(); V.d.mapq = B[4].as<uint8_t>();
    V.d.seq = B[5].as<uint8_t>() - sb0; V.d.qual = B[6].as<uint8_t>() - nb0; V.d.cigar = B[7].as<uint32_t>() - c0;
    V.d_regions = B[8].as<pb_region_t>();
    V.d_ref = B[9].as<char>() - ref0;
    return PB_OK;
}

__global__ void k_add_offset_i32(int32_t *p, int64_t n, int32_t off) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] += off;
}

// single-shot host entry: everything copied up front (small batches, or regions whose read ranges are not ascending)
static int variant_call_host_single(pb_variant_encoder_t *e, pb_variant_net_t *net, const pb_reads_t *h_reads,
                                    const pb_region_t *h_regions, int64_t n_regions, const char *h_ref, int64_t ref_bytes,
                                    const pb_variant_params_t *params, int64_t capacity, int8_t *h_images,
                                    int64_t *h_positions, uint8_t *h_depths, uint8_t *h_freqs, char *h_keys,
                                    int32_t *h_region_of, float *h_probs, int64_t *n_out, void *stream_) {
    cudaStream_t st = (cudaStream_t) stream_;
    DevBuf *rb[8] = {&e->h_pos, &e->h_seq_off, &e->h_cigar_off, &e->h_flags, &e->h_mapq, &e->h_seq, &e->h_qual, &e->h_cigar};
    pb_reads_t d;
    PB_TRY(upload_reads(h_reads, rb, &d, st));
    PB_TRY(upload(e->h_regions, h_regions, sizeof(pb_region_t) * n_regions, st));
    PB_TRY(upload(e->h_ref, h_ref, (size_t) ref_bytes, st));
    const int64_t cap = std::max<int64_t>(capacity, 1);
    PB_TRY(e->p_images.reserve((size_t) cap * 33 * 26));
    PB_TRY(e->p_positions.reserve(sizeof(int64_t) * cap));
    PB_TRY(e->p_depths.reserve(cap));
    PB_TRY(e->p_freqs.reserve(cap));
    PB_TRY(e->p_keys.reserve((size_t) cap * PB_ALLELE_STRIDE));
    PB_TRY(e->p_region_of.reserve(sizeof(int32_t) * cap));
    PB_TRY(e->p_probs.reserve(sizeof(float) * 3 * cap));
    PB_TRY(pb_variant_call_device(e, net, &d, e->h_regions.as<pb_region_t>(), n_regions, h_regions, e->h_ref.as<char>(), ref_bytes,
                                  params, capacity, e->p_images.as<int8_t>(), e->p_positions.as<int64_t>(), e->p_depths.as<uint8_t>(),
                                  e->p_freqs.as<uint8_t>(), e->p_keys.as<char>(), e->p_region_of.as<int32_t>(), e->p_probs.as<float>(),
                                  n_out, stream_));
    const int64_t n = *n_out;
    if (n > 0) {
        if (h_images) PB_CUDA(cudaMemcpyAsync(h_images, e->p_images.p, (size_t) n * 33 * 26, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_positions, e->p_positions.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_depths, e->p_depths.p, n, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_freqs, e->p_freqs.p, n, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_keys, e->p_keys.p, (size_t) n * PB_ALLELE_STRIDE, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_region_of, e->p_region_of.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_probs, e->p_probs.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    }
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------- streaming session
// The unit of work is a GROUP of regions.  A session accumulates the candidates of the groups pushed so far in the
// library's device buffers, runs the network over WHOLE 9,472-candidate chunks of the accumulated candidates as they fill
// (the remainder rides along to the next group, so only the very last chunk of a session is partial), and lets the caller
// stage group g+1 (copy stream, pinned source) while the kernels of group g run.  pb_variant_call_host is a session over a
// fixed group list; the multi-GPU callers (pepper_b200/dist.py) push whatever group a rank claims next.
constexpr int64_t CALL_GROUP = 96;
constexpr int64_t CALL_FIRST_GROUP = 24;

// network chunk of the pipelined host entry == VARIANT_CHUNK of nets.cu (74 row tiles of 128 candidates)
static constexpr int64_t PIPE_NET_CHUNK = 9472;

namespace pb {
struct VariantStream {
    bool active = false;
    pb_variant_params_t params;
    int64_t capacity = 0, done = 0, net_done = 0;
    pb_pred_record_t *d_records = nullptr;
    cudaStream_t st = nullptr;
    GroupView V[2];
    int32_t region_id0[2] = {0, 0};
    bool staged[2] = {false, false}, from_host[2] = {false, false};
    int next_stage = 0, next_run = 0;
    // one event triple per group (encoder start, encoder end = network start, network end): read back at _sync / _end only, so
    // that run(g+1) — whose host-side table building is not short — can start while the network of group g is still running
    std::vector<cudaEvent_t> ev;
    int64_t ev_done = 0;               // groups whose times have been added to enc_ms / net_ms
    float enc_ms = 0.f, net_ms = 0.f;
    float enc_phase_ms[5] = {0, 0, 0, 0, 0};        // prefix, count, sites, alleles, windows — summed over the groups
    int64_t enc_launches = 0, net_launches = 0, groups = 0;
};
}  // namespace pb

static int session(pb_variant_encoder_t *e, VariantStream **S, bool need_active) {
    if (!e) { set_error("null handle"); return PB_ERR_ARG; }
    if (!e->vstream) e->vstream = new VariantStream();
    *S = e->vstream;
    if (need_active && !(*S)->active) { set_error("no active stream session: call pb_variant_stream_begin first"); return PB_ERR_STATE; }
    return PB_OK;
}

extern "C" int pb_variant_stream_begin(pb_variant_encoder_t *e, pb_variant_net_t *net, const pb_variant_params_t *params,
                                       int64_t capacity, pb_pred_record_t *d_records, void *stream_) {
    if (!e || !net || !params) { set_error("null argument"); return PB_ERR_ARG; }
    VariantStream *S;
    PB_TRY(session(e, &S, false));
    PB_CUDA(cudaSetDevice(e->device));
    if (!e->copy_stream) {
        PB_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) PB_CUDA(cudaEventCreateWithFlags(&e->copied[b], cudaEventDisableTiming));
    }
    PB_TRY(ensure_events(e->pevt, 3));
    const int64_t cap = std::max<int64_t>(capacity, 1);
    PB_TRY(e->p_images.reserve((size_t) cap * 33 * 26));
    PB_TRY(e->p_positions.reserve(sizeof(int64_t) * cap));
    PB_TRY(e->p_depths.reserve(cap));
    PB_TRY(e->p_freqs.reserve(cap));
    PB_TRY(e->p_keys.reserve((size_t) cap * PB_ALLELE_STRIDE));
    PB_TRY(e->p_region_of.reserve(sizeof(int32_t) * cap));
    PB_TRY(e->p_probs.reserve(sizeof(float) * 3 * cap));
    S->active = true; S->params = *params; S->capacity = capacity; S->done = 0; S->net_done = 0; S->d_records = d_records;
    S->st = (cudaStream_t) stream_;
    S->staged[0] = S->staged[1] = false; S->next_stage = 0; S->next_run = 0; S->ev_done = 0;
    S->enc_ms = S->net_ms = 0.f;
    for (int i = 0; i < 5; i++) S->enc_phase_ms[i] = 0.f;
    S->enc_launches = S->net_launches = S->groups = 0;
    return PB_OK;
}

// stage regions [g0, g1) of a host-resident workload into the free staging buffer (asynchronous H2D on the copy stream)
extern "C" int pb_variant_stream_stage_host(pb_variant_encoder_t *e, const pb_reads_t *h_reads, const pb_region_t *h_regions,
                                            int64_t g0, int64_t g1, const char *h_ref, int32_t region_id0) {
    VariantStream *S;
    PB_TRY(session(e, &S, true));
    if (!h_reads || !h_regions || g0 < 0 || g1 <= g0) { set_error("bad group [%lld, %lld)", (long long) g0, (long long) g1); return PB_ERR_ARG; }
    const int b = S->next_stage;
    if (S->staged[b]) { set_error("both staging buffers are in use: run a staged group first"); return PB_ERR_STATE; }
    for (int64_t r = g0; r < g1; r++) {
        if (h_regions[r].read_begin < 0 || h_regions[r].read_end > h_reads->n_reads || h_regions[r].read_begin > h_regions[r].read_end ||
            (r > g0 && h_regions[r].read_begin < h_regions[r - 1].read_end)) {
            set_error("region %lld: read range out of bounds or not ascending", (long long) r);
            return PB_ERR_ARG;
        }
    }
    PB_TRY(stage_group(e, b, h_reads, h_regions, g0, g1, h_ref, S->V[b]));
    S->staged[b] = true; S->from_host[b] = true; S->region_id0[b] = region_id0; S->next_stage = b ^ 1;
    return PB_OK;
}

// the same for a workload that is already resident in HBM: no copies of the reads, only the re-based region rows
extern "C" int pb_variant_stream_stage_device(pb_variant_encoder_t *e, const pb_reads_t *d_reads, const pb_region_t *h_regions,
                                              int64_t g0, int64_t g1, const char *d_ref, int32_t region_id0) {
    VariantStream *S;
    PB_TRY(session(e, &S, true));
    if (!d_reads || !h_regions || g0 < 0 || g1 <= g0) { set_error("bad group [%lld, %lld)", (long long) g0, (long long) g1); return PB_ERR_ARG; }
    const int b = S->next_stage;
    if (S->staged[b]) { set_error("both staging buffers are in use: run a staged group first"); return PB_ERR_STATE; }
    GroupView &V = S->V[b];
    const int64_t r0 = h_regions[g0].read_begin, r1 = h_regions[g1 - 1].read_end;
    if (r0 < 0 || r1 > d_reads->n_reads || r1 < r0) { set_error("group read range out of bounds"); return PB_ERR_ARG; }
    V.h_regions.assign(h_regions + g0, h_regions + g1);
    for (auto &rg : V.h_regions) { rg.read_begin -= r0; rg.read_end -= r0; }
    PB_TRY(upload(e->g_buf[b][8], V.h_regions.data(), sizeof(pb_region_t) * V.h_regions.size(), e->copy_stream));
    PB_CUDA(cudaEventRecord(e->copied[b], e->copy_stream));
    V.d = *d_reads;
    V.d.n_reads = r1 - r0;
    V.d.pos += r0; V.d.seq_off += r0; V.d.cigar_off += r0; V.d.flags += r0; V.d.mapq += r0;     // seq / qual / cigar keep absolute offsets
    V.d_regions = e->g_buf[b][8].as<pb_region_t>();
    V.d_ref = d_ref;
    S->staged[b] = true; S->from_host[b] = false; S->region_id0[b] = region_id0; S->next_stage = b ^ 1;
    return PB_OK;
}

// adds the times of every group whose events have completed (call after a stream synchronisation)
static void stream_collect_timing(VariantStream *S) {
    for (; S->ev_done < S->groups; S->ev_done++) {
        float a = 0.f, b = 0.f;
        cudaEvent_t *E = &S->ev[(size_t) 3 * S->ev_done];
        cudaEventElapsedTime(&a, E[0], E[1]);
        cudaEventElapsedTime(&b, E[1], E[2]);
        S->enc_ms += a; S->net_ms += b;
    }
}
static int stream_events(VariantStream *S, int64_t group, cudaEvent_t **E) {
    while ((int64_t) S->ev.size() < 3 * (group + 1)) {
        cudaEvent_t x;
        PB_CUDA(cudaEventCreate(&x));
        S->ev.push_back(x);
    }
    *E = &S->ev[(size_t) 3 * group];
    return PB_OK;
}

static int stream_network(pb_variant_encoder_t *e, pb_variant_net_t *net, VariantStream *S, int64_t run) {
    if (run <= 0) return PB_OK;
    pb_candidate_columns_t cols{e->p_positions.as<int64_t>() + S->net_done, e->p_region_of.as<int32_t>() + S->net_done,
                                e->p_depths.as<uint8_t>() + S->net_done, e->p_freqs.as<uint8_t>() + S->net_done,
                                e->p_keys.as<char>() + S->net_done * PB_ALLELE_STRIDE};
    PB_TRY(pb_variant_net_forward_records_device(net, e->p_images.as<int8_t>() + S->net_done * 33 * 26, run, e->p_probs.as<float>() + S->net_done * 3,
                                                 &cols, S->d_records ? S->d_records + S->net_done : nullptr, (void *) S->st));
    S->net_done += run;
    int64_t nl = 0;
    pb_variant_net_launches(net, &nl);
    S->net_launches += nl;
    return PB_OK;
}

// Encode the oldest staged group behind the accumulated candidates and queue the network over the whole chunks available
// (`flush` != 0: over everything, partial tail included).  Returns with the network still running: stage the next group, then
// call pb_variant_stream_sync.  On PB_ERR_CAPACITY *n_total holds the candidates needed so far (this group included).
extern "C" int pb_variant_stream_run(pb_variant_encoder_t *e, pb_variant_net_t *net, int flush, int64_t *n_total) {
    VariantStream *S;
    PB_TRY(session(e, &S, true));
    if (!net) { set_error("null handle"); return PB_ERR_ARG; }
    const int b = S->next_run;
    if (!S->staged[b]) { set_error("no staged group"); return PB_ERR_STATE; }
    cudaStream_t st = S->st;
    cudaEvent_t *E;
    PB_TRY(stream_events(S, S->groups, &E));
    PB_CUDA(cudaStreamWaitEvent(st, e->copied[b], 0));
    GroupView &V = S->V[b];
    int64_t n_g = 0;
    const int64_t done = S->done, room = std::max<int64_t>(S->capacity - done, 0);
    PB_CUDA(cudaEventRecord(E[0], st));
    int rc = pb_variant_encode_device(e, &V.d, V.d_regions, (int64_t) V.h_regions.size(), V.h_regions.data(), V.d_ref, 0, &S->params, room,
                                      e->p_images.as<int8_t>() + done * 33 * 26, e->p_positions.as<int64_t>() + done,
                                      e->p_depths.as<uint8_t>() + done, e->p_freqs.as<uint8_t>() + done,
                                      e->p_keys.as<char>() + done * PB_ALLELE_STRIDE, e->p_region_of.as<int32_t>() + done, nullptr, &n_g, (void *) st);
    S->staged[b] = false; S->next_run = b ^ 1;
    if (n_total) *n_total = done + n_g;
    if (rc != PB_OK) { if (rc == PB_ERR_CAPACITY) cudaStreamSynchronize(e->copy_stream); return rc; }
    for (int i = 0; i < 5; i++) S->enc_phase_ms[i] += e->ms[i];
    S->enc_launches += e->launches;
    if (n_g > 0 && S->region_id0[b] != 0)
        k_add_offset_i32<<<(unsigned) ceil_div(n_g, 256), 256, 0, st>>>(e->p_region_of.as<int32_t>() + done, n_g, S->region_id0[b]);
    PB_CUDA(cudaEventRecord(E[1], st));
    S->done = done + n_g;
    const int64_t avail = S->done - S->net_done;
    PB_TRY(stream_network(e, net, S, flush ? avail : avail / PIPE_NET_CHUNK * PIPE_NET_CHUNK));
    PB_CUDA(cudaEventRecord(E[2], st));
    S->groups++;
    return PB_OK;
}

extern "C" int pb_variant_stream_sync(pb_variant_encoder_t *e) {
    VariantStream *S;
    PB_TRY(session(e, &S, true));
    PB_CUDA(cudaStreamSynchronize(S->st));
    stream_collect_timing(S);
    return PB_OK;
}

// network over whatever is still waiting, wait for everything; *n_out = candidates of the session
extern "C" int pb_variant_stream_end(pb_variant_encoder_t *e, pb_variant_net_t *net, int64_t *n_out) {
    VariantStream *S;
    PB_TRY(session(e, &S, true));
    if (!net || !n_out) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = S->st;
    PB_CUDA(cudaStreamSynchronize(st));
    stream_collect_timing(S);
    if (S->done > S->net_done) {
        PB_CUDA(cudaEventRecord(e->pevt[1], st));
        PB_TRY(stream_network(e, net, S, S->done - S->net_done));
        PB_CUDA(cudaEventRecord(e->pevt[2], st));
        PB_CUDA(cudaStreamSynchronize(st));
        float b = 0.f;
        cudaEventElapsedTime(&b, e->pevt[1], e->pevt[2]);
        S->net_ms += b;
    }
    PB_CUDA(cudaStreamSynchronize(e->copy_stream));
    e->pms[0] = S->enc_ms; e->pms[1] = S->net_ms;
    *n_out = S->done;
    S->active = false;
    return PB_OK;
}

// per-phase encoder device time (ms: prefix, count, sites, alleles, windows) and kernel launches (encoder, network) of the
// last session, summed over its groups
extern "C" int pb_variant_stream_stats(pb_variant_encoder_t *e, float *ms5, int64_t *launches2, int64_t *groups) {
    if (!e || !e->vstream) { set_error("no session"); return PB_ERR_STATE; }
    VariantStream *S = e->vstream;
    if (ms5) for (int i = 0; i < 5; i++) ms5[i] = S->enc_phase_ms[i];
    if (launches2) { launches2[0] = S->enc_launches; launches2[1] = S->net_launches; }
    if (groups) *groups = S->groups;
    return PB_OK;
}

// device pointers of the session's accumulated columns and probabilities (valid until the next session begins)
extern "C" int pb_variant_stream_columns(pb_variant_encoder_t *e, pb_candidate_columns_t *cols, const float **d_probs, const int8_t **d_images) {
    if (!e || !cols) { set_error("null argument"); return PB_ERR_ARG; }
    cols->positions = e->p_positions.as<int64_t>(); cols->region_of = e->p_region_of.as<int32_t>();
    cols->depths = e->p_depths.as<uint8_t>(); cols->freqs = e->p_freqs.as<uint8_t>(); cols->keys = e->p_keys.as<char>();
    if (d_probs) *d_probs = e->p_probs.as<float>();
    if (d_images) *d_images = e->p_images.as<int8_t>();
    return PB_OK;
}

// D2H of the first n candidates of the last session (any pointer may be NULL)
extern "C" int pb_variant_stream_fetch(pb_variant_encoder_t *e, int64_t n, int8_t *h_images, int64_t *h_positions, uint8_t *h_depths,
                                       uint8_t *h_freqs, char *h_keys, int32_t *h_region_of, float *h_probs, void *stream_) {
    if (!e) { set_error("null handle"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    if (n > 0) {
        if (h_images) PB_CUDA(cudaMemcpyAsync(h_images, e->p_images.p, (size_t) n * 33 * 26, cudaMemcpyDeviceToHost, st));
        if (h_positions) PB_CUDA(cudaMemcpyAsync(h_positions, e->p_positions.p, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
        if (h_depths) PB_CUDA(cudaMemcpyAsync(h_depths, e->p_depths.p, n, cudaMemcpyDeviceToHost, st));
        if (h_freqs) PB_CUDA(cudaMemcpyAsync(h_freqs, e->p_freqs.p, n, cudaMemcpyDeviceToHost, st));
        if (h_keys) PB_CUDA(cudaMemcpyAsync(h_keys, e->p_keys.p, (size_t) n * PB_ALLELE_STRIDE, cudaMemcpyDeviceToHost, st));
        if (h_region_of) PB_CUDA(cudaMemcpyAsync(h_region_of, e->p_region_of.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        if (h_probs) PB_CUDA(cudaMemcpyAsync(h_probs, e->p_probs.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    }
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

// Host entry: regions are processed in groups; the H2D copy of group g+1 (copy stream, pinned source) overlaps the
// encoder + network kernels of group g, so PCIe time hides behind compute for all but the first group.
extern "C" int pb_variant_call_host(pb_variant_encoder_t *e, pb_variant_net_t *net, const pb_reads_t *h_reads,
                                    const pb_region_t *h_regions, int64_t n_regions, const char *h_ref, int64_t ref_bytes,
                                    const pb_variant_params_t *params, int64_t capacity, int8_t *h_images,
                                    int64_t *h_positions, uint8_t *h_depths, uint8_t *h_freqs, char *h_keys,
                                    int32_t *h_region_of, float *h_probs, int64_t *n_out, void *stream_) {
    if (!e || !net || !h_reads || !h_regions || !params || !n_out) { set_error("null argument"); return PB_ERR_ARG; }
    PB_CUDA(cudaSetDevice(e->device));
    *n_out = 0;
    if (n_regions <= 0) return PB_OK;
    bool ascending = true;
    for (int64_t r = 0; r < n_regions; r++) {
        if (h_regions[r].read_begin < 0 || h_regions[r].read_end > h_reads->n_reads || h_regions[r].read_begin > h_regions[r].read_end) {
            set_error("region %lld: read range out of bounds", (long long) r);
            return PB_ERR_ARG;
        }
        if (r > 0 && h_regions[r].read_begin < h_regions[r - 1].read_end) ascending = false;
    }
    if (n_regions <= CALL_GROUP || !ascending)
        return variant_call_host_single(e, net, h_reads, h_regions, n_regions, h_ref, ref_bytes, params, capacity, h_images, h_positions,
                                        h_depths, h_freqs, h_keys, h_region_of, h_probs, n_out, stream_);
    // group boundaries: a short first group keeps the un-overlapped first copy small
    std::vector<int64_t> gb;
    gb.push_back(0);
    for (int64_t r = std::min<int64_t>(CALL_FIRST_GROUP, n_regions); r < n_regions; r += CALL_GROUP) gb.push_back(r);
    gb.push_back(n_regions);
    const int64_t n_groups = (int64_t) gb.size() - 1;
    PB_TRY(pb_variant_stream_begin(e, net, params, capacity, nullptr, stream_));
    PB_TRY(pb_variant_stream_stage_host(e, h_reads, h_regions, gb[0], gb[1], h_ref, 0));
    for (int64_t g = 0; g < n_groups; g++) {
        int64_t n_tot = 0;
        const int rc = pb_variant_stream_run(e, net, g + 1 == n_groups, &n_tot);
        if (rc == PB_ERR_CAPACITY) {
            // the caller retries with the returned size: extrapolate from the regions seen so far (retried again if short)
            const int64_t need = (int64_t) ((double) n_tot * (double) n_regions / (double) gb[g + 1] * 1.25) + 4096;
            e->vstream->active = false;
            *n_out = need;
            set_error("candidate capacity %lld too small (estimated need %lld)", (long long) capacity, (long long) need);
            return PB_ERR_CAPACITY;
        }
        if (rc != PB_OK) { e->vstream->active = false; return rc; }
        // the network of this group is queued: issue the next group's copies now, so that neither the host work of staging nor
        // the copies themselves leave the compute stream idle
        // no sync here: the next run() builds its host-side tables while this group's network runs, and its encoder synchronises the
        // stream before any staging buffer is reused
        if (g + 1 < n_groups) PB_TRY(pb_variant_stream_stage_host(e, h_reads, h_regions, gb[g + 1], gb[g + 2], h_ref, (int32_t) gb[g + 1]));
    }
    PB_TRY(pb_variant_stream_end(e, net, n_out));
    return pb_variant_stream_fetch(e, *n_out, h_images, h_positions, h_depths, h_freqs, h_keys, h_region_of, h_probs, stream_);
}

extern "C" int pb_variant_call_timings(pb_variant_encoder_t *e, float *ms2) {
    if (!e || !ms2) return PB_ERR_ARG;
    ms2[0] = e->pms[0]; ms2[1] = e->pms[1];
    return PB_OK;
}

extern "C" int pb_polish_call_device(pb_polish_encoder_t *e, pb_polish_net_t *net, const pb_reads_t *d_reads,
                                     const pb_region_t *d_regions, int64_t n_regions, const pb_region_t *h_regions,
                                     int64_t capacity_images, uint8_t *d_bases, uint8_t *d_phred, int64_t *d_position,
                                     int32_t *d_index, int32_t *d_image_region, int32_t *d_chunk_id, int64_t *n_images_out,
                                     void *stream_) {
    if (!e || !net || !n_images_out) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_TRY(ensure_events(e->pevt, 3));
    PB_CUDA(cudaEventRecord(e->pevt[0], st));
    *n_images_out = 0;
    // encode into library scratch; column capacity: positions + generous insert allowance, retried on demand
    int64_t span = 0;
    for (int64_t r = 0; r < n_regions; r++) span += h_regions[r].ref_end - h_regions[r].ref_start + 1;
    int64_t cap_cols = 2 * span + 1024, n_cols = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        PB_TRY(e->p_image.reserve((size_t) cap_cols * 10));
        PB_TRY(e->p_pos.reserve(sizeof(int64_t) * cap_cols));
        PB_TRY(e->p_idx.reserve(sizeof(int32_t) * cap_cols));
        PB_TRY(e->p_col_off.reserve(sizeof(int64_t) * (n_regions + 1)));
        int rc = pb_polish_encode_device(e, d_reads, d_regions, n_regions, h_regions, cap_cols, e->p_image.as<uint8_t>(),
                                         e->p_pos.as<int64_t>(), e->p_idx.as<int32_t>(), e->p_col_off.as<int64_t>(), &n_cols, stream_);
        if (rc == PB_ERR_CAPACITY && attempt == 0) { cap_cols = n_cols + 16; continue; }
        if (rc != PB_OK) return rc;
        break;
    }
    std::vector<int64_t> col_off((size_t) n_regions + 1, 0);
    if (n_regions > 0) PB_CUDA(cudaMemcpyAsync(col_off.data(), e->p_col_off.p, sizeof(int64_t) * (n_regions + 1), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    std::vector<ChunkRow> rows;
    for (int64_t r = 0; r < n_regions; r++) {
        const int64_t n = col_off[r + 1] - col_off[r];
        int64_t start = 0, end = std::min<int64_t>(n, PB_POLISH_SEQ_LEN);
        int cid = 0;
        while (true) {
            ChunkRow c; c.col_start = col_off[r] + start; c.nvalid = (int32_t) (end - start); c.region = (int32_t) r; c.chunk_id = cid++; c.pad = 0;
            rows.push_back(c);
            if (end == n) break;
            start = end - 50;                                                  // SEQ_OVERLAP
            end = std::min<int64_t>(n, start + PB_POLISH_SEQ_LEN);
        }
    }
    const int64_t n_img = (int64_t) rows.size();
    *n_images_out = n_img;
    if (n_img > capacity_images) {
        set_error("image capacity %lld < %lld needed", (long long) capacity_images, (long long) n_img);
        return PB_ERR_CAPACITY;
    }
    if (n_img > 0) {
        PB_TRY(upload(e->p_chunks, rows.data(), sizeof(ChunkRow) * n_img, st));
        PB_TRY(e->p_imgs.reserve((size_t) n_img * PB_POLISH_SEQ_LEN * PB_POLISH_FEATURES));
        k_polish_chunk<<<(unsigned) n_img, 256, 0, st>>>(e->p_chunks.as<ChunkRow>(), e->p_image.as<uint8_t>(), e->p_pos.as<int64_t>(),
                                                        e->p_idx.as<int32_t>(), e->p_imgs.as<uint8_t>(), d_position, d_index,
                                                        d_image_region, d_chunk_id);
        PB_CUDA(cudaGetLastError());
    }
    PB_CUDA(cudaEventRecord(e->pevt[1], st));
    if (n_img > 0) PB_TRY(pb_polish_net_forward_device(net, e->p_imgs.as<uint8_t>(), n_img, d_bases, d_phred, nullptr, nullptr, stream_));
    PB_CUDA(cudaEventRecord(e->pevt[2], st));
    PB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&e->pms[0], e->pevt[0], e->pevt[1]);
    cudaEventElapsedTime(&e->pms[1], e->pevt[1], e->pevt[2]);
    return PB_OK;
}

extern "C" int pb_polish_call_host(pb_polish_encoder_t *e, pb_polish_net_t *net, const pb_reads_t *h_reads,
                                   const pb_region_t *h_regions, int64_t n_regions, int64_t capacity_images, uint8_t *h_bases,
                                   uint8_t *h_phred, int64_t *h_position, int32_t *h_index, int32_t *h_image_region,
                                   int32_t *h_chunk_id, int64_t *n_images_out, void *stream_) {
    if (!e || !net || !h_reads || !h_regions || !n_images_out) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(e->device));
    DevBuf *rb[8] = {&e->h_pos, &e->h_seq_off, &e->h_cigar_off, &e->h_flags, &e->h_mapq, &e->h_seq, &e->h_qual, &e->h_cigar};
    pb_reads_t d;
    PB_TRY(upload_reads(h_reads, rb, &d, st));
    PB_TRY(upload(e->h_regions, h_regions, sizeof(pb_region_t) * n_regions, st));
    const int64_t cap = std::max<int64_t>(capacity_images, 1);
    PB_TRY(e->p_bases.reserve((size_t) cap * PB_POLISH_SEQ_LEN));
    PB_TRY(e->p_phred.reserve((size_t) cap * PB_POLISH_SEQ_LEN));
    PB_TRY(e->p_position.reserve(sizeof(int64_t) * cap * PB_POLISH_SEQ_LEN));
    PB_TRY(e->p_index.reserve(sizeof(int32_t) * cap * PB_POLISH_SEQ_LEN));
    PB_TRY(e->p_iregion.reserve(sizeof(int32_t) * cap));
    PB_TRY(e->p_cid.reserve(sizeof(int32_t) * cap));
    PB_TRY(pb_polish_call_device(e, net, &d, e->h_regions.as<pb_region_t>(), n_regions, h_regions, capacity_images,
                                 e->p_bases.as<uint8_t>(), e->p_phred.as<uint8_t>(), e->p_position.as<int64_t>(),
                                 e->p_index.as<int32_t>(), e->p_iregion.as<int32_t>(), e->p_cid.as<int32_t>(), n_images_out, stream_));
    const int64_t n = *n_images_out;
    if (n > 0) {
        PB_CUDA(cudaMemcpyAsync(h_bases, e->p_bases.p, (size_t) n * PB_POLISH_SEQ_LEN, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_phred, e->p_phred.p, (size_t) n * PB_POLISH_SEQ_LEN, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_position, e->p_position.p, sizeof(int64_t) * n * PB_POLISH_SEQ_LEN, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_index, e->p_index.p, sizeof(int32_t) * n * PB_POLISH_SEQ_LEN, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_image_region, e->p_iregion.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaMemcpyAsync(h_chunk_id, e->p_cid.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    }
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

extern "C" int pb_polish_call_timings(pb_polish_encoder_t *e, float *ms2) {
    if (!e || !ms2) return PB_ERR_ARG;
    ms2[0] = e->pms[0]; ms2[1] = e->pms[1];
    return PB_OK;
}
