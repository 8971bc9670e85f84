// Candidate selection on the GPU: one thread per prediction record; predictions, keys and the
// reference stay in HBM between the network and the VCF writer.
#include "common.cuh"
#include <algorithm>

namespace pb {

__global__ void k_find_candidates(const int64_t *__restrict__ positions, const int32_t *__restrict__ region_of,
                                  const uint8_t *__restrict__ depths, const uint8_t *__restrict__ freqs, const char *__restrict__ keys,
                                  const float *__restrict__ probs, int64_t n, const pb_region_t *__restrict__ regions,
                                  const char *__restrict__ ref, pb_candidate_options_t O, uint8_t *__restrict__ flags,
                                  uint8_t *__restrict__ genotype) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const pb_region_t rg = regions[region_of[i]];
    const int64_t x = positions[i] - rg.ref_start;
    // reference window [pos-10, pos+10) clipped to the available reference (CandidateFinder.py:389-393)
    char w[20];
    int nw = 0, pi = 0;
    for (int64_t k = x - 10; k < x + 10; k++) {
        if (k == x) pi = nw;
        if (k >= 0 && k < rg.ref_len) w[nw++] = to_upper(ref[rg.ref_off + k]);
    }
    if (x >= rg.ref_len) pi = nw;
    // homopolymer run length containing each position, maximum over [pi-5, pi+4)  (:395-406)
    int maxrun = 0;
    const int lo = max(0, pi - 5), hi = min(nw, pi + 4);
    for (int k = 0; k < nw;) {
        int e = k + 1;
        while (e < nw && w[e] == w[k]) e++;
        if (e > lo && k < hi) maxrun = max(maxrun, e - k);      // the run [k, e) intersects the look-up window
        k = e;
    }
    const bool in_repeat = maxrun >= 5;
    const char ref_base = (x >= 0 && x < rg.ref_len) ? w[pi] : '\0';
    const bool ref_ok = is_upper_acgt(ref_base);                                   // :408
    const float p0 = probs[i * 3], p1 = probs[i * 3 + 1], p2 = probs[i * 3 + 2];
    int g = 0; float pv = p0;                                                      // np.argmax: first maximum (:411)
    if (p1 > pv) { g = 1; pv = p1; }
    if (p2 > pv) { g = 2; pv = p2; }
    const char *key = keys + i * PB_ALLELE_STRIDE;
    const char t = key[0];
    bool valid = true;
    for (int k = 1; k < PB_ALLELE_STRIDE && key[k]; k++) valid = valid && is_upper_acgt(key[k]);
    uint8_t f = 0;
    if (in_repeat) f |= 4;
    if (ref_ok) f |= 16;
    if (ref_ok && valid) {
        if (t == '1' && g != 0) f |= 1;                                            // Margin: SNPs with a non-ref genotype
        const double na = (double) fmaxf(p1, p2);
        // depth 0 cannot come out of the encoder (a candidate needs support >= 2); the reference would raise ZeroDivisionError
        // (:456) — here such a record is simply never selected by the frequency rule.  freqs / depths are the encoder's uint8
        // values, saturated at 125 like the reference's stores (region_summary.cpp:857, DataStore.py:66).
        const double vaf = depths[i] ? (double) freqs[i] / (double) depths[i] : -1.0;
        const double pthr = t == '1' ? (in_repeat ? O.snp_p_value_in_lc : O.snp_p_value)
                          : t == '2' ? (in_repeat ? O.insert_p_value_in_lc : O.insert_p_value)
                                     : (in_repeat ? O.delete_p_value_in_lc : O.delete_p_value);
        const double fthr = t == '1' ? O.report_snp_above_freq : O.report_indel_above_freq;
        if (na >= pthr) { f |= 2; if (t == '3') f |= 8; }
        else if (0.0 < fthr && fthr <= vaf) f |= 2;
    }
    flags[i] = f;
    genotype[i] = (uint8_t) g;
}

}  // namespace pb

using namespace pb;

extern "C" int pb_variant_find_candidates_device(const int64_t *d_positions, const int32_t *d_region_of, const uint8_t *d_depths,
                                                 const uint8_t *d_freqs, const char *d_keys, const float *d_probs, int64_t n,
                                                 const pb_region_t *d_regions, const char *d_ref, const pb_candidate_options_t *opt,
                                                 uint8_t *d_flags, uint8_t *d_genotype, void *stream_) {
    if (!opt) { set_error("null options"); return PB_ERR_ARG; }
    if (n <= 0) return PB_OK;
    k_find_candidates<<<(unsigned) ceil_div(n, 256), 256, 0, (cudaStream_t) stream_>>>(d_positions, d_region_of, d_depths, d_freqs, d_keys, d_probs, n,
                                                                                     d_regions, d_ref, *opt, d_flags, d_genotype);
    PB_CUDA(cudaGetLastError());
    return PB_OK;
}

extern "C" int pb_variant_find_candidates_host(const int64_t *h_positions, const int32_t *h_region_of, const uint8_t *h_depths,
                                               const uint8_t *h_freqs, const char *h_keys, const float *h_probs, int64_t n,
                                               const pb_region_t *h_regions, int64_t n_regions, const char *h_ref, int64_t ref_bytes,
                                               const pb_candidate_options_t *opt, uint8_t *h_flags, uint8_t *h_genotype, void *stream_) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { set_error("no CUDA device: libpepper_b200 has no CPU fallback"); return PB_ERR_CUDA; }
    if (n <= 0) return PB_OK;
    cudaStream_t st = (cudaStream_t) stream_;
    DevBuf p, r, d, f, k, pr, rg, rf, of, og;
    PB_TRY(upload(p, h_positions, sizeof(int64_t) * n, st));
    PB_TRY(upload(r, h_region_of, sizeof(int32_t) * n, st));
    PB_TRY(upload(d, h_depths, n, st));
    PB_TRY(upload(f, h_freqs, n, st));
    PB_TRY(upload(k, h_keys, (size_t) n * PB_ALLELE_STRIDE, st));
    PB_TRY(upload(pr, h_probs, sizeof(float) * 3 * n, st));
    PB_TRY(upload(rg, h_regions, sizeof(pb_region_t) * n_regions, st));
    PB_TRY(upload(rf, h_ref, (size_t) ref_bytes, st));
    PB_TRY(of.reserve(n)); PB_TRY(og.reserve(n));
    int rc = pb_variant_find_candidates_device(p.as<int64_t>(), r.as<int32_t>(), d.as<uint8_t>(), f.as<uint8_t>(), k.as<char>(), pr.as<float>(), n,
                                               rg.as<pb_region_t>(), rf.as<char>(), opt, of.as<uint8_t>(), og.as<uint8_t>(), stream_);
    if (rc == PB_OK) {
        cudaError_t e = cudaMemcpyAsync(h_flags, of.p, n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_genotype, og.p, n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("find_candidates: %s", cudaGetErrorString(e)); rc = PB_ERR_CUDA; }
    }
    DevBuf *bufs[] = {&p, &r, &d, &f, &k, &pr, &rg, &rf, &of, &og};
    for (auto *b : bufs) b->release();
    return rc;
}
