// Shared host/device helpers for libpepper_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include "../../include/pepper_b200.h"

namespace pb {

void set_error(const char *fmt, ...);

#define PB_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            pb::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,        \
                          cudaGetErrorString(e_));                                     \
            return PB_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

#define PB_TRY(call)                  \
    do {                              \
        int rc_ = (call);             \
        if (rc_ != PB_OK) return rc_; \
    } while (0)

// Grow-only device buffer owned by a handle.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return PB_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return PB_ERR_CUDA;
        }
        cap = want;
        return PB_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

int upload(DevBuf &b, const void *h, size_t bytes, cudaStream_t st);
int upload_reads(const pb_reads_t *h, DevBuf *const bufs[8], pb_reads_t *d, cudaStream_t st);

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// BAM 4-bit code -> ASCII (seq_nt16_str), the characters the reference sees after get_reads
// (bam_handler.cpp:213 toupper(seq_nt16_str[bam_seqi(..)])).
__device__ __forceinline__ char nt16_char(int code) { return "=ACMGRSVTWYHKDBN"[code & 15]; }

// base (nibble) i of a packed 4-bit sequence
__device__ __forceinline__ int seq_code_at(const uint8_t *__restrict__ seq, int64_t nib) {
    const uint8_t b = __ldg(seq + (nib >> 1));
    return (nib & 1) ? (b & 15) : (b >> 4);
}

// rank of NT16 characters in ASCII order ('=' < 'A' < 'B' < 'C' < 'D' < 'G' < 'H' < 'K' < 'M' < 'N' <
// 'R' < 'S' < 'T' < 'V' < 'W' < 'Y'): std::set<string> ordering of allele keys, region_summary.cpp:670
__device__ __forceinline__ int nt16_ascii_rank(int code) {
    // ranks by code 0..15: 0,1,3,8,5,10,11,13,12,14,15,6,7,4,2,9  (4 bits each)
    return (int) ((0x92476FECDBA58310ULL >> (4 * (code & 15))) & 15ULL);
}
__device__ __forceinline__ int nt16_code_of_rank(int rank) {
    // codes by rank 0..15: 0,1,14,2,13,4,11,12,3,15,5,6,8,7,9,10
    return (int) ((0xA97865F3CB4D2E10ULL >> (4 * (rank & 15))) & 15ULL);
}
__device__ __forceinline__ bool is_upper_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
__device__ __forceinline__ char to_upper(char c) { return (c >= 'a' && c <= 'z') ? (char) (c - 32) : c; }
// ------------------------------------------------------------------ single-CTA exclusive scan (n <= a few 1e6)
// out[i] = sum(in[0..i)), out[n] = total (also *total).  Each thread scans 8 consecutive elements per round, so one round of
// the 1024-thread block covers 8192 inputs with two block barriers.
static __global__ void k_scan_excl(const int32_t *__restrict__ in, int64_t *__restrict__ out, int64_t n, int64_t *__restrict__ total) {
    constexpr int PER = 8;
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry, s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = (int) (blockDim.x >> 5);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < n; b += (int64_t) blockDim.x * PER) {
        const int64_t i0 = b + (int64_t) tid * PER;
        int32_t v[PER];
        int64_t mine = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { v[k] = (i0 + k < n) ? in[i0 + k] : 0; mine += v[k]; }
        int64_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = (lane < nwarps) ? s_warp[lane] : 0;
            int64_t winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t u = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += u;
            }
            s_warp[lane] = winc - w;
            if (lane == 31) s_total = winc;
        }
        __syncthreads();
        int64_t run = s_carry + s_warp[warp] + inc - mine;
#pragma unroll
        for (int k = 0; k < PER; k++) { if (i0 + k < n) out[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (tid == 0) s_carry += s_total;
        __syncthreads();
    }
    if (tid == 0) { out[n] = s_carry; if (total) *total = s_carry; }
}

#endif

}  // namespace pb
