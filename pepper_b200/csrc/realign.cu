// Read -> reference realignment on the device.
//
// Replaces ReadAligner(ref_start, ref_end, ref_seq).align_reads_to_reference(reads)
// (pepper/modules/src/local_reassembly/simple_aligner.cpp:66-106, pybind_api.h ReadAligner) for every region of a batch:
// each read is aligned with the reference's SSW configuration (match 4, mismatch 6, gap open 8, gap extend 2,
// simple_aligner.h:19-25) against the region reference from the read's own start, and gets the SSW CIGAR / position when
// sw_score > 1.  Results are bit-identical to the SSE2 library the reference vendors (ssw.c / ssw_cpp.cpp), including:
//   * E opens only from the H of the striped main loop, i.e. from an F chain restricted to one stripe segment
//     (segLen = ceil(readLen / 16) in byte mode, / 8 in word mode; ssw.c:282,487 "disallow adjacent insertion and then
//     deletion"), while H itself gets the full Lazy-F correction;
//   * byte mode first, word mode when max + bias >= 255 (ssw.c:330, 826-830);
//   * best cell = highest score, first column in iteration order, smallest read index (ssw.c:519-531);
//   * the reverse pass over the prefixes that stops at the first column whose maximum equals the score (ssw.c:512);
//   * banded_sw's band arithmetic, band doubling, tie rules and corner-anchored trace-back (ssw.c:571-757);
//   * '=' / 'X' splitting and soft clips of ssw_cpp.cpp:52-215, mapped to M / S tuples by CigarStringToVector.
//
// Kernel 1 (k_sw): one warp per read.  The read's positions are blocked over the 32 lanes (R rows per lane in registers);
// reference columns are processed one at a time: a descending pass builds max(diag, E, 0), two max-plus warp scans give
// every lane the F value entering its block (segment-restricted and unrestricted chains), an ascending pass finishes
// H / E / F and tracks the lane-local best cell; __reduce_max_sync gives the column maximum for the overflow / terminate
// tests.  Kernel 2 (k_banded): one warp per read, rows sequential, band columns over the lanes, F by a warp scan; one
// packed direction byte per cell in a per-warp scratch slot; lane 0 traces back and writes the final tuples.
#include "common.cuh"
#include <vector>
#include <algorithm>
#include <stdlib.h>

using namespace pb;

namespace {

constexpr int GO = 8, GE = 2, MATCH = 4, MISM = 6, BIAS = 6;
constexpr int NEG = -(1 << 28);
constexpr unsigned FULL = 0xffffffffu;
constexpr int RMAX = 42;                      // rows per lane held in registers: reads up to 32 * 42 = 1344 bases
constexpr int RDYN = 0;                       // fallback for longer reads (e.g. a soft clip of tens of kb ending inside the region): rows in a
                                              // global scratch slab [row][lane], any length — the reference's SSW has no limit either

struct Task {                                 // one read against the region reference from its own start
    int64_t q_off;                            // offset of the read's codes in `codes`
    int64_t r_off;                            // offset of the reference suffix in `rcodes`
    int32_t q_len, r_len;
};
struct Aln { int32_t score, ref_begin, ref_end, read_begin, read_end, status; };   // status: 0 none, 1 needs cigar, 2 done, 3 band slot overflow, <0 error

__device__ __forceinline__ int ssw_code_of_nt16(int c) { return c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4; }
__device__ __forceinline__ int ssw_code_of_ascii(int ch) {                       // kBaseTranslation (ssw_cpp.cpp:12-30)
    switch (ch) {
        case 'A': case 'a': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

__global__ void k_codes_reads(const uint8_t *__restrict__ seq, int64_t nb, int8_t *__restrict__ codes) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) codes[i] = (int8_t) ssw_code_of_nt16(seq_code_at(seq, i));
}
__global__ void k_codes_ref(const char *__restrict__ ref, int64_t n, int8_t *__restrict__ rcodes) {
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rcodes[i] = (int8_t) ssw_code_of_ascii((unsigned char) ref[i]);
}

// thread per read: its task; flags reads that start before their region (the reference drops those)
__global__ void k_tasks(pb_reads_t R, const pb_region_t *__restrict__ regions, int64_t n_regions, Task *__restrict__ tasks,
                        int32_t *__restrict__ region_of, int32_t *__restrict__ err) {
    const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R.n_reads) return;
    int64_t lo = 0, hi = n_regions;                       // last region with read_begin <= r (regions are in read order)
    while (hi - lo > 1) { const int64_t m = (lo + hi) >> 1; if (regions[m].read_begin <= r) lo = m; else hi = m; }
    const pb_region_t g = regions[lo];
    Task t;
    t.q_off = R.seq_off[r];
    t.q_len = (int32_t) (R.seq_off[r + 1] - R.seq_off[r]);
    const int64_t start_index = R.pos[r] - g.ref_start;
    if (r < g.read_begin || r >= g.read_end) { atomicExch(err, 2); t.q_len = 0; }
    if (start_index < 0) atomicExch(err, 1);
    t.r_off = g.ref_off + max((int64_t) 0, start_index);
    t.r_len = (int32_t) max((int64_t) 0, g.ref_len - start_index);
    if (start_index < 0) t.r_len = 0;
    tasks[r] = t;
    region_of[r] = (int32_t) lo;
}

struct SwBest { int score, ref, read; };

// One striped-semantics pass.  q[p] = codes[qbase + qstep * p], p < readLen; reference column order: ref_dir 0 ascending,
// 1 descending over [0, refLen).  lanesL = 16 (byte mode) / 8 (word mode).
template <int R>
__device__ __noinline__ SwBest sw_pass(const int8_t *__restrict__ codes, int64_t qbase, int qstep, int readLen, const int8_t *__restrict__ ref,
                                       int refLen, int ref_dir, int lanesL, bool byte_mode, int terminate, int lane, int *__restrict__ gHE = nullptr) {
    constexpr bool DYN = R == RDYN;                           // rows in the global slab gHE: H at [2r][lane], E at [2r+1][lane]
    constexpr int RA = DYN ? 1 : R;
    const int Rl = (readLen + 31) >> 5;
    const int S = (readLen + lanesL - 1) / lanesL;
    const int p0 = lane * Rl;
    const int cnt = max(0, min(Rl, readLen - p0));
    constexpr int UNR = DYN ? 1 : R;
    int Hreg[RA], Ereg[RA];
    uint32_t qc[(RA + 7) / 8];
    auto Hr = [&](int r) -> int & { if constexpr (DYN) return gHE[((int64_t) 2 * r) * 32 + lane]; else return Hreg[r]; };
    auto Er = [&](int r) -> int & { if constexpr (DYN) return gHE[((int64_t) 2 * r + 1) * 32 + lane]; else return Ereg[r]; };
    auto Qr = [&](int r) -> int {
        if constexpr (DYN) return (int) (codes[qbase + (int64_t) qstep * (p0 + r)] & 15);
        else return (int) ((qc[r >> 3] >> (4 * (r & 7))) & 15);
    };
#pragma unroll UNR
    for (int k = 0; k < (RA + 7) / 8; k++) qc[k] = 0;
#pragma unroll UNR
    for (int r = 0; r < (DYN ? cnt : R); r++) {
        Hr(r) = 0; Er(r) = 0;
        if constexpr (!DYN) { if (r < cnt) qc[r >> 3] |= (uint32_t) (codes[qbase + (int64_t) qstep * (p0 + r)] & 15) << (4 * (r & 7)); }
    }
    int reset_r = (S - p0 % S) % S;                           // row of this block that starts a stripe segment
    if (reset_r >= cnt) reset_r = -1;
    const int dfull = -GE * Rl;
    int hlast = 0;                                            // H of this lane's last row (diag source of the next lane)
    int lbest = 0, lcol = 0, lrow = 0;
    bool overflow = false;
    for (int c = 0; c < refLen; c++) {
        const int i = ref_dir ? refLen - 1 - c : c;
        const int rc = ref[i];
        int up = __shfl_up_sync(FULL, hlast, 1);
        if (lane == 0) up = 0;
        // pass 1 (descending): H[r] <- max(diag + s, E, 0) using the previous column's H
#pragma unroll UNR
        for (int r = (DYN ? cnt : R) - 1; r >= 0; r--) {
            if (r < cnt) {
                const int code = Qr(r);
                const int s = (code == rc && rc < 4) ? MATCH : -MISM;
                const int diag = (r == 0 ? up : Hr(r - 1)) + s;
                Hr(r) = max(max(diag, Er(r)), 0);
            }
        }
        // block transfer of the two F chains: F_out = max(F_in + d, A)
        int aloc = 0, afull = 0;
#pragma unroll UNR
        for (int r = 0; r < (DYN ? cnt : R); r++) {
            if (r < cnt) {
                if (r == reset_r) aloc = 0;
                const int open = max(Hr(r) - GO, 0);
                aloc = max(aloc - GE, open);
                afull = max(afull - GE, open);
            }
        }
        // unrestricted chain: prefix max of (A + GE*Rl*lane); restricted chain: (d, A) pair scan with d = NEG at a reset
        int bfull = afull + GE * Rl * lane;
        int dl = (reset_r >= 0) ? NEG : dfull, al = aloc;
        if (cnt == 0) { bfull = NEG; dl = 0; al = NEG; }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int ub = __shfl_up_sync(FULL, bfull, d);
            const int ud = __shfl_up_sync(FULL, dl, d), ua = __shfl_up_sync(FULL, al, d);
            if (lane >= d) {
                bfull = max(bfull, ub);
                al = max(al, ua + dl < NEG ? NEG : ua + dl);
                dl = max(NEG, ud + dl);
            }
        }
        int ffull = __shfl_up_sync(FULL, bfull, 1) - GE * Rl * (lane - 1);
        int floc = __shfl_up_sync(FULL, al, 1);
        if (lane == 0) { ffull = 0; floc = 0; }
        ffull = max(ffull, 0); floc = max(floc, 0);
        // pass 2 (ascending): finish H / E / F, lane-local best in (column, row) order
        int cm = 0;
#pragma unroll UNR
        for (int r = 0; r < (DYN ? cnt : R); r++) {
            if (r < cnt) {
                if (r == reset_r) floc = 0;
                const int hm = max(Hr(r), floc);
                const int open = max(hm - GO, 0);
                Er(r) = max(Er(r) - GE, open);
                floc = max(floc - GE, open);
                const int hf = max(hm, ffull);
                ffull = max(ffull - GE, open);
                Hr(r) = hf;
                cm = max(cm, hf);
                if (hf > lbest) { lbest = hf; lcol = c; lrow = p0 + r; }
                if (r == cnt - 1) hlast = hf;
            }
        }
        const int colmax = __reduce_max_sync(FULL, cm);
        if (byte_mode && colmax + BIAS >= 255) { overflow = true; break; }
        if (colmax == terminate) break;
    }
    // lexicographic reduce: score desc, column order asc, row asc
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const int os = __shfl_xor_sync(FULL, lbest, d), oc = __shfl_xor_sync(FULL, lcol, d), orow = __shfl_xor_sync(FULL, lrow, d);
        const bool take = os > lbest || (os == lbest && (oc < lcol || (oc == lcol && orow < lrow)));
        if (take) { lbest = os; lcol = oc; lrow = orow; }
    }
    SwBest b;
    if (overflow) { b.score = 255; b.ref = 0; b.read = 0; return b; }
    if (lbest == 0) { b.score = 0; b.ref = byte_mode ? -1 : 0; b.read = readLen - 1; return b; }
    b.score = lbest; b.ref = ref_dir ? refLen - 1 - lcol : lcol; b.read = lrow;
    return b;
}

// ssw_align steps 1-2 (ssw.c:801-866): score / end (byte then word), begin by the reverse pass
template <int R>
__global__ void __launch_bounds__(128) k_sw(const Task *__restrict__ tasks, const int32_t *__restrict__ order, int64_t n, const int8_t *__restrict__ codes,
                                            const int8_t *__restrict__ rcodes, Aln *__restrict__ out, int *__restrict__ slab = nullptr,
                                            const int64_t *__restrict__ slab_off = nullptr) {
    const int64_t w = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int64_t a = order[w];
    const Task t = tasks[a];
    int *gHE = (R == RDYN) ? slab + slab_off[w] : nullptr;      // this warp's [2 * rows][32] slab
    Aln res = {0, -1, 0, -1, 0, 0};
    if (t.q_len > 0 && t.r_len > 0) {
        const int8_t *ref = rcodes + t.r_off;
        bool word = false;
        SwBest b = sw_pass<R>(codes, t.q_off, 1, t.q_len, ref, t.r_len, 0, 16, true, 255, lane, gHE);
        if (b.score == 255) { b = sw_pass<R>(codes, t.q_off, 1, t.q_len, ref, t.r_len, 0, 8, false, 65535, lane, gHE); word = true; }
        res.score = b.score; res.ref_end = b.ref; res.read_end = b.read;
        if (b.score > 1 && b.ref >= 0) {                       // results with score <= 1 are never used (simple_aligner.cpp:84)
            const SwBest rb = word ? sw_pass<R>(codes, t.q_off + b.read, -1, b.read + 1, ref, b.ref + 1, 1, 8, false, b.score, lane, gHE)
                                   : sw_pass<R>(codes, t.q_off + b.read, -1, b.read + 1, ref, b.ref + 1, 1, 16, true, b.score, lane, gHE);
            res.ref_begin = rb.ref; res.read_begin = b.read - rb.read;
            res.status = 1;
        }
    }
    if (lane == 0) out[a] = res;
}

// ---------------------------------------------------------------------------------------------- packed s16x2 pass
// Same recurrences, two rows per 32-bit register (VIADDMNMX.S16x2 / VIMNMX.S16x2 are native on sm_100a).
// Row mapping: stripe segment s (length S) is split over G = 32 / lanesL consecutive GPU lanes (Rl = ceil(S / G) rows each), so
// a stripe boundary is always a lane boundary and the restricted F chain differs from the unrestricted one only in the
// cross-lane scan.  Inside a lane the rows are split in a low half (first ceil(cnt/2) rows, low 16 bits) and a high half
// (the rest, high 16 bits): two independent sub-blocks that advance in the same instruction.  Pads sit at the end of each
// half (mask registers keep them out of the maxima; -30000 keeps them out of the chain transfer).
constexpr int NPMAX = RMAX / 2;                            // row pairs per lane (largest instantiation)
constexpr int SW16_SMEM_WARP = (5 * NPMAX + NPMAX) * 32 * 4;   // profile [5][NP][32] + best-column snapshot [NP][32]

__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return ((uint32_t) hi << 16) | ((uint32_t) lo & 0xffffu); }
__device__ __forceinline__ int lo16(uint32_t v) { return (int) (int16_t) (v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int) (int16_t) (v >> 16); }

template <int NP>
__device__ __noinline__ SwBest sw_pass16(const int8_t *__restrict__ codes, int64_t qbase, int qstep, int readLen, const int8_t *__restrict__ ref,
                                         int refLen, int ref_dir, int lanesL, bool byte_mode, int terminate, int lane, uint32_t *smem) {
    uint32_t *prof = smem;                                  // [rc][pair][lane]
    uint32_t *snap = smem + 5 * NP * 32;                    // [pair][lane]
    const int G = 32 / lanesL;
    const int S = (readLen + lanesL - 1) / lanesL;
    const int Rl = (S + G - 1) / G;
    const int seg = lane / G, sub = lane % G;
    const int start = seg * S + sub * Rl;
    const int endx = min(seg * S + min((sub + 1) * Rl, S), readLen);
    const int cnt = max(0, endx - start);
    const int cl = (cnt + 1) >> 1, ch = cnt - cl;
    const bool seg_start = sub == 0;
    constexpr int PADNEG = -30000;
    uint32_t H2[NP], E2[NP], M2[NP];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NP; r++) {
        H2[r] = 0; E2[r] = 0;
        M2[r] = (r < cl ? 0x0000ffffu : 0u) | (r < ch ? 0xffff0000u : 0u);
        const int ql = r < cl ? codes[qbase + (int64_t) qstep * (start + r)] : 9;
        const int qh = r < ch ? codes[qbase + (int64_t) qstep * (start + cl + r)] : 9;
#pragma unroll
        for (int rc = 0; rc < 5; rc++) {
            const int sl = ql == 9 ? PADNEG : ((ql == rc && rc < 4) ? MATCH : -MISM);
            const int sh = qh == 9 ? PADNEG : ((qh == rc && rc < 4) ? MATCH : -MISM);
            prof[(rc * NP + r) * 32 + lane] = pack2(sl, sh);
        }
    }
    __syncwarp();
    const unsigned nonempty = __ballot_sync(FULL, cnt > 0);
    const unsigned below = nonempty & ((1u << lane) - 1u);
    const int src_lane = below ? 31 - __clz(below) : -1;
    const uint32_t NGO2 = pack2(-GO, -GO), NGE2 = pack2(-GE, -GE), PAD2 = pack2(PADNEG, PADNEG);
    const int npad_l = NP - cl, npad_h = NP - ch;
    int hlast = 0, hmid = 0;                                 // H of the lane's last row / of the low half's last row (previous column)
    int gbest = 0, gcol = 0;
    bool overflow = false, have = false;
    for (int c = 0; c < refLen; c++) {
        const int i = ref_dir ? refLen - 1 - c : c;
        const int rc = ref[i];
        int up = __shfl_sync(FULL, hlast, src_lane < 0 ? 0 : src_lane);
        if (src_lane < 0) up = 0;
        const uint32_t up2 = pack2(up, hmid);
        const uint32_t *pr = prof + rc * NP * 32 + lane;
        // pass 1 (ascending, previous column's H carried in a register): H <- max(diag + s, E, 0); in the same sweep the
        // transfer of each half: chain value after its last valid row (pads only decay: compensated below)
        uint32_t a2 = 0, diag2 = up2;
        uint32_t O2[NP];
#pragma unroll
        for (int r = 0; r < NP; r++) {
            const uint32_t old = H2[r];
            H2[r] = __viaddmax_s16x2_relu(diag2, pr[r * 32], E2[r]);
            diag2 = old;
            O2[r] = __viaddmax_s16x2_relu(H2[r], NGO2, 0u);            // max(H - gapO, 0): reused by pass 2
            a2 = __viaddmax_s16x2(a2, NGE2, (O2[r] & M2[r]) | (PAD2 & ~M2[r]));
        }
        const int a_l = cl > 0 ? lo16(a2) + GE * npad_l : NEG;     // cl == 0: empty lane
        const int a_h = ch > 0 ? hi16(a2) + GE * npad_h : NEG;
        const int d_l = -GE * cl, d_h = -GE * ch;
        // lane composite, then exclusive (d, A) scans over the lanes: unrestricted chain and stripe-restricted chain
        int A = max(a_l + d_h, a_h), D = d_l + d_h;
        if (cnt == 0) { A = NEG; D = 0; }
        int af = A, df = D, al = A, dl = seg_start ? NEG : D;
        if (cnt == 0) dl = seg_start ? NEG : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int uaf = __shfl_up_sync(FULL, af, d), udf = __shfl_up_sync(FULL, df, d);
            const int ual = __shfl_up_sync(FULL, al, d), udl = __shfl_up_sync(FULL, dl, d);
            if (lane >= d) {
                af = max(af, max(uaf + df, NEG)); df = max(NEG, udf + df);
                al = max(al, max(ual + dl, NEG)); dl = max(NEG, udl + dl);
            }
        }
        int fin_full = __shfl_up_sync(FULL, af, 1), fin_loc = __shfl_up_sync(FULL, al, 1);
        if (lane == 0) { fin_full = 0; fin_loc = 0; }
        fin_full = max(fin_full, 0);
        fin_loc = seg_start ? 0 : max(fin_loc, 0);
        uint32_t ffull2 = pack2(fin_full, max(max(fin_full + d_l, a_l), 0));
        uint32_t floc2 = pack2(fin_loc, max(max(fin_loc + d_l, a_l), 0));
        // pass 2 (ascending)
        uint32_t cm2 = 0;
#pragma unroll
        for (int r = 0; r < NP; r++) {
            // open = max(max(H, floc) - gapO, 0) = max(floc - gapO, max(H - gapO, 0)); the unrestricted chain dominates the
            // restricted one, so the final H is max(H, ffull)
            const uint32_t open2 = __viaddmax_s16x2(floc2, NGO2, O2[r]);
            E2[r] = __viaddmax_s16x2(E2[r], NGE2, open2);
            floc2 = __viaddmax_s16x2(floc2, NGE2, open2);
            const uint32_t hf2 = __vmaxs2(H2[r], ffull2);
            ffull2 = __viaddmax_s16x2(ffull2, NGE2, open2);
            H2[r] = hf2;
            cm2 = __vmaxs2(cm2, hf2 & M2[r]);
        }
        // last rows of the halves (diag sources of the next column)
        uint32_t hl2 = 0, hp2 = 0;
        switch (cl) {
#define PB_CASE(k) case k + 1: if (k < NP) { hl2 = H2[k < NP ? k : 0]; hp2 = H2[(k > 0 && k < NP) ? k - 1 : 0]; } break;
            PB_CASE(0) PB_CASE(1) PB_CASE(2) PB_CASE(3) PB_CASE(4) PB_CASE(5) PB_CASE(6) PB_CASE(7) PB_CASE(8) PB_CASE(9) PB_CASE(10)
            PB_CASE(11) PB_CASE(12) PB_CASE(13) PB_CASE(14) PB_CASE(15) PB_CASE(16) PB_CASE(17) PB_CASE(18) PB_CASE(19) PB_CASE(20)
#undef PB_CASE
            default: break;
        }
        hmid = lo16(hl2);
        hlast = ch == 0 ? hmid : (ch == cl ? hi16(hl2) : hi16(hp2));
        const int colmax = __reduce_max_sync(FULL, max(lo16(cm2), hi16(cm2)));
        if (colmax > gbest) {
            gbest = colmax; gcol = c; have = true;
            if (byte_mode && colmax + BIAS >= 255) { overflow = true; break; }
#pragma unroll
            for (int r = 0; r < NP; r++) snap[r * 32 + lane] = H2[r];
        }
        if (colmax == terminate) break;
    }
    SwBest b;
    if (overflow) { b.score = 255; b.ref = 0; b.read = 0; return b; }
    if (!have) { b.score = 0; b.ref = byte_mode ? -1 : 0; b.read = readLen - 1; return b; }
    __syncwarp();
    int row = 0x7fffffff;
    for (int r = cl - 1; r >= 0; r--) if (lo16(snap[r * 32 + lane]) == gbest) row = start + r;
    if (row == 0x7fffffff) for (int r = ch - 1; r >= 0; r--) if (hi16(snap[r * 32 + lane]) == gbest) row = start + cl + r;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) row = min(row, __shfl_xor_sync(FULL, row, d));
    b.score = gbest; b.ref = ref_dir ? refLen - 1 - gcol : gcol; b.read = row;
    return b;
}

template <int NP>
__global__ void __launch_bounds__(128) k_sw16(const Task *__restrict__ tasks, const int32_t *__restrict__ order, int64_t n, const int8_t *__restrict__ codes,
                                              const int8_t *__restrict__ rcodes, Aln *__restrict__ out) {
    extern __shared__ uint32_t sw_smem[];
    const int64_t w = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    uint32_t *smem = sw_smem + (threadIdx.x >> 5) * (SW16_SMEM_WARP / 4);
    const int64_t a = order[w];
    const Task t = tasks[a];
    Aln res = {0, -1, 0, -1, 0, 0};
    if (t.q_len > 0 && t.r_len > 0) {
        const int8_t *ref = rcodes + t.r_off;
        bool word = false;
        SwBest b = sw_pass16<NP>(codes, t.q_off, 1, t.q_len, ref, t.r_len, 0, 16, true, 255, lane, smem);
        if (b.score == 255) { b = sw_pass16<NP>(codes, t.q_off, 1, t.q_len, ref, t.r_len, 0, 8, false, 65535, lane, smem); word = true; }
        res.score = b.score; res.ref_end = b.ref; res.read_end = b.read;
        if (b.score > 1 && b.ref >= 0) {
            const SwBest rb = word ? sw_pass16<NP>(codes, t.q_off + b.read, -1, b.read + 1, ref, b.ref + 1, 1, 8, false, b.score, lane, smem)
                                   : sw_pass16<NP>(codes, t.q_off + b.read, -1, b.read + 1, ref, b.ref + 1, 1, 16, true, b.score, lane, smem);
            res.ref_begin = rb.ref; res.read_begin = b.read - rb.read;
            res.status = 1;
        }
    }
    if (lane == 0) out[a] = res;
}

// ---------------------------------------------------------------------------------------------- banded_sw + cigar
__device__ __forceinline__ int score_of(int a, int b) { return (a == b && a < 4) ? MATCH : -MISM; }

// one warp per alignment; scratch slot: 3 int32 band arrays of band_words, path_words cigar words, then direction bytes.
// Small bands and short sequences live in shared memory (band arrays, base codes); the trace-back stages 32 rows x 96
// direction bytes at a time in shared memory so that the sequential walk of lane 0 never waits on HBM / L2.
constexpr int BW_SH = 128;                                  // largest band width whose arrays stay in shared memory
constexpr int BAND_SH = 2 * BW_SH + 8;
constexpr int CODE_SH = 2816;                               // read + reference codes staged per warp
constexpr int WIN_W = 96, WIN_H = 32;
struct BandShared { int32_t band[3 * BAND_SH]; int8_t code[CODE_SH]; uint8_t win[WIN_H * WIN_W]; };

__global__ void __launch_bounds__(128) k_banded(const Task *__restrict__ tasks, const int32_t *__restrict__ list, int64_t n, const int8_t *__restrict__ codes,
                                                const int8_t *__restrict__ rcodes, Aln *__restrict__ alns, uint8_t *__restrict__ scratch,
                                                int64_t slot_bytes, int band_words, int path_words, uint32_t *__restrict__ pool, unsigned long long pool_cap,
                                                unsigned long long *__restrict__ pool_used, int64_t *__restrict__ cig_off, int32_t *__restrict__ cig_len) {
    __shared__ BandShared sh_all[4];
    BandShared &sh = sh_all[threadIdx.x >> 5];
    const int64_t w = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    uint8_t *slot = scratch + w * slot_bytes;
    int32_t *gband = reinterpret_cast<int32_t *>(slot);
    uint32_t *path = reinterpret_cast<uint32_t *>(gband + 3 * (size_t) band_words);
    uint8_t *dir = slot + sizeof(int32_t) * (3 * (size_t) band_words + path_words);
    const int64_t dir_bytes = slot_bytes - (int64_t) sizeof(int32_t) * (3 * (int64_t) band_words + path_words);
    for (int64_t it = w; it < n; it += n_warps) {
        const int64_t a = list[it];
        Aln al = alns[a];
        if (al.status != 1 && al.status != 3) continue;
        const Task t = tasks[a];
        const int refLen = al.ref_end - al.ref_begin + 1, readLen = al.read_end - al.read_begin + 1;
        if (refLen <= 0 || readLen <= 0) { if (lane == 0) { alns[a].status = -1; } continue; }
        const int8_t *ref = rcodes + t.r_off + al.ref_begin;
        const int8_t *read = codes + t.q_off + al.read_begin;
        __syncwarp();
        if (refLen + readLen <= CODE_SH) {                       // stage the two sequences
            for (int k = lane; k < readLen; k += 32) sh.code[k] = read[k];
            for (int k = lane; k < refLen; k += 32) sh.code[readLen + k] = ref[k];
            read = sh.code; ref = sh.code + readLen;
        }
        __syncwarp();
        int bw = abs(refLen - readLen) + 1;
        int maxv = 0, width_d = 0;
        bool fits = true;
        for (;;) {
            const int width = bw * 2 + 3;
            width_d = bw * 2 + 1;
            if (width + 2 > band_words || (int64_t) width_d * readLen > dir_bytes) { fits = false; break; }
            const bool in_sh = width + 2 <= BAND_SH;
            int32_t *prev = in_sh ? sh.band : gband, *cur = prev + (in_sh ? BAND_SH : band_words), *e_b = cur + (in_sh ? BAND_SH : band_words);
            for (int k = lane; k < width + 2; k += 32) { prev[k] = 0; cur[k] = 0; e_b[k] = 0; }
            __syncwarp();
            int lmax = 0;
            for (int i = 0; i < readLen; i++) {
                const int beg = max(0, i - bw), end = min(refLen - 1, i + bw);
                const int edge = min(end + 1, width - 1);
                const int x = beg, xp = max(0, i - 1 - bw);
                if (lane == 0) { prev[0] = 0; e_b[0] = 0; prev[edge] = 0; e_b[edge] = 0; cur[0] = 0; }
                __syncwarp();
                const int rd = read[i];
                uint8_t *line = dir + (size_t) width_d * i;
                int hcar = 0, fcar = 0;                              // h_c[0] = 0, f = 0 at the row start
                for (int jb = beg; jb <= end; jb += 32) {
                    const int j = jb + lane;
                    const bool on = j <= end;
                    int ev = 0, de = 2, diag = 0, u = 0;
                    if (on) {
                        u = j - x + 1;
                        const int e_i = j - xp + 1;
                        const int t1 = i == 0 ? -GO : prev[e_i] - GO;
                        const int t2 = i == 0 ? -GE : e_b[e_i] - GE;
                        ev = max(t1, t2);
                        de = t1 > t2 ? 3 : 2;
                        diag = prev[e_i - 1] + score_of(ref[j], rd);
                    }
                    __syncwarp();                                   // all reads of e_b done before the in-place writes
                    const int e1 = max(ev, 0);
                    const int g = max(e1, diag);                    // h without the F term (>= 0)
                    // f(j) = max(h(j-1) - GO, f(j-1) - GE); sources: carry (lane 0) and g of earlier lanes
                    const int F0 = max(hcar - GO, fcar - GE);
                    int v = on ? g - GO + lane * GE : NEG;
                    int pm = v;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, pm, d); if (lane >= d) pm = max(pm, o); }
                    int ex = __shfl_up_sync(FULL, pm, 1);
                    int f = F0 - lane * GE;
                    if (lane > 0) f = max(f, ex - (lane - 1) * GE);
                    const int f1 = max(f, 0);
                    const int t1 = max(e1, f1);
                    const int h = max(t1, diag);
                    // direction of f: compare the previous column's true h and f
                    int hp = __shfl_up_sync(FULL, h, 1), fp = __shfl_up_sync(FULL, f, 1);
                    if (lane == 0) { hp = hcar; fp = fcar; }
                    const int df = (hp - GO > fp - GE) ? 5 : 4;
                    const int dh = (t1 <= diag) ? 1 : (e1 > f1 ? de : df);
                    if (on) {
                        e_b[u] = ev;
                        cur[u] = h;
                        line[j - x] = (uint8_t) ((de == 3 ? 1 : 0) | (df == 5 ? 2 : 0) | (dh << 2));
                        lmax = max(lmax, h);
                    }
                    const int last = min(31, end - jb);
                    hcar = __shfl_sync(FULL, h, last);
                    fcar = __shfl_sync(FULL, f, last);
                }
                __syncwarp();
                int32_t *tmp = prev; prev = cur; cur = tmp;
            }
            maxv = max(maxv, __reduce_max_sync(FULL, lmax));
            if (maxv >= al.score) break;
            if (bw > max(refLen, readLen) + 1) break;               // the band already covers everything: cannot improve
            bw *= 2;
        }
        if (!fits) { if (lane == 0) alns[a].status = 3; continue; }
        __syncwarp();
        __threadfence_block();
        // trace back from the bottom-right corner (ssw.c:665-733): 32 rows x 96 band columns are staged at a time
        int i = readLen - 1, j = refLen - 1, status = 2;
        int e = 0, l = 0, state = 2, op = 0, prev_op = 0;        // lane 0 only: 0 M, 1 I, 2 D
        while (i > 0 && status == 2) {
            const int x_top = max(0, i - bw);
            const int c0 = (j - x_top) - WIN_W / 2;              // window start in band coordinates (same for the staged rows)
            // all loads of 8 rows are issued before the first store (the walk below is sequential: this is its only HBM / L2 wait)
            for (int rr0 = 0; rr0 < WIN_H; rr0 += 8) {
                uint8_t tmp[8][WIN_W / 32];
#pragma unroll
                for (int r8 = 0; r8 < 8; r8++) {
                    const int row = i - (rr0 + r8);
                    const uint8_t *src = dir + (size_t) width_d * max(row, 0);
#pragma unroll
                    for (int k = 0; k < WIN_W / 32; k++) {
                        const int cc = c0 + 32 * k + lane;
                        tmp[r8][k] = (row >= 1 && cc >= 0 && cc < width_d) ? src[cc] : (uint8_t) 0;
                    }
                }
#pragma unroll
                for (int r8 = 0; r8 < 8; r8++)
#pragma unroll
                    for (int k = 0; k < WIN_W / 32; k++) sh.win[(rr0 + r8) * WIN_W + 32 * k + lane] = tmp[r8][k];
            }
            __syncwarp();
            if (lane == 0) {
                const int i_top = i;
                while (i > 0) {
                    const int rr = i_top - i;
                    if (rr >= WIN_H) break;
                    const int col = j - max(0, i - bw);
                    if (col < 0 || col >= width_d) { status = -2; break; }
                    const int wc = col - c0;
                    if (wc < 0 || wc >= WIN_W) break;            // drifted out of the staged window: restage
                    const int b = sh.win[rr * WIN_W + wc];
                    const int code = state == 2 ? (b >> 2) : state == 0 ? ((b & 1) ? 3 : 2) : ((b & 2) ? 5 : 4);
                    if (code == 1) { i--; j--; state = 2; op = 0; }
                    else if (code == 2) { i--; state = 0; op = 1; }
                    else if (code == 3) { i--; state = 2; op = 1; }
                    else if (code == 4) { j--; state = 1; op = 2; }
                    else if (code == 5) { j--; state = 2; op = 2; }
                    else { status = -2; break; }
                    if (op == prev_op) e++;
                    else {
                        if (l >= path_words) { status = -3; break; }
                        path[l++] = (uint32_t) e << 4 | (uint32_t) prev_op;
                        prev_op = op; e = 1;
                    }
                }
            }
            i = __shfl_sync(FULL, i, 0); j = __shfl_sync(FULL, j, 0); status = __shfl_sync(FULL, status, 0);
            __syncwarp();
        }
        if (lane == 0) {
            if (status == 2) {
                if (l + 2 > path_words) status = -3;
                else if (op == 0) path[l++] = (uint32_t) (e + 1) << 4;
                else { path[l++] = (uint32_t) e << 4 | (uint32_t) op; path[l++] = 1u << 4; }
            }
            if (status == 2) {
                // forward walk (ConvertAlignment + CalculateNumberMismatch + CigarStringToVector): count, reserve, write
                const int tail = t.q_len - al.read_end - 1;
                for (int pass = 0; pass < 2 && status == 2; pass++) {
                    int n_out = 0;
                    unsigned long long base = 0;
                    if (pass == 1) {
                        base = atomicAdd(pool_used, (unsigned long long) cig_len[a]);
                        if (base + (unsigned long long) cig_len[a] > pool_cap) { status = -4; break; }
                        cig_off[a] = (int64_t) base;
                    }
                    uint32_t *dst = pool + base;
                    const int8_t *rp = ref, *qp = read;
                    int in_m = 0, in_x = 0, len_m = 0, len_x = 0;
                    if (al.read_begin > 0) { if (pass) dst[n_out] = (uint32_t) al.read_begin << 4 | 4; n_out++; }
                    for (int k = l - 1; k >= 0; k--) {
                        const int pop = (int) (path[k] & 15), len = (int) (path[k] >> 4);
                        if (pop == 0) {
                            for (int s = 0; s < len; s++, rp++, qp++) {
                                if (*rp != *qp) {
                                    if (in_m) { if (pass) dst[n_out] = (uint32_t) len_m << 4; n_out++; }
                                    len_m = 0; len_x++; in_m = 0; in_x = 1;
                                } else {
                                    if (in_x) { if (pass) dst[n_out] = (uint32_t) len_x << 4; n_out++; }
                                    len_m++; len_x = 0; in_m = 1; in_x = 0;
                                }
                            }
                        } else {
                            if (pop == 1) qp += len; else rp += len;
                            if (in_m) { if (pass) dst[n_out] = (uint32_t) len_m << 4; n_out++; }
                            else if (in_x) { if (pass) dst[n_out] = (uint32_t) len_x << 4; n_out++; }
                            in_m = in_x = 0; len_m = len_x = 0;
                            if (pass) dst[n_out] = path[k];
                            n_out++;
                        }
                    }
                    if (in_m) { if (pass) dst[n_out] = (uint32_t) len_m << 4; n_out++; }
                    else if (in_x) { if (pass) dst[n_out] = (uint32_t) len_x << 4; n_out++; }
                    if (tail > 0) { if (pass) dst[n_out] = (uint32_t) tail << 4 | 4; n_out++; }
                    if (pass == 0) cig_len[a] = n_out;
                }
            }
            alns[a].status = status;
        }
        __syncwarp();
    }
}

// new cigar length per read: realigned (status 2, score > 1, non-empty) or the original
__global__ void k_out_counts(pb_reads_t R, const Aln *__restrict__ alns, const int32_t *__restrict__ cig_len, int32_t *__restrict__ out_nc) {
    const int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R.n_reads) return;
    const bool re = alns[r].status == 2 && alns[r].score > 1 && cig_len[r] > 0;
    out_nc[r] = re ? cig_len[r] : (int32_t) (R.cigar_off[r + 1] - R.cigar_off[r]);
}

// warp per read: final position and cigar
__global__ void __launch_bounds__(256) k_out_write(pb_reads_t R, const Aln *__restrict__ alns, const int32_t *__restrict__ cig_len,
                                                   const int64_t *__restrict__ cig_off, const uint32_t *__restrict__ pool,
                                                   const int64_t *__restrict__ o_cigar_off, int64_t *__restrict__ o_pos, uint32_t *__restrict__ o_cigar) {
    const int64_t r = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R.n_reads) return;
    const bool re = alns[r].status == 2 && alns[r].score > 1 && cig_len[r] > 0;
    const int64_t o = o_cigar_off[r], n = o_cigar_off[r + 1] - o;
    const uint32_t *src = re ? pool + cig_off[r] : R.cigar + R.cigar_off[r];
    for (int64_t k = lane; k < n; k += 32) o_cigar[o + k] = src[k];
    if (lane == 0) o_pos[r] = re ? R.pos[r] + alns[r].ref_begin : R.pos[r];
}

}  // namespace

struct pb_realigner {
    int device = 0;
    int sms = 148;
    DevBuf codes, rcodes, tasks, region_of, err, alns, order, list, scratch, pool, pool_used, cig_off, cig_len, out_nc;
    DevBuf o_pos, o_cigar_off, o_cigar, scal, slab, slab_off;
    int64_t n_long = 0;                          // reads that took the global-slab kernel in the last call
    DevBuf h_pos, h_seq_off, h_cigar_off, h_flags, h_mapq, h_seq, h_qual, h_cigar, h_regions, h_ref;
    pb_reads_t in{};
    int64_t out_cigar = 0;
    int64_t n_realigned = 0, n_sw = 0;
    float ms[2] = {0, 0};
    cudaEvent_t evt[3] = {nullptr, nullptr, nullptr};
};

extern "C" int pb_realigner_create(pb_realigner_t **out, int device) {
    if (!out) { set_error("null out"); return PB_ERR_ARG; }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= device) {
        set_error("no CUDA device %d (found %d): libpepper_b200 has no CPU fallback", device, n);
        return PB_ERR_CUDA;
    }
    PB_CUDA(cudaSetDevice(device));
    auto *t = new pb_realigner();
    t->device = device;
    cudaDeviceProp prop;
    PB_CUDA(cudaGetDeviceProperties(&prop, device));
    t->sms = prop.multiProcessorCount;
    for (auto &e : t->evt) PB_CUDA(cudaEventCreate(&e));
    *out = t;
    return PB_OK;
}

extern "C" int pb_realigner_destroy(pb_realigner_t *t) {
    if (!t) return PB_OK;
    DevBuf *bufs[] = {&t->codes, &t->rcodes, &t->tasks, &t->region_of, &t->err, &t->alns, &t->order, &t->list, &t->scratch, &t->pool,
                      &t->pool_used, &t->cig_off, &t->cig_len, &t->out_nc, &t->o_pos, &t->o_cigar_off, &t->o_cigar, &t->scal, &t->slab, &t->slab_off, &t->h_pos,
                      &t->h_seq_off, &t->h_cigar_off, &t->h_flags, &t->h_mapq, &t->h_seq, &t->h_qual, &t->h_cigar, &t->h_regions, &t->h_ref};
    for (auto *b : bufs) b->release();
    for (auto &e : t->evt) if (e) cudaEventDestroy(e);
    delete t;
    return PB_OK;
}

// host: read lengths (one D2H of seq_off) -> length-sorted order, split at the register kernel's limit
extern "C" int pb_realign_device(pb_realigner_t *t, const pb_reads_t *dr, const pb_region_t *d_regions, const pb_region_t *h_regions,
                                 int64_t n_regions, const char *d_ref, int64_t ref_bytes, pb_reads_t *out, void *stream_) {
    if (!t || !dr || !out || (!h_regions && n_regions) || (!d_regions && n_regions)) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(t->device));
    const int64_t n = dr->n_reads;
    t->in = *dr;
    *out = *dr;
    t->out_cigar = 0; t->n_realigned = 0; t->n_sw = 0; t->n_long = 0;
    if (n == 0) return PB_OK;
    std::vector<int64_t> so(n + 1), co2(2);
    PB_CUDA(cudaMemcpyAsync(so.data(), dr->seq_off, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(co2.data(), dr->cigar_off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    const int64_t nb = so[n], nc_in = co2[0];
    int64_t max_len = 0;
    for (int64_t r = 0; r < n; r++) max_len = std::max(max_len, so[r + 1] - so[r]);
    int64_t max_ref = 0;
    for (int64_t g = 0; g < n_regions; g++) max_ref = std::max(max_ref, h_regions[g].ref_len);
    PB_CUDA(cudaEventRecord(t->evt[0], st));
    PB_TRY(t->codes.reserve((size_t) nb + 16));
    PB_TRY(t->rcodes.reserve((size_t) ref_bytes + 16));
    PB_TRY(t->tasks.reserve(sizeof(Task) * n));
    PB_TRY(t->region_of.reserve(sizeof(int32_t) * n));
    PB_TRY(t->err.reserve(16));
    PB_TRY(t->alns.reserve(sizeof(Aln) * n));
    PB_TRY(t->order.reserve(sizeof(int32_t) * n));
    PB_TRY(t->cig_off.reserve(sizeof(int64_t) * n));
    PB_TRY(t->cig_len.reserve(sizeof(int32_t) * n));
    PB_TRY(t->out_nc.reserve(sizeof(int32_t) * n));
    PB_TRY(t->pool_used.reserve(16));
    PB_TRY(t->scal.reserve(64));
    PB_CUDA(cudaMemsetAsync(t->err.p, 0, 16, st));
    PB_CUDA(cudaMemsetAsync(t->cig_len.p, 0, sizeof(int32_t) * n, st));
    if (nb) k_codes_reads<<<(unsigned) ceil_div(nb, 256), 256, 0, st>>>(dr->seq, nb, t->codes.as<int8_t>());
    if (ref_bytes) k_codes_ref<<<(unsigned) ceil_div(ref_bytes, 256), 256, 0, st>>>(d_ref, ref_bytes, t->rcodes.as<int8_t>());
    k_tasks<<<(unsigned) ceil_div(n, 256), 256, 0, st>>>(*dr, d_regions, n_regions, t->tasks.as<Task>(), t->region_of.as<int32_t>(), t->err.as<int32_t>());
    PB_CUDA(cudaGetLastError());
    // longest reads first (they dominate the tail); reads that fit the register kernel first block
    std::vector<int32_t> order(n);
    for (int64_t r = 0; r < n; r++) order[r] = (int32_t) r;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return so[a + 1] - so[a] > so[b + 1] - so[b]; });
    int64_t n_big = 0;
    while (n_big < n && so[order[n_big] + 1] - so[order[n_big]] > 32 * RMAX) n_big++;
    PB_TRY(upload(t->order, order.data(), sizeof(int32_t) * n, st));
    if (n_big) {
        // reads beyond the register kernels (no upper bound on the length): one slab of 2 x ceil(len / 32) rows x 32 lanes per read
        std::vector<int64_t> slab_off((size_t) n_big + 1, 0);
        for (int64_t i = 0; i < n_big; i++) {
            const int64_t len = so[order[i] + 1] - so[order[i]];
            slab_off[i + 1] = slab_off[i] + 2 * ((len + 31) / 32) * 32;
        }
        PB_TRY(t->slab.reserve(sizeof(int32_t) * (size_t) slab_off[n_big]));
        PB_TRY(upload(t->slab_off, slab_off.data(), sizeof(int64_t) * (n_big + 1), st));
        k_sw<RDYN><<<(unsigned) ceil_div(n_big * 32, 128), 128, 0, st>>>(t->tasks.as<Task>(), t->order.as<int32_t>(), n_big, t->codes.as<int8_t>(),
                                                                        t->rcodes.as<int8_t>(), t->alns.as<Aln>(), t->slab.as<int>(), t->slab_off.as<int64_t>());
        t->n_long = n_big;
    }
    if (n - n_big) {
        static const bool force_i32 = getenv("PB_REALIGN_I32") && atoi(getenv("PB_REALIGN_I32")) != 0;     // debug: int32 kernel for every read
        if (force_i32) {
            k_sw<RMAX><<<(unsigned) ceil_div((n - n_big) * 32, 128), 128, 0, st>>>(t->tasks.as<Task>(), t->order.as<int32_t>() + n_big, n - n_big,
                                                                                  t->codes.as<int8_t>(), t->rcodes.as<int8_t>(), t->alns.as<Aln>());
        } else {
            // length-sorted order: launch each length class with the smallest register footprint that holds it
            const int64_t lim[4] = {32 * 2 * 21, 32 * 2 * 16, 32 * 2 * 12, 32 * 2 * 8};
            int64_t lo = n_big;
            for (int cls = 0; cls < 4; cls++) {
                int64_t hi = lo;
                const int64_t next_lim = cls < 3 ? lim[cls + 1] : -1;
                while (hi < n && so[order[hi] + 1] - so[order[hi]] > next_lim) hi++;
                const int64_t m = hi - lo;
                if (m > 0) {
                    const unsigned grid = (unsigned) ceil_div(m * 32, 128);
                    const int32_t *ord = t->order.as<int32_t>() + lo;
#define PB_LAUNCH_SW16(NPT)                                                                                                          \
    do {                                                                                                                             \
        PB_CUDA(cudaFuncSetAttribute(k_sw16<NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SW16_SMEM_WARP));                 \
        k_sw16<NPT><<<grid, 128, 4 * SW16_SMEM_WARP, st>>>(t->tasks.as<Task>(), ord, m, t->codes.as<int8_t>(), t->rcodes.as<int8_t>(), \
                                                           t->alns.as<Aln>());                                                       \
    } while (0)
                    if (cls == 0) PB_LAUNCH_SW16(21);
                    else if (cls == 1) PB_LAUNCH_SW16(16);
                    else if (cls == 2) PB_LAUNCH_SW16(12);
                    else PB_LAUNCH_SW16(8);
#undef PB_LAUNCH_SW16
                }
                lo = hi;
            }
        }
    }
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaEventRecord(t->evt[1], st));
    int32_t h_err = 0;
    PB_CUDA(cudaMemcpyAsync(&h_err, t->err.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    if (h_err == 1) { set_error("a read starts before its region start: the reference drops such reads (simple_aligner.cpp:73-77); fetch reads with get_reads(start = region start)"); return PB_ERR_ARG; }
    if (h_err == 2) { set_error("region read ranges do not cover the reads in order"); return PB_ERR_ARG; }
    // banded pass: persistent warps with one scratch slot each; alignments that outgrow the slot are retried with bigger slots
    const int64_t pool_cap = nb + 4 * n + 1024;             // worst case: one tuple per base (alternating = / X) + clips
    PB_TRY(t->pool.reserve(sizeof(uint32_t) * (size_t) pool_cap));
    PB_CUDA(cudaMemsetAsync(t->pool_used.p, 0, 16, st));
    int band_bw = 128;
    int64_t warps = (int64_t) t->sms * 24;
    for (int attempt = 0; attempt < 6; attempt++) {
        const int band_words = 2 * band_bw + 8;
        const int path_words = (int) (max_len + max_ref + 16);
        const int64_t slot_bytes = ((int64_t) sizeof(int32_t) * (3 * band_words + path_words) + (int64_t) (2 * band_bw + 1) * (max_len + 1) + 255) / 256 * 256;
        warps = std::max<int64_t>(4, std::min<int64_t>(warps, (int64_t) (6ll << 30) / slot_bytes));
        warps = std::min<int64_t>(warps, n);
        PB_TRY(t->scratch.reserve((size_t) (slot_bytes * warps)));
        k_banded<<<(unsigned) ceil_div(warps * 32, 128), 128, 0, st>>>(t->tasks.as<Task>(), t->order.as<int32_t>(), n, t->codes.as<int8_t>(), t->rcodes.as<int8_t>(),
                                                                     t->alns.as<Aln>(), t->scratch.as<uint8_t>(), slot_bytes, band_words, path_words, t->pool.as<uint32_t>(),
                                                                     (unsigned long long) pool_cap, t->pool_used.as<unsigned long long>(),
                                                                     t->cig_off.as<int64_t>(), t->cig_len.as<int32_t>());
        PB_CUDA(cudaGetLastError());
        // any alignment left in state 3 (slot too small)?  check on the host
        std::vector<Aln> h(n);
        PB_CUDA(cudaMemcpyAsync(h.data(), t->alns.p, sizeof(Aln) * n, cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        int64_t left = 0, bad = 0;
        for (auto &x : h) { if (x.status == 3) left++; if (x.status < 0) bad++; }
        if (bad) { set_error("realignment failed for %lld reads (trace-back / pool overflow)", (long long) bad); return PB_ERR_STATE; }
        if (!left) {
            t->n_sw = n;
            for (auto &x : h) if (x.status == 2 && x.score > 1) t->n_realigned++;
            break;
        }
        if (band_bw > max_len + max_ref + 2) { set_error("band could not be grown enough"); return PB_ERR_STATE; }
        band_bw *= 4;
    }
    // assemble the output cigar arrays
    k_out_counts<<<(unsigned) ceil_div(n, 256), 256, 0, st>>>(*dr, t->alns.as<Aln>(), t->cig_len.as<int32_t>(), t->out_nc.as<int32_t>());
    PB_CUDA(cudaGetLastError());
    PB_TRY(t->o_cigar_off.reserve(sizeof(int64_t) * (n + 1)));
    k_scan_excl<<<1, 1024, 0, st>>>(t->out_nc.as<int32_t>(), t->o_cigar_off.as<int64_t>(), n, t->scal.as<int64_t>());
    PB_CUDA(cudaGetLastError());
    int64_t tot = 0;
    PB_CUDA(cudaMemcpyAsync(&tot, t->scal.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    t->out_cigar = tot;
    PB_TRY(t->o_cigar.reserve(sizeof(uint32_t) * (size_t) (tot + 4)));
    PB_TRY(t->o_pos.reserve(sizeof(int64_t) * n));
    k_out_write<<<(unsigned) ceil_div(n * 32, 256), 256, 0, st>>>(*dr, t->alns.as<Aln>(), t->cig_len.as<int32_t>(), t->cig_off.as<int64_t>(), t->pool.as<uint32_t>(),
                                                                 t->o_cigar_off.as<int64_t>(), t->o_pos.as<int64_t>(), t->o_cigar.as<uint32_t>());
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaEventRecord(t->evt[2], st));
    PB_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&t->ms[0], t->evt[0], t->evt[1]);
    cudaEventElapsedTime(&t->ms[1], t->evt[1], t->evt[2]);
    (void) nc_in;
    out->pos = t->o_pos.as<int64_t>();
    out->cigar_off = t->o_cigar_off.as<int64_t>();
    out->cigar = t->o_cigar.as<uint32_t>();
    return PB_OK;
}

extern "C" int pb_realign_host(pb_realigner_t *t, const pb_reads_t *h, const pb_region_t *h_regions, int64_t n_regions, const char *h_ref,
                               int64_t ref_bytes, int64_t *o_pos, int64_t *o_cigar_off, uint32_t *o_cigar, int64_t cigar_capacity,
                               int64_t *n_cigar, void *stream_) {
    if (!t || !h || !n_cigar) { set_error("null argument"); return PB_ERR_ARG; }
    cudaStream_t st = (cudaStream_t) stream_;
    PB_CUDA(cudaSetDevice(t->device));
    DevBuf *rb[8] = {&t->h_pos, &t->h_seq_off, &t->h_cigar_off, &t->h_flags, &t->h_mapq, &t->h_seq, &t->h_qual, &t->h_cigar};
    pb_reads_t d;
    PB_TRY(upload_reads(h, rb, &d, st));
    PB_TRY(upload(t->h_regions, h_regions, sizeof(pb_region_t) * n_regions, st));
    PB_TRY(upload(t->h_ref, h_ref, (size_t) ref_bytes, st));
    pb_reads_t o;
    PB_TRY(pb_realign_device(t, &d, t->h_regions.as<pb_region_t>(), h_regions, n_regions, t->h_ref.as<char>(), ref_bytes, &o, stream_));
    const int64_t n = h->n_reads;
    *n_cigar = n ? t->out_cigar : 0;
    if (n == 0) { if (o_cigar_off) o_cigar_off[0] = 0; return PB_OK; }
    if (t->out_cigar > cigar_capacity) return PB_ERR_CAPACITY;
    PB_CUDA(cudaMemcpyAsync(o_pos, o.pos, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(o_cigar_off, o.cigar_off, sizeof(int64_t) * (n + 1), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaMemcpyAsync(o_cigar, o.cigar, sizeof(uint32_t) * t->out_cigar, cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    return PB_OK;
}

extern "C" int pb_realign_stats(pb_realigner_t *t, int64_t *n_aligned, int64_t *n_realigned, float *ms_sw, float *ms_cigar) {
    if (!t) return PB_ERR_ARG;
    if (n_aligned) *n_aligned = t->n_sw;
    if (n_realigned) *n_realigned = t->n_realigned;
    if (ms_sw) *ms_sw = t->ms[0];
    if (ms_cigar) *ms_cigar = t->ms[1];
    return PB_OK;
}
