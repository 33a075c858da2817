// Device code of libindelgpu.so: split-read realignment kernels for sm_100a.
//
// Shared pieces: reference packing, the one-diagonal alignment (band 1), junction scoring and
// segment stitching.  The warp-per-read pipeline that uses them is in realign_kernel.cuh, the
// k-mer vote in warp_vote.cuh, the banded DP in band_dp.cuh.
// Everything between the read bytes coming in and the segment words going out stays in
// shared memory / registers; the reference windows come from the resident 2-bit packed copy
// (voting, TMA-staged) and the raw bytes (DP) in HBM/L2.
//
// Reference behaviour reproduced (file:line of ratan-lab/indelMINER):
//   find_best_band        src/alignment.c:393-447 (+ :29-181)
//   local_align           src/localalign.c:15-196
//   ALIGN / align         src/globalalign.c:66-401
//   fetch_cigar           src/globalalign.c:507-604
//   attempt_diagonal_alignments  src/alignment.c:539-759
//   update_readsegs       src/readaln.c:348-458
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/indelgpu.h"

namespace indelgpu {

constexpr int kThreads = 128;             // threads per CTA of the realign kernel
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;
constexpr int kNeg = -9999999;            // MININT (localalign.c:3, globalalign.h:16)

enum { OP_INS = 1, OP_DEL = 2, OP_SOFT = 4, OP_EQ = 7, OP_X = 8 };   // bam.h:138-155, readaln.h:10-11
constexpr int ST_ASSERT = 7;              // the reference would have hit a forceassert / exit()

struct DevParams {
    int k, g, maxdel, ethr;
    int match, mismatch, G, H;
    uint32_t kmask;                       // 2k low bits
};

struct RefView {
    const uint8_t* raw;                   // concatenated contigs, each starting on a 64-base boundary
    const uint32_t* packed;               // 2 bits per base, 16 bases per word, non-ACGT -> 0
    const int64_t* contig_off;            // base offset of each contig in raw
    const int64_t* contig_len;
    int ncontigs;
};

// shared-memory layout, computed identically on host and device
struct SmemLayout {
    int ht_slots;        // power of two
    int hist_words;      // uint32 words (2 x uint16 counters each)
    int read_bytes;      // padded
    int ops_cap;         // words per CIGAR buffer
    int off_keys, off_vals, off_hist, off_read, off_bits, off_psum, off_cig1, off_cig2, off_segs;
    int total;
};

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

__host__ __device__ inline SmemLayout make_layout(int max_read, int max_numdiag)
{
    SmemLayout L;
    int ht = 256;
    #pragma unroll 1
    while (ht < 4 * max_read) ht <<= 1;
    L.ht_slots = ht;
    L.hist_words = round_up(max_numdiag + 2, 8) / 2 + 4;
    L.read_bytes = round_up(max_read + 32, 16);
    L.ops_cap = max_read + 4;
    int o = 0;
    L.off_keys = o; o += ht * 4;
    L.off_vals = o; o += ht * 4;
    L.off_hist = o; o += round_up(L.hist_words * 4, 16);
    L.off_read = o; o += L.read_bytes;
    L.off_bits = o; o += round_up(((max_read + 127) / 128 * 4 + 4) * 4, 16);
    L.off_psum = o; o += round_up((max_read + 2) * 4, 16);
    L.off_cig1 = o; o += round_up(L.ops_cap * 4, 16);
    L.off_cig2 = o; o += round_up(L.ops_cap * 4, 16);
    L.off_segs = o; o += round_up((2 * L.ops_cap + 4) * 4, 16);
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------------
// 2-bit packing of the reference (once per upload)
// ---------------------------------------------------------------------------------------

// base2bits of alignment.c:11-24 up to a relabelling (only equality of k-mers matters):
// A/a -> 0, C/c -> 1, T/t -> 2, G/g -> 3, every other byte -> 0 (same as A).
__host__ __device__ inline uint32_t base_code(uint32_t b)
{
    uint32_t u = b & 0xDFu;
    uint32_t idx = u - 0x41u;
    uint32_t ok = (idx < 20u) ? ((0x80045u >> idx) & 1u) : 0u;
    return ok ? ((u >> 1) & 3u) : 0u;
}

__global__ void pack_reference_kernel(const uint8_t* __restrict__ raw, uint32_t* __restrict__ packed,
                                      int64_t nwords)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    #pragma unroll 1
    for (; w < nwords; w += stride) {
        const uint4 v = reinterpret_cast<const uint4*>(raw)[w];
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
        uint32_t out = 0;
#pragma unroll
        for (int t = 0; t < 16; t++) {
            uint32_t b = (q[t >> 2] >> ((t & 3) * 8)) & 0xFFu;
            out |= base_code(b) << (2 * t);
        }
        packed[w] = out;
    }
}

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t hash_slot(uint32_t code, int mask)
{
    return ((code * 2654435761u) >> 12) & (uint32_t)mask;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

struct Cta {
    // dynamic shared memory views
    uint32_t* keys; uint32_t* vals; uint32_t* hist; uint8_t* read; uint32_t* bits; int* psum;
    uint32_t* cig1; uint32_t* cig2; uint32_t* segs;
    SmemLayout L;
};

// result of one attempt_band_alignment (alignment.c:343-391); coordinates already absolute
struct Aln {
    int low, up, score, r1, r2, q1, q2, n;
    int cells_fwd, cells_rev, cells_glob;
};

// ---------------------------------------------------------------------------------------
// band of one diagonal: local_align degenerates to a max-segment scan (SURVEY.md 8a'),
// ALIGN takes its all-REP exit (globalalign.c:358-365).  Executed by warp 0.
// ---------------------------------------------------------------------------------------
__device__ void align_diag1(const DevParams& P, Cta& S, const uint8_t* __restrict__ win,
                            int N, int zs2, int M, int d, uint32_t* cig, int* s_out /* >= 9 ints */)
{
    const int lane = threadIdx.x & 31;
    const uint8_t* r = S.read + zs2;
    const int si = max(0, -d), ei = min(M, N - d);               // localalign.c:86-87
    // Forward sweep (localalign.c:100-131) as a prefix-sum problem: run_i = P_i - min_{j<=i} P_j.
    // Four consecutive rows per lane: 128 rows per step, one warp scan of the lane totals and one of
    // the lane minima.
    int best = 0, endi = si;
    int carryP = 0, carryMin = 0;
    if (lane == 0) S.psum[0] = 0;
    const int rows = ei - si;
    #pragma unroll 1
    for (int base = 0; base < rows; base += 128) {
        const int t0 = base + 4 * lane;                          // row index of this lane's first cell (row i = si + 1 + t)
        const int nval = min(4, max(0, rows - t0));
        uint32_t x = 0xFFFFFFFFu;                                // read bytes XOR window bytes
        if (nval > 0) {
            const uintptr_t ra = reinterpret_cast<uintptr_t>(r + si + t0), wa = reinterpret_cast<uintptr_t>(win + si + d + t0);
            const uint32_t* rp = reinterpret_cast<const uint32_t*>(ra & ~(uintptr_t)3);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(wa & ~(uintptr_t)3);
            const uint32_t rv = __funnelshift_r(rp[0], rp[1], 8 * (int)(ra & 3));
            const uint32_t wv = __funnelshift_r(__ldg(wp), __ldg(wp + 1), 8 * (int)(wa & 3));
            x = rv ^ wv;
        }
        int p[4]; uint32_t nib = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool eq = ((x >> (8 * j)) & 0xFFu) == 0u;
            const int w = j < nval ? (eq ? P.match : P.mismatch) : 0;
            if (eq && j < nval) nib |= 1u << j;
            p[j] = (j ? p[j - 1] : 0) + w;
        }
        int tot = p[3];                                          // inclusive scan of the lane totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, tot, o); if (lane >= o) tot += t; }
        const int offs = carryP + tot - p[3];
#pragma unroll
        for (int j = 0; j < 4; j++) p[j] += offs;                // prefix sums P_i
        int mn = min(min(p[0], p[1]), min(p[2], p[3]));          // inclusive scan of the lane minima
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, mn, o); if (lane >= o) mn = min(mn, t); }
        int prev = __shfl_up_sync(0xFFFFFFFFu, mn, 1);
        prev = lane ? min(prev, carryMin) : carryMin;            // min over everything before this lane (P_si = 0 included)
        int lrun = -1, lj = 0, m = prev;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            m = min(m, p[j]);
            const int run = j < nval ? p[j] - m : -1;            // max(0, run + w) recursion
            if (run > lrun) { lrun = run; lj = j; }
            if (j < nval) S.psum[t0 + j + 1] = p[j];
        }
        const int cmax = __reduce_max_sync(0xFFFFFFFFu, lrun);
        if (cmax > best) {                                       // strict: the first maximum wins
            const uint32_t who = __ballot_sync(0xFFFFFFFFu, lrun == cmax);
            const int src = __ffs(who) - 1;
            best = cmax; endi = si + base + 4 * src + __shfl_sync(0xFFFFFFFFu, lj, src) + 1;
        }
        const uint32_t v = nib << (4 * (lane & 7));
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const uint32_t word = __reduce_or_sync(0xFFFFFFFFu, (lane >> 3) == w ? v : 0u);
            if (lane == 0) S.bits[(base >> 5) + w] = word;
        }
        carryP = __shfl_sync(0xFFFFFFFFu, p[3], 31);
        carryMin = min(carryMin, __shfl_sync(0xFFFFFFFFu, mn, 31));
    }
    __syncwarp();
    int starti = 0;
    if (best > 0) {
        // reverse sweep (localalign.c:144-176): first row, walking down from endi, whose suffix sum equals best
        const int target = S.psum[endi - si] - best;
        #pragma unroll 1
        for (int base = endi; base > si; base -= 32) {
            const int i = base - lane;
            const bool hit = i > si && S.psum[i - 1 - si] == target;
            const uint32_t who = __ballot_sync(0xFFFFFFFFu, hit);
            if (who) { starti = base - (__ffs(who) - 1); break; }
        }
    }
    const bool none = best <= 0 || starti == 0 || endi == starti;   // localalign.c:191-193
    if (lane == 0) {
        int n = 0;
        if (!none) {
            // fetch_cigar (globalalign.c:507-604) on an all-REP script
            if (starti - 1 > 0) cig[n++] = ((uint32_t)(starti - 1) << 4) | OP_SOFT;
            int pos = starti - 1 - si, end = endi - 1 - si;      // bit positions in S.bits
            #pragma unroll 1
            while (pos <= end) {
                const uint32_t word = S.bits[pos >> 5];
                const int bit = (word >> (pos & 31)) & 1;
                // length of the run starting at pos
                int q = pos;
                #pragma unroll 1
                while (true) {
                    const uint32_t wq = S.bits[q >> 5];
                    uint32_t diff = (bit ? ~wq : wq) >> (q & 31);   // 1 where the run is broken
                    const int room = 32 - (q & 31);
                    if (diff) { q += __ffs(diff) - 1; break; }
                    q += room;
                    if (q > end) break;
                }
                if (q > end + 1) q = end + 1;
                cig[n++] = ((uint32_t)(q - pos) << 4) | (bit ? OP_EQ : OP_X);
                pos = q;
            }
            if (M - endi > 0) cig[n++] = ((uint32_t)(M - endi) << 4) | OP_SOFT;
        }
        s_out[0] = none ? 0 : best;
        s_out[1] = none ? 0 : starti;        // q1 (1-based inclusive)
        s_out[2] = none ? 0 : starti + d;    // r1
        s_out[3] = none ? 0 : endi;          // q2
        s_out[4] = none ? 0 : endi + d;      // r2
        s_out[5] = n;
        s_out[6] = ei - si;                                  // forward cells
        s_out[7] = best > 0 ? endi - starti + 1 : 0;         // reverse cells until the hit
        s_out[8] = 0;                                        // ALIGN exits before any sweep (band <= 1)
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// scalar pieces of attempt_diagonal_alignments, executed by one thread
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int cig_op(uint32_t c) { return (int)(c & 15u); }
__device__ __forceinline__ int cig_len(uint32_t c) { return (int)(c >> 4); }

// alignment.c:219-303
__device__ int count_matches(const uint32_t* c1, int n1, int q1, int q2,
                             const uint32_t* c2, int n2, int q3, int q4, int* pmm)
{
    int i, j, matches = 0, mm = 0;
    #pragma unroll 1
    for (i = 0, j = q1; i < n1; i++) {
        const int len = cig_len(c1[i]), op = cig_op(c1[i]);
        if (op != OP_DEL) j += len;
        if (j < q2) { if (op == OP_EQ) matches += len; else if (op == OP_X) mm += len; }
        if (j >= q2) {
            if (op == OP_EQ) matches += q2 - (j - len); else if (op == OP_X) mm += q2 - (j - len);
            break;
        }
    }
    #pragma unroll 1
    for (i = 0, j = 0; i < n2; i++) {
        const int len = cig_len(c2[i]), op = cig_op(c2[i]);
        if (op != OP_DEL) j += len;
        if (j >= q3) {
            if (op == OP_EQ) matches += j - q3; else if (op == OP_X) mm += j - q3;
            i++; break;
        }
    }
    #pragma unroll 1
    for (; i < n2; i++) {
        const int len = cig_len(c2[i]), op = cig_op(c2[i]);
        if (op != OP_DEL) j += len;
        if (j < q4) { if (op == OP_EQ) matches += len; else if (op == OP_X) mm += len; }
        if (j >= q4) {
            if (op == OP_EQ) matches += q4 - (j - len); else if (op == OP_X) mm += q4 - (j - len);
            break;
        }
    }
    *pmm = mm;
    return matches;
}

// alignment.c:306-339, candidates spread over the lanes of warp 0:
// arg-max by (matches desc, mismatches asc, i asc) == the reference's first-wins scan with its early exit
__device__ int best_junction_warp(int q1, int q2, const uint32_t* c1, int n1,
                                  int q3, int q4, const uint32_t* c2, int n2)
{
    const int lane = threadIdx.x & 31;
    unsigned long long best = 0;
    #pragma unroll 1
    for (int base = q3; base <= q2; base += 32) {
        const int i = base + lane;
        unsigned long long key = 0;
        if (i <= q2) {
            int mm;
            const int matches = count_matches(c1, n1, q1, i, c2, n2, i, q4, &mm);
            key = ((unsigned long long)(uint32_t)(matches + 1) << 42) |
                  ((unsigned long long)(0x1FFFFFu - (uint32_t)mm) << 21) |
                  (unsigned long long)(0x1FFFFFu - (uint32_t)(i - q3));
        }
        key = warp_max_u64(key);
        best = key > best ? key : best;
    }
    return q3 + (int)(0x1FFFFFu - (uint32_t)(best & 0x1FFFFFu));
}

struct SegWriter {
    uint32_t* segs; int n; int refindx; int readindx;
    __device__ void emit(uint32_t c)
    {
        const int op = cig_op(c), len = cig_len(c);              // new_readseg, readaln.c:24-99
        segs[n++] = c;
        if (op == OP_EQ || op == OP_X) { readindx += len; refindx += len; }
        else if (op == OP_DEL) refindx += len;
        else readindx += len;
    }
};

// update_readsegs (readaln.c:348-458): returns the number of segment words
__device__ int stitch_segments(uint32_t* segs, int r1, const uint32_t* c1, int n1, int index,
                               int q2, int r2, const uint32_t* c2, int n2)
{
    SegWriter w{segs, 0, r1, 0};
    int i, j;
    #pragma unroll 1
    for (i = 0, j = 0; i < n1; i++) {
        const int op = cig_op(c1[i]), len = cig_len(c1[i]);
        if (op != OP_DEL) j += len;
        if (j <= index) w.emit(c1[i]);
        if (j > index) {
            const int part = index - (j - len);
            if (part > 0) w.emit(((uint32_t)part << 4) | (uint32_t)op);
            break;
        }
    }
    int rindex = r2, nextindex = index;
    if (index >= q2) {
        int offset = 0;
        #pragma unroll 1
        for (i = 0, j = 0; i < n2; i++) {
            const int op = cig_op(c2[i]), len = cig_len(c2[i]);
            if (op != OP_DEL) j += len;
            if (j <= q2) { }
            else if (j <= index) {
                if (op != OP_INS) { offset += len; if ((j - len) <= q2) offset -= q2 - (j - len); }
            } else {
                if (op != OP_INS && (j - len) <= index) offset += index - (j - len);
            }
        }
        rindex = r2 + offset;
    } else {
        w.emit(((uint32_t)(q2 - index) << 4) | OP_INS);
        nextindex += q2 - index;
    }
    if (w.refindx < rindex) w.emit(((uint32_t)(rindex - w.refindx) << 4) | OP_DEL);
    #pragma unroll 1
    for (i = 0, j = 0; i < n2; i++) {
        const int op = cig_op(c2[i]), len = cig_len(c2[i]);
        if (op != OP_DEL) j += len;
        if (j > nextindex) { w.emit(((uint32_t)(j - nextindex) << 4) | (uint32_t)op); i++; break; }
    }
    #pragma unroll 1
    for (; i < n2; i++) w.emit(c2[i]);
    return w.n;
}

}  // namespace indelgpu
