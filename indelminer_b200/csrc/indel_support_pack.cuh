// Row f1 of SURVEY.md section 8, second generation: realign_with_indel (src/variant.c:1246-1424) as a two-pass
// wavefront on 16-bit SIMD lanes.
//
// indel_support.cuh carries the three path counters forward with every score (26 integer instructions per cell, the
// ALU pipe's limit).  Here the scores of TWO pairs ride in the halves of one register (s16x2 DPX instructions:
// VIADDMNMX.U16x2, VIMNMX3.U16x2), pass 1 (indel_support_pack_kernel, a wavefront over a segment of lanes) writes
// three bits per cell -- which predecessor the reference's walk back would take, and whether the score is positive --
// to a scratch buffer, and pass 2 (indel_support_walk_kernel) is the reference's walk back itself
// (variant.c:1405-1417), one THREAD per pair, reading those bits.
//
// Score domain.  A value x of cell (i, j) (i = row = query index, j = column = target index, both 1-based) is held
// as  x4 = 4 * (x + i + j + BIAS)  per 16-bit half.
//  * The shift by the anti-diagonal turns both gap recurrences into ONE add-max each:
//        ifins~[i][j] = max(ifins~[i-1][j], V~[i-1][j] - 4)          (variant.c:1331: max(F, up - open) - extend)
//        ifdel~[i][j] = max(ifdel~[i][j-1], V~[i][j-1] - 4)          (variant.c:1333)
//        ifsub~[i][j] = V~[i-1][j-1] + (match ? 4 : 1)               (variant.c:1325: +2 / -1)
//  * The factor 4 leaves two low bits for the tie rules of variant.c:1336-1342 -- the substitution wins a tie against
//    a gap, the insertion wins a tie against the deletion: the three candidates carry 2, 1 and 0 there, ONE
//    three-input maximum picks the value, and its low bits say which candidate it was.
//  * BIAS keeps every half non-negative (V + i + j >= 0: a row's E restarts at 0, variant.c:1322), so plain 32-bit adds
//    never carry between the halves.
// The maximum and the sign test need V itself: key = 16 * (V + 2047) + (15 - column within the lane), two
// multiply-adds on the otherwise idle FMA pipe; bit 15 of a key says V > 0, and the largest key of a lane's row is its
// best cell with the reference's "first maximum" order inside the row.
//
// Lengths: target <= 512, query <= kPackMaxQuery = 500 (the 16-bit ranges: 16 * (V + 2047) needs |V| < 2048; a
// position is 9 + 10 bits); anything longer goes to indel_support_kernel, one pair per thread.  Reproduced quirks as in indel_support.cuh.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "indel_support.cuh"

namespace indelgpu {

enum { kPackMaxQuery = 500 };       // = kWaveMaxQuery: 9 bits of row in a position, V + 2047 in 12 bits
enum : unsigned { kPkBias = 64u, kPkPastTarget = 0xFFF0u, kPkPastQuery = 0xFFE0u };

__host__ __device__ __forceinline__ unsigned pk2(unsigned lo, unsigned hi) { return (lo & 0xFFFFu) | (hi << 16); }
__host__ __device__ __forceinline__ unsigned pk_both(unsigned v) { return v * 0x00010001u; }

// bytes 1 and 3 of x, then bytes 1 and 3 of y, each replaced by eight copies of its top bit
__host__ __device__ __forceinline__ unsigned pk_signs(unsigned x, unsigned y)
{
#ifdef __CUDA_ARCH__
    unsigned r;                                            // selector nibbles 9, B, D, F: bytes 1, 3, 5, 7 in sign mode
    asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(r) : "r"(x), "r"(y));
    return r;
#else
    unsigned r = 0;
    if (x & 0x00008000u) r |= 0x000000FFu;
    if (x & 0x80000000u) r |= 0x0000FF00u;
    if (y & 0x00008000u) r |= 0x00FF0000u;
    if (y & 0x80000000u) r |= 0xFF000000u;
    return r;
#endif
}

// a * b + c on the FMA pipe (the ALU pipe is the one this kernel fills); b is a run-time value where a constant would
// let the compiler turn the multiply-add back into an ALU add
__host__ __device__ __forceinline__ unsigned pk_mad(unsigned a, unsigned b, unsigned c)
{
#ifdef __CUDA_ARCH__
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#else
    return a * b + c;
#endif
}

// direction words per lane and step: the 2-bit codes of up to 8 columns per word (2 = substitution, 1 = insertion,
// 0 = deletion; first column in the highest bits used), then one word of "V > 0" bits
__host__ __device__ __forceinline__ int pack_planes(int cpl) { return cpl <= 8 ? 2 : 3; }

// the state one lane keeps for its CPL columns of two pairs (low half: pair A, high half: pair B)
template <int CPL>
struct PackLane {
    unsigned A[CPL];        // target codes (upper-cased byte << 4; past the target: matches nothing)
    unsigned V[CPL];        // V4 + 18 of the previous row (the substitution's constant, so that it is one multiply-add)
    unsigned F[CPL];        // ifins4 + 1 of the previous row
    unsigned best;          // largest key so far, per half
    unsigned rowA, rowB;    // the row it was seen in
    unsigned vdiag;         // V4[i-1][j0] for the row this lane does next
    unsigned outV, outE;    // the right-hand boundary of the row just done

    __host__ __device__ __forceinline__ void init(const uint8_t* t1a, int len1a, const uint8_t* t1b, int len1b, int j0)
    {
#pragma unroll
        for (int c = 0; c < CPL; c++) {
            const int j = j0 + c;
            const unsigned ca = j < len1a ? (unsigned)up_case(t1a[j]) << 4 : (unsigned)kPkPastTarget;
            const unsigned cb = j < len1b ? (unsigned)up_case(t1b[j]) << 4 : (unsigned)kPkPastTarget;
            A[c] = pk2(ca, cb);
            V[c] = pk_both(4 * (kPkBias - 4) + 18);                // row 0: V = -4 - j (variant.c:1303-1306), shifted by j; V is kept + 18
            F[c] = pk_both(4 * (kPkBias + (unsigned)(j + 1)) + 1);   // F = 0, shifted by j; an insertion's tie bits
        }
        best = pk_both(2047u << 4); rowA = rowB = 0;               // score 0: nothing yet
        vdiag = pk_both(4 * (kPkBias - 4) + 18);
        outV = outE = 0;
    }

    // row i (1-based) of this lane's columns.  inV / inE: V4 + 18 and ifdel4 of (i, j0) from the lane on the left (first
    // lane: the column-0 values); nb: the negated query codes of row i; one: the number 1 at run time.
    // w[]: the direction words of the row's cells (pack_planes(CPL) of them).
    __host__ __device__ __forceinline__ void row(int i, int j0, bool first, unsigned inV, unsigned inE, unsigned nb, unsigned one,
                                                 unsigned* w)
    {
        unsigned left = first ? pk_both(4 * (kPkBias - 4) + 18) : inV;          // V[i][0] = -4 - i, shifted by i (+ 18)
        unsigned E = first ? pk_both(4 * (kPkBias + (unsigned)i)) : inE;         // E restarts at 0 on every row (:1322)
        unsigned diag = vdiag;
        vdiag = left;
        const unsigned rowK = pk_both((2047u - kPkBias - 17u - (unsigned)i - (unsigned)j0) * 16u);
        const unsigned mone = 0u - one;                     // -1 at run time: keeps v4 - code a multiply-add
        unsigned acc0 = 0, acc1 = 0, P = 0, rowbest = 0;
#pragma unroll
        for (int c = 0; c < CPL; c += 2) {
            unsigned key[2];
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const unsigned up = V[c + d];
                const unsigned m12 = __viaddmin_u16x2(A[c + d], nb, 0x000C000Cu);      // 0 if the bases match, else 12
                const unsigned ifsub = pk_mad(m12, mone, diag);                        // (V + 18) - m12: 4 * (4 | 1) + tie bits 2
                const unsigned ifins = __viaddmax_u16x2(up, 0xFFDFFFDFu, F[c + d]);   // (V + 18) - 33: 4 * -4 + tie bits 1
                F[c + d] = ifins;
                const unsigned ifdel = __viaddmax_u16x2(left, 0xFFDEFFDEu, E);         // (V + 18) - 34: tie bits 0
                E = ifdel;
                const unsigned v4 = __vimax3_u16x2(ifsub, ifins, ifdel);              // variant.c:1336-1342 in one instruction
                const unsigned code = v4 & 0x00030003u;                                // which candidate it was
                const unsigned vc = pk_mad(code, mone, v4);                            // v4 - code, on the FMA pipe
                if (c + d < 8) acc0 = pk_mad(acc0, 4u, code); else acc1 = pk_mad(acc1, 4u, code);
                // 16 * (V + 2047) + 15 - (c + d):  vc * 4 = 16 * (V + i + j0 + c + d + 1 + BIAS)
                key[d] = pk_mad(pk_mad(vc, 4u, rowK), one, pk_both((unsigned)((16 - (c + d)) * 16 + 15 - (c + d))));
                const unsigned v18 = pk_mad(vc, one, 0x00120012u);
                V[c + d] = v18; diag = up; left = v18;
            }
            rowbest = __vimax3_u16x2(rowbest, key[0], key[1]);
            P |= pk_signs(key[0], key[1]) & (0x01010101u << (c >> 1));
        }
        // strictly better than every earlier row of this lane?  (inside the row the column bits of the key decide)
        bool ph, pl;
        (void)__vibmax_u16x2(best & 0xFFF0FFF0u, rowbest & 0xFFF0FFF0u, &ph, &pl);
        if (!pl) { best = (best & 0xFFFF0000u) | (rowbest & 0x0000FFFFu); rowA = (unsigned)i; }
        if (!ph) { best = (best & 0x0000FFFFu) | (rowbest & 0xFFFF0000u); rowB = (unsigned)i; }
        w[0] = acc0;
        if (CPL > 8) w[1] = acc1;
        w[pack_planes(CPL) - 1] = P;
        outV = left; outE = E;
    }

    // this lane's best cell of one half as  score << 19 | ~((i << 10) | j),  0 if no score is positive
    __host__ __device__ __forceinline__ unsigned best_key(int h, int j0) const
    {
        const unsigned b = h ? best >> 16 : best & 0xFFFFu;
        const unsigned score = (b >> 4) - 2047u;
        if ((b >> 4) <= 2047u) return 0u;
        const unsigned pos = ((h ? rowB : rowA) << 10) | (unsigned)(j0 + 15 - (int)(b & 15u) + 1);
        return (score << 19) | (0x7FFFFu - pos);
    }
};

// pass 2: the reference's walk back from the first maximum (variant.c:1405-1417) on the stored directions of one
// segment of lanes.  `dirs`: plane 0 of the segment's lane 0 at step 0; `plane`: words from one plane to the next;
// `stride`: words from one LANE to the next -- a lane's steps are contiguous, so the cells of a diagonal (same lane,
// one step back per cell) share 32-byte sectors.  Cell (i, j) of half h: lane (j-1) / cpl at step i - 1 + lane,
// column c = (j-1) % cpl of that lane.
// Every word is a dependent load from a buffer far larger than L2.  One word holds a whole row of a lane's columns,
// so the words of the next kPackAhead ROWS -- in the lanes the current diagonal passes through -- are fetched
// together and serve substitutions, insertions (one row up, same word as the diagonal's) and deletions (same row,
// same word) alike until the path leaves those lanes.  Substitutions are only counted as runs of a diagonal; their
// bases are compared (raw bytes, variant.c:1412) when the run ends, with loads that do not depend on each other.
enum { kPackAhead = 8 };
__host__ __device__ __forceinline__ int pack_count_mismatches(const uint8_t* __restrict__ t1, const uint8_t* __restrict__ t2, int i, int j, int run)
{
    int ns = 0;                                            // the run's cells: (i + k, j + k), k = 1 .. run
    for (int k = 1; k <= run; k++) ns += t1[j + k - 1] != t2[i + k - 1];
    return ns;
}

__host__ __device__ __forceinline__ void pack_walk_back(const uint32_t* __restrict__ dirs, size_t plane, int cpl, int stride, int h, unsigned pos,
                                                        const uint8_t* __restrict__ t1, const uint8_t* __restrict__ t2,
                                                        int& subs, int& indels, int& aligned)
{
    int i = (int)(pos >> 10), j = (int)(pos & 1023u);
    int ns = 0, ni = 0, na = 1;                            // the NUL column (:1405)
    const size_t pplane = (size_t)(pack_planes(cpl) - 1) * plane;
    int lj = (j - 1) / cpl, c = (j - 1) - lj * cpl;        // kept incrementally: no division inside the loop
    int run = 0;                                           // substitutions since the last gap, not compared yet
    bool walking = true, flush = false;
    while (walking && i > 0 && j > 0) {                    // row 0 and column 0 hold negative scores
        if (flush) { ns += pack_count_mismatches(t1, t2, i, j, run); run = 0; flush = false; }
        unsigned wp[kPackAhead], w0[kPackAhead], w1[kPackAhead];
        int lane_of[kPackAhead];
        {
            int ljt = lj, ct = c;
#pragma unroll
            for (int t = 0; t < kPackAhead; t++) {
                lane_of[t] = ljt;
                wp[t] = w0[t] = w1[t] = 0;
                if (i - t > 0 && ljt >= 0) {
                    const uint32_t* w = dirs + (size_t)ljt * stride + (i - t - 1 + ljt);
                    wp[t] = w[pplane];
                    w0[t] = w[0];
                    if (cpl > 8) w1[t] = w[plane];
                }
                if (--ct < 0) { ct = cpl - 1; ljt--; }
            }
        }
        bool on = true;                                    // still inside the words that were fetched
#pragma unroll
        for (int t = 0; t < kPackAhead; t++) {             // row i0 - t
            while (on) {                                   // its cells, leftwards over deletions
                if (i <= 0 || j <= 0) { walking = false; on = false; break; }
                if (lane_of[t] != lj) { on = false; break; }
                if (!((wp[t] >> ((((c & 1) << 1) + h) * 8 + (c >> 1))) & 1u)) { walking = false; on = false; break; }   // V <= 0
                const int cnt = (c >> 3) ? cpl - 8 : (cpl < 8 ? cpl : 8);
                const unsigned code = (((c >> 3) ? w1[t] : w0[t]) >> (16 * h + 2 * (cnt - 1 - (c & 7)))) & 3u;
                if (code != 2 && run > 0) { flush = true; on = false; break; }          // the run of substitutions ends here
                if (code != 1) { j--; if (--c < 0) { c = cpl - 1; lj--; } }           // one column to the left
                if (code == 0) { ni++; continue; }                                       // deletion: same row
                if (code == 1) ni++; else run++;
                na++; i--;
                break;                                                                   // one row up: the next word
            }
        }
    }
    ns += pack_count_mismatches(t1, t2, i, j, run);
    subs = ns; indels = ni; aligned = na;
}

#ifdef __CUDACC__
// The work of one launch: `n` pairs of `order` (sorted by target length), 2 * (32 / SEG) of them per warp pass.
// Pass g = pairs [g * per_pass, (g + 1) * per_pass) writes its direction words to dirs + g * steps_pad * 32 * PLANES
// (plane, lane, step; steps_pad = steps_cap rounded up to 8),
// the columns per lane it used to pass_cpl[g], and the first maximum of every pair (score << 19 | ~position, 0:
// no positive score) to best[]; indel_support_walk_kernel turns those into the three counts.
struct PackArgs {
    int n;
    const int32_t* order;
    const uint8_t* targets; const int64_t* target_off;
    const uint8_t* queries; const int64_t* query_off;
    int32_t* subs; int32_t* indels; int32_t* aligned;
    uint32_t* dirs; int32_t* pass_cpl; uint32_t* best;
    int steps_cap;         // wavefront steps a pass's region holds
    int seg, maxcpl;       // the walk kernel's copy of the pass kernel's template arguments
    unsigned one;          // 1 (pk_mad)
    int* error_flag;
};

template <int CPL, int SEG>
__device__ __forceinline__ void support_pack_pass(const uint8_t* t1a, int len1a, const uint8_t* t2a, int len2a,
                                                  const uint8_t* t1b, int len1b, const uint8_t* t2b, int len2b,
                                                  int lane, uint32_t* __restrict__ dirs, size_t plane, int steps_pad, uint32_t* stage /* this lane's column of the warp's 3 x 8 x 32 words */,
                                                  unsigned one, unsigned& keya, unsigned& keyb)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int li = lane & (SEG - 1), j0 = li * CPL;
    PackLane<CPL> L;
    L.init(t1a, len1a, t1b, len1b, j0);
    const int len1 = max(len1a, len1b), len2 = max(len2a, len2b);
    const int nl = (len1 + CPL - 1) / CPL;
    const int steps = (len1 > 0 && len2 > 0) ? len2 + nl - 1 : 0;
    const int wsteps = __reduce_max_sync(FULL, steps);     // the segments of a warp step together
    // query bases, SEG rows per load and one load AHEAD of their use (the load's latency would otherwise stall the warp
    // every SEG steps): qnext holds rows [s0 + SEG, s0 + 2 SEG) while the wavefront consumes qcur and qprev
    auto load_q = [&](int r) {
        const unsigned qa = r < len2a ? (unsigned)up_case(t2a[r]) << 4 : (unsigned)kPkPastQuery;
        const unsigned qb = r < len2b ? (unsigned)up_case(t2b[r]) << 4 : (unsigned)kPkPastQuery;
        return pk2(0u - qa, 0u - qb);
    };
    unsigned qcur = 0, qprev = 0, qnext = load_q(li);
    uint32_t* sp = stage;                                  // this step's slot of the staging tile
#pragma unroll 1
    for (int s = 0; s < wsteps; s++) {
        const int sl = s & (SEG - 1);
        if (sl == 0) { qprev = qcur; qcur = qnext; qnext = load_q(s + SEG + li); }
        // lane li works on query index s - li: lane (s - li) mod SEG holds it, in qcur if it was loaded this round
        const unsigned give = li <= sl ? qcur : qprev;
        const unsigned nb = __shfl_sync(FULL, give, (s - li) & (SEG - 1), SEG);
        const unsigned inV = __shfl_up_sync(FULL, L.outV, 1, SEG);
        const unsigned inE = __shfl_up_sync(FULL, L.outE, 1, SEG);
        const int i = s - li + 1;
        if (i >= 1 && i <= len2 && li < nl) {
            unsigned w[3];
            L.row(i, j0, li == 0, inV, inE, nb, one, w);
#pragma unroll
            for (int p = 0; p < pack_planes(CPL); p++) sp[p * 256] = w[p];
        }
        sp += 32;
        if ((s & 7) == 7 || s == wsteps - 1) {             // eight steps of this lane = one 32-byte sector per plane
            sp = stage;
#pragma unroll
            for (int p = 0; p < pack_planes(CPL); p++) {
                uint4 lo, hi;
                lo.x = stage[(p * 8 + 0) * 32]; lo.y = stage[(p * 8 + 1) * 32]; lo.z = stage[(p * 8 + 2) * 32]; lo.w = stage[(p * 8 + 3) * 32];
                hi.x = stage[(p * 8 + 4) * 32]; hi.y = stage[(p * 8 + 5) * 32]; hi.z = stage[(p * 8 + 6) * 32]; hi.w = stage[(p * 8 + 7) * 32];
                uint4* out = reinterpret_cast<uint4*>(dirs + p * plane + (size_t)lane * steps_pad + (s & ~7));
                __stcs(out, lo); __stcs(out + 1, hi);      // streaming: read once, by another kernel
            }
        }
    }
    // first maximum in row-major order over the lanes of the segment, per half
    unsigned ka = L.best_key(0, j0), kb = L.best_key(1, j0);
#pragma unroll
    for (int o = SEG / 2; o > 0; o >>= 1) {
        ka = max(ka, __shfl_xor_sync(FULL, ka, o));
        kb = max(kb, __shfl_xor_sync(FULL, kb, o));
    }
    keya = ka; keyb = kb;
}

template <int SEG, int MAXCPL>
__global__ void __launch_bounds__(128, 4)
indel_support_pack_kernel(const __grid_constant__ PackArgs a)
{
    constexpr int SLOTS = 32 / SEG;                        // register halves come in pairs: 2 * SLOTS pairs per pass
    constexpr int PLANES = MAXCPL <= 8 ? 2 : 3;
    const int lane = threadIdx.x & 31;
    const int seg = lane / SEG, li = lane & (SEG - 1);
    const int steps_pad = (a.steps_cap + 7) & ~7;
    const size_t plane = (size_t)steps_pad * 32;
    __shared__ uint32_t stage_all[4][3 * 8 * 32];
    uint32_t* const stage = stage_all[threadIdx.x >> 5] + lane;
    const long long gwarp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long npass = ((long long)a.n + 2 * SLOTS - 1) / (2 * SLOTS);
    for (long long g = gwarp; g < npass; g += nwarps) {
        const long long ka = g * (2 * SLOTS) + 2 * seg, kb = ka + 1;
        const bool havea = ka < a.n, haveb = kb < a.n;
        const long long ia = havea ? a.order[ka] : 0, ib = haveb ? a.order[kb] : 0;
        const uint8_t* t1a = a.targets; const uint8_t* t2a = a.queries; const uint8_t* t1b = a.targets; const uint8_t* t2b = a.queries;
        int len1a = 0, len2a = 0, len1b = 0, len2b = 0;
        if (havea) {
            const int64_t to = a.target_off[ia], qo = a.query_off[ia];
            t1a += to; t2a += qo; len1a = (int)(a.target_off[ia + 1] - to); len2a = (int)(a.query_off[ia + 1] - qo);
        }
        if (haveb) {
            const int64_t to = a.target_off[ib], qo = a.query_off[ib];
            t1b += to; t2b += qo; len1b = (int)(a.target_off[ib + 1] - to); len2b = (int)(a.query_off[ib + 1] - qo);
        }
        const int wl1 = __reduce_max_sync(0xFFFFFFFFu, max(len1a, len1b));
        const int wl2 = __reduce_max_sync(0xFFFFFFFFu, max(len2a, len2b));
        const int need = (wl1 + SEG - 1) / SEG;            // columns per lane, the warp's widest target
        const int cpl = max(2, (need + 1) & ~1);
        unsigned keya = 0, keyb = 0;
        if (need > MAXCPL || wl2 + SEG - 1 > a.steps_cap) {                           // the host's classes rule this out
            if (lane == 0) atomicExch(a.error_flag, 1);
        } else {
            uint32_t* const dirs = a.dirs + (size_t)g * plane * PLANES;
#define PACK(C) support_pack_pass<C, SEG>(t1a, len1a, t2a, len2a, t1b, len1b, t2b, len2b, lane, dirs, plane, steps_pad, stage, a.one, keya, keyb)
            switch (cpl) {
                case 2: PACK(2); break;
                case 4: PACK(4); break;
                case 6: PACK(6); break;
                case 8: PACK(8); break;
                case 10: if (MAXCPL > 8) PACK(10); break;
                case 12: if (MAXCPL > 8) PACK(12); break;
                case 14: if (MAXCPL > 8) PACK(14); break;
                default: if (MAXCPL > 8) PACK(16); break;
            }
#undef PACK
        }
        if (lane == 0) a.pass_cpl[g] = cpl;
        if (li == 0 && havea) a.best[ka] = keya;
        if (li == 1 && haveb) a.best[kb] = keyb;
    }
}

// pass 2, one THREAD per pair: a warp walks 32 paths at once instead of leaving 30 lanes idle behind the wavefront.
// One launch serves the pass launches that are waiting for it (blocks [first_block[q], first_block[q + 1]) walk job q):
// the kernel is a chain of dependent loads per thread, so its duration hardly depends on the number of pairs.
enum { kWalkJobs = 3 };
struct WalkArgs { int njobs; int first_block[kWalkJobs + 1]; PackArgs job[kWalkJobs]; };

__global__ void __launch_bounds__(128, 8)
indel_support_walk_kernel(const __grid_constant__ WalkArgs wa)
{
    int q = 0;
    while (q + 1 < wa.njobs && (int)blockIdx.x >= wa.first_block[q + 1]) q++;
    const PackArgs& a = wa.job[q];
    const long long k = (blockIdx.x - wa.first_block[q]) * (long long)blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int per_pass = 2 * (32 / a.seg);
    const int steps_pad = (a.steps_cap + 7) & ~7;
    const size_t plane = (size_t)steps_pad * 32;
    const long long g = k / per_pass;
    const int r = (int)(k - g * per_pass), seg = r >> 1, h = r & 1;
    const long long idx = a.order[k];
    const unsigned top = a.best[k];
    int subs = 0, indels = 0, aligned = 1;
    if (top)
        pack_walk_back(a.dirs + (size_t)g * plane * (a.maxcpl <= 8 ? 2 : 3) + (size_t)seg * a.seg * steps_pad, plane, a.pass_cpl[g], steps_pad, h,
                       0x7FFFFu - (top & 0x7FFFFu), a.targets + a.target_off[idx], a.queries + a.query_off[idx], subs, indels, aligned);
    a.subs[idx] = subs; a.indels[idx] = indels; a.aligned[idx] = aligned;
}

// direction words of one launch
static inline size_t pack_dirs_bytes(long long npairs, int seg, int maxcpl, int steps_cap)
{
    const long long per_pass = 2 * (32 / seg);
    return (size_t)((npairs + per_pass - 1) / per_pass) * (size_t)((steps_cap + 7) & ~7) * 32 * 4 * (maxcpl <= 8 ? 2 : 3);
}
#endif  // __CUDACC__

}  // namespace indelgpu
