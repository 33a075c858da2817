// Row f1 of SURVEY.md section 8: realign_with_indel (src/variant.c:1246-1424), the full-matrix affine DP
// behind is_indel_supported (variant.c:1561-1573, annotate mode): would this read support a known indel?
//
// Two kernels.
//  * indel_support_wave_kernel -- the one that runs for real reads (target <= 512, query <= 500 bases).
//    A segment of 8, 16 or 32 lanes per (target, query) pair (4, 2 or 1 pairs per warp, by target length).
//    Lane l of the segment owns CPL consecutive target columns and sweeps the rows as
//    a wavefront (lane l works on row s - l + 1 at step s); its slice of the previous row, of F and of the
//    path counters lives in registers, the row boundary travels to the right-hand lane with two shuffles
//    per step.  The reference fills a score matrix and a direction matrix and then walks back from the
//    first maximum while the score stays positive, counting substitutions, gap columns and aligned
//    columns.  That walk follows ONE predecessor per cell (the direction matrix of V only), so the three
//    counters of "the path that ends here" are carried forward with the cell instead:
//        C[i][j] = V[i][j] > 0 ? step(i, j) + C[pred(i, j)] : 0,
//    and the answer is C at the first maximum.  No matrix is stored and nothing is traced back: the kernel
//    reads each sequence once and is bound by the integer pipe.
//  * indel_support_kernel -- any length up to 8000: one pair per THREAD, score (16 bit) and direction
//    matrices in lane-interleaved global scratch, explicit traceback.  Used for the pairs the wavefront
//    kernel's packing cannot hold; also the plain statement of the algorithm the tests compare against.
//
// Reproduced quirks: E restarts at 0 on every row (variant.c:1322) and F starts at 0 (:1305); the
// comparison is case-insensitive in the DP (:1325) but case-sensitive when substitutions are counted
// (:1412); the counting loop starts on the terminating NUL, so `aligned` is one more than the number of
// aligned columns (:1405-1417).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace indelgpu {

struct SupportArgs {
    int n;
    const int32_t* index;  // optional: task idx = index[k], k < n (the pairs left to the thread-per-pair kernel)
    const uint8_t* targets; const int64_t* target_off;     // target j = targets[target_off[j] .. target_off[j+1])
    const uint8_t* queries; const int64_t* query_off;
    int32_t* subs; int32_t* indels; int32_t* aligned;
    int16_t* V;            // per warp: 32 * cells_cap int16, lane-interleaved
    int8_t*  I;            // per warp: 32 * cells_cap
    int32_t* F;            // per warp: 32 * (max_len1 + 1)
    long long cells_cap;   // (max_len1 + 1) * (max_len2 + 1)
    int max_len1, max_len2;
    unsigned long long* cell_total;
    int* error_flag;
};

__host__ __device__ __forceinline__ int up_case(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

__global__ void __launch_bounds__(128)
indel_support_kernel(const __grid_constant__ SupportArgs a)
{
    enum { SUB = 0, INS = 1, DEL = 2 };
    const int match = 2, mismatch = 1, gopen = 4, gextend = 1;     // variant.c:1289-1292
    const int lane = threadIdx.x & 31;
    const long long gwarp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    int16_t* V = a.V + gwarp * 32 * a.cells_cap + lane;
    int8_t*  I = a.I + gwarp * 32 * a.cells_cap + lane;
    int32_t* F = a.F + gwarp * 32 * (a.max_len1 + 1) + lane;
    unsigned long long cells = 0;
#define AT(p, e) (p)[(long long)(e) * 32]
    for (long long k = gwarp * 32 + lane; k < a.n; k += nwarps * 32) {
        const long long idx = a.index ? a.index[k] : k;
        const uint8_t* t1 = a.targets + a.target_off[idx];
        const uint8_t* t2 = a.queries + a.query_off[idx];
        const int len1 = (int)(a.target_off[idx + 1] - a.target_off[idx]);
        const int len2 = (int)(a.query_off[idx + 1] - a.query_off[idx]);
        if (len1 < 0 || len2 < 0 || len1 > a.max_len1 || len2 > a.max_len2) {
            a.subs[idx] = a.indels[idx] = a.aligned[idx] = 0;
            atomicExch(a.error_flag, 1);
            continue;
        }
        const int W = len1 + 1;
        for (int j = 0; j <= len1; j++) { AT(V, j) = (int16_t)(-gopen - j * gextend); AT(F, j) = 0; }
        int max_score = 0, max_i = -1, max_j = -1;
        for (int i = 1; i <= len2; i++) {
            const int b = up_case(t2[i - 1]);
            const long long row = (long long)i * W, prow = row - W;
            int left = -gopen - i * gextend;                       // V[i][0]
            AT(V, row) = (int16_t)left;
            int diag = AT(V, prow);                                // V[i-1][0]
            int E = 0;                                             // restarts on every row
            for (int j = 1; j <= len1; j++) {
                const int up = AT(V, prow + j);
                const int ifsub = diag + (up_case(t1[j - 1]) == b ? match : -mismatch);
                const int f = AT(F, j);
                const int ifins = max(f, up - gopen) - gextend;
                AT(F, j) = ifins;
                const int ifdel = max(E, left - gopen) - gextend;
                E = ifdel;
                const int ifindel = max(ifins, ifdel);
                int v = ifsub, d = SUB;
                if (v < ifindel) { d = (ifins >= ifdel) ? INS : DEL; v = ifindel; }     // :1336-1342
                AT(V, row + j) = (int16_t)v; AT(I, row + j) = (int8_t)d;
                if (v > max_score) { max_score = v; max_i = i; max_j = j; }             // strict: first maximum
                diag = up; left = v;
            }
        }
        cells += (unsigned long long)len1 * (unsigned long long)len2;
        int subs = 0, ins = 0, dels = 0, aligned = 1;              // the NUL column (:1405)
        int score = max_score, i = max_i, j = max_j;
        while (score > 0) {
            const int d = AT(I, (long long)i * W + j);
            if (d == SUB) { if (t1[j - 1] != t2[i - 1]) subs++; aligned++; i--; j--; }
            else if (d == INS) { ins++; aligned++; i--; }
            else { dels++; j--; }
            score = AT(V, (long long)i * W + j);
        }
        a.subs[idx] = subs; a.indels[idx] = ins + dels; a.aligned[idx] = aligned;
    }
#undef AT
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_xor_sync(0xFFFFFFFFu, cells, o);
    if (lane == 0 && cells) atomicAdd(a.cell_total, cells);
}

// ---------------------------------------------------------------------------------------------------
// wavefront kernel
// ---------------------------------------------------------------------------------------------------
enum { kWaveMaxTarget = 512, kWaveMaxQuery = 500 };
// path counters packed in one word: aligned columns (<= 500) | substitutions (<= 500) | gap columns (<= 1012)
enum { kAln1 = 1, kSub1 = 1 << 9, kInd1 = 1 << 18 };

// One wavefront of SEG lanes (8, 16 or 32; a warp runs 32 / SEG pairs side by side, each on its own
// segment of lanes): a short target on all 32 lanes would leave most of them without a column and pay
// 31 steps of pipeline fill for nothing.
template <int CPL, int SEG>
__device__ __forceinline__ void support_wavefront(const uint8_t* __restrict__ t1, int len1,
                                                  const uint8_t* __restrict__ t2, int len2, int lane,
                                                  int& best_out, int& pos_out, int& c_out)
{
    const unsigned FULL = 0xFFFFFFFFu;
    int araw[CPL], aup[CPL], V[CPL], F[CPL], Cn[CPL];
    const int li = lane & (SEG - 1);                   // lane within the segment
    const int j0 = li * CPL;                           // this lane owns columns j0+1 .. j0+CPL
#pragma unroll
    for (int c = 0; c < CPL; c++) {
        const int j = j0 + c;
        const int ch = j < len1 ? t1[j] : 0;
        araw[c] = ch;
        aup[c] = j < len1 ? up_case(ch) : 0x100;       // past the target: matches nothing, so never a maximum
        V[c] = -4 - (j + 1);                           // row 0 (variant.c:1303-1306)
        F[c] = 0; Cn[c] = 0;
    }
    int best = 0, bestpos = 0, bestC = 0;
    int vdiag = -4 - j0, cdiag = 0;                    // V and C of (row - 1, j0) for the row this lane does next
    unsigned out_pk = 0; int out_c = 0;                // boundary handed to lane + 1: V:12 | E:12 | query base:8, and C
    const int nl = (len1 + CPL - 1) / CPL;             // lanes of the segment that own a column
    const int steps = (len1 > 0 && len2 > 0) ? len2 + nl - 1 : 0;
    const int wsteps = __reduce_max_sync(FULL, steps); // the segments of a warp step together
    unsigned qreg = 0;
#pragma unroll 1
    for (int s = 0; s < wsteps; s++) {
        if ((s & (SEG - 1)) == 0) qreg = (s + li < len2) ? t2[s + li] : 0;
        const unsigned b0 = __shfl_sync(FULL, qreg, s & (SEG - 1), SEG);
        const unsigned in_pk = __shfl_up_sync(FULL, out_pk, 1, SEG);
        const int in_c = __shfl_up_sync(FULL, out_c, 1, SEG);
        const int i = s - li + 1;
        if (i >= 1 && i <= len2 && li < nl) {
            int left, E, cleft; unsigned braw;
            if (li == 0) { left = -4 - i; E = 0; cleft = 0; braw = b0; }            // V[i][0]; E restarts per row
            else { left = (int)in_pk >> 20; E = (int)(in_pk << 12) >> 20; braw = in_pk & 0xFFu; cleft = in_c; }
            const int bup = up_case((int)braw);
            int diag = vdiag, cd = cdiag;
            vdiag = left; cdiag = cleft;
#pragma unroll
            for (int c = 0; c < CPL; c++) {
                const int up = V[c], cup = Cn[c];
                const int ifsub = diag + (aup[c] == bup ? 2 : -1);
                const int ifins = max(F[c], up - 4) - 1;
                F[c] = ifins;
                const int ifdel = max(E, left - 4) - 1;
                E = ifdel;
                const int ifindel = max(ifins, ifdel);
                const int cgap = (ifins >= ifdel) ? cup + (kInd1 + kAln1) : cleft + kInd1;     // variant.c:1339
                const int csub = cd + kAln1 + (araw[c] != (int)braw ? kSub1 : 0);
                int v = ifsub, cc = csub;
                if (ifsub < ifindel) { v = ifindel; cc = cgap; }                                 // :1336
                cc = v > 0 ? cc : 0;                       // the walk back stops where the score is not positive
                if (v > best) { best = v; bestC = cc; bestpos = (i << 10) | (j0 + c + 1); }      // first maximum
                V[c] = v; Cn[c] = cc;
                diag = up; cd = cup; left = v; cleft = cc;
            }
            out_pk = ((unsigned)left << 20) | (((unsigned)E & 0xFFFu) << 8) | braw;
            out_c = cleft;
        }
    }
    best_out = best; pos_out = bestpos; c_out = bestC;
}

struct WaveArgs {
    int n;                 // pairs in `order`
    const int32_t* order;  // pair indices of this launch, sorted by target length (neighbours share a tile width)
    const uint8_t* targets; const int64_t* target_off;
    const uint8_t* queries; const int64_t* query_off;
    int32_t* subs; int32_t* indels; int32_t* aligned;
};

template <int SEG>
__global__ void __launch_bounds__(128, 3)
indel_support_wave_kernel(const __grid_constant__ WaveArgs a)
{
    constexpr int GROUP = 32 / SEG;                    // pairs per warp and pass
    const int lane = threadIdx.x & 31;
    const int seg = lane / SEG;
    const long long gwarp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = gwarp * GROUP; base < a.n; base += nwarps * GROUP) {
        const long long k = base + seg;
        const bool have = k < a.n;
        const long long idx = have ? a.order[k] : 0;
        const uint8_t* t1 = a.targets; const uint8_t* t2 = a.queries;
        int len1 = 0, len2 = 0;
        if (have) {
            const int64_t to = a.target_off[idx], qo = a.query_off[idx];
            t1 += to; t2 += qo;
            len1 = (int)(a.target_off[idx + 1] - to);
            len2 = (int)(a.query_off[idx + 1] - qo);
        }
        int best, pos, cnt;
        const int need = (__reduce_max_sync(0xFFFFFFFFu, len1) + SEG - 1) / SEG;       // columns per lane, the warp's widest target
#define WAVE(C) support_wavefront<C, SEG>(t1, len1, t2, len2, lane, best, pos, cnt)
        switch (need) {
            case 0: case 1: case 2: WAVE(2); break;
            case 3: case 4: WAVE(4); break;
            case 5: case 6: WAVE(6); break;
            case 7: case 8: WAVE(8); break;
            case 9: case 10: WAVE(10); break;
            case 11: case 12: WAVE(12); break;
            case 13: case 14: WAVE(14); break;
            default: WAVE(16); break;
        }
#undef WAVE
        // first maximum in row-major order over the lanes of the segment: largest score, then smallest (i, j)
        const unsigned key = best > 0 ? ((unsigned)best << 19) | (0x7FFFFu - (unsigned)pos) : 0u;
        unsigned top = key;
#pragma unroll
        for (int o = SEG / 2; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xFFFFFFFFu, top, o));
        const unsigned segmask = (SEG == 32 ? 0xFFFFFFFFu : ((1u << SEG) - 1u)) << (seg * SEG);
        const int owner = __ffs(__ballot_sync(0xFFFFFFFFu, key == top) & segmask) - 1;
        const int c = __shfl_sync(0xFFFFFFFFu, cnt, owner);
        if ((lane & (SEG - 1)) == 0 && have) {
            const int cc = top ? c : 0;
            a.subs[idx] = (cc >> 9) & 0x1FF;
            a.indels[idx] = (cc >> 18) & 0x3FF;
            a.aligned[idx] = (cc & 0x1FF) + 1;         // the NUL column (variant.c:1405)
        }
    }
}

}  // namespace indelgpu
