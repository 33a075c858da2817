// Realignment with bands wider than one diagonal (-g N > 0) as a pipeline of kernels.
//
// The k-mer vote wants a warp per read; the banded DP wants 32 independent alignments per warp (one
// per thread, band_dp.cuh) because one anti-diagonal of a (g+1)-wide band holds too few independent
// cells to feed a warp.  So the two rounds of attempt_diagonal_alignments (alignment.c:539-759) are
// cut at the alignment calls, with a few dozen bytes per read of intermediate state in HBM:
//
//   pipe_vote_kernel(round 0)   warp / read    find_best_band on window 1                    -> low[0]
//   pipe_dp_kernel(round 0)     thread / read  local_align + ALIGN + fetch_cigar             -> aln[0], cig[0]
//   pipe_vote_kernel(round 1)   warp / read    plan (:568-717) + find_best_band on window 2  -> plan, low[1]
//   pipe_dp_kernel(round 1)     thread / read                                               -> aln[1], cig[1]
//   pipe_combine_kernel         warp / read    junction, update_readsegs, results
#pragma once

#include "realign_kernel.cuh"

namespace indelgpu {

enum { PF_VOTE1_OK = 1, PF_GO = 2, PF_VOTE2_OK = 4 };

struct PipeBufs {
    int32_t* low;         // [2][n]
    Aln* aln;             // [2][n]
    uint32_t* cig;        // [2][n * cig_stride]
    Plan* plan;           // [n]
    int32_t* flags;       // [n]  PF_*
    int cig_stride;
};

// ---- warp per read: (plan +) vote ------------------------------------------------------------
// what one read needs from the vote of `round`; staged one read ahead like in the fused kernel
struct VoteJob {
    ReadCtx c;
    uint32_t zs1, e1, zs2, e2, anc;
    int flags;
    bool go;                  // false: nothing to stage or vote for this read in this round
};

template <bool DIRECT, int HB>
__global__ void __launch_bounds__(256)
pipe_vote_kernel(const __grid_constant__ RealignArgs a, const PipeBufs p, const int round)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpView V;
    bind_warp(V, smem + (size_t)warp * a.L.total, a.L);
    init_warp_tables(V);
    if (lane == 0) { mbar_init(V.bar + 0, 1); mbar_init(V.bar + 1, 1); mbar_fence_init(); }
    __syncwarp();
    uint32_t phase = 0;                                           // bit b = parity of buffer b's barrier
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    Plan* s_plan = reinterpret_cast<Plan*>(V.misc);

    // derive the job of read idx (round 1: plan first, alignment.c:568-717) and start its copies
    auto prepare = [&](int idx, int buf) -> VoteJob {
        VoteJob j;
        j.go = false; j.flags = 0;
        j.zs1 = j.e1 = j.zs2 = j.e2 = j.anc = 0;
        if (idx >= a.n) { j.c.bad = true; return j; }
        j.c = load_read_ctx(a, idx);
        if (j.c.bad) { if (round == 0 && lane == 0) p.flags[idx] = 0; return j; }
        j.zs1 = (uint32_t)j.c.left1; j.e1 = (uint32_t)j.c.right1; j.zs2 = 0; j.e2 = (uint32_t)j.c.readlen; j.anc = (uint32_t)j.c.position;
        j.go = true;
        if (round == 1) {
            j.flags = p.flags[idx];
            if (!(j.flags & PF_VOTE1_OK)) { j.go = false; return j; }
            __syncwarp();
            if (lane == 0) {
                const Aln a1 = p.aln[idx];
                make_plan(a.P, a1, p.cig + (int64_t)idx * p.cig_stride, j.c.position, j.c.left2, j.c.right2, (unsigned)j.c.readlen, s_plan);
                p.plan[idx] = *s_plan;
            }
            __syncwarp();
            j.go = s_plan->go != 0;
            j.zs1 = s_plan->zs1; j.e1 = s_plan->e1; j.zs2 = s_plan->zs2; j.e2 = s_plan->e2; j.anc = s_plan->anc;
            __syncwarp();
        }
        if (j.go && lane == 0) stage_read(a, V, j.c, buf);
        return j;
    };

    int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int it = 0;
    VoteJob cur = prepare(idx, 0);
#pragma unroll 1
    for (; idx < a.n; idx += nwarps, it++) {
        const int buf = it & 1;
        const VoteJob nxt = prepare(idx + nwarps, buf ^ 1);
        if (cur.go) {
            const bool landed = mbar_wait(V.bar + buf, (phase >> buf) & 1u);
            phase ^= 1u << buf;
            if (!landed) { if (lane == 0) atomicExch(a.error_flag, 3); break; }
            const ReadCtx& c = cur.c;
            const uint8_t* read = read_buf(V, buf) + (int)(c.roff & 15);
            pack_read_warp(V, read, c.readlen);
            const int64_t sw0 = ((c.cbase + c.left2) & ~(int64_t)63) >> 4;
            bool ok;
            const int low = vote_band_warp<DIRECT, HB>(a.P, V, win_buf(V, buf), sw0, c.cbase + cur.zs1, (int)(cur.e1 - cur.zs1),
                                                       (int)cur.zs2, (int)(cur.e2 - cur.zs2), (int)(cur.anc - cur.zs1), &ok);
            if (lane == 0) {
                p.low[(int64_t)round * a.n + idx] = low;
                if (round == 0) p.flags[idx] = ok ? PF_VOTE1_OK : 0;
                else p.flags[idx] = cur.flags | PF_GO | (ok ? PF_VOTE2_OK : 0);
                if (!ok) atomicExch(a.error_flag, 1);         // numdiagonals <= numgaps: the reference aborts (alignment.c:405)
            }
            __syncwarp();
        }
        cur = nxt;
    }
}

// ---- thread per read: attempt_band_alignment (alignment.c:343-391) ----------------------------
// which window / read slice an alignment of `round` uses; false when the read takes no part in it
__device__ __forceinline__ bool pipe_task(const RealignArgs& a, const PipeBufs& p, int round, int idx, ReadCtx* c,
                                          uint32_t* zs1, uint32_t* e1, uint32_t* zs2, uint32_t* e2, int* low)
{
    *c = load_read_ctx(a, idx);
    if (c->bad) return false;
    const int flags = p.flags[idx];
    if (!(round == 0 ? (flags & PF_VOTE1_OK) : (flags & PF_VOTE2_OK))) return false;
    *low = p.low[(int64_t)round * a.n + idx];
    if (round == 0) { *zs1 = (uint32_t)c->left1; *e1 = (uint32_t)c->right1; *zs2 = 0; *e2 = (uint32_t)c->readlen; }
    else { const Plan pl = p.plan[idx]; *zs1 = pl.zs1; *e1 = pl.e1; *zs2 = pl.zs2; *e2 = pl.e2; }
    return true;
}

// sweeps of local_align + the unique-diagonal shortcut for every read; the alignments that need ALIGN's
// divide and conquer run it 32 at a time (banded_two_phase_loop; see band_tasks_kernel for why)
__global__ void __launch_bounds__(128, BAND_MIN_BLOCKS)
pipe_dp_kernel(const __grid_constant__ RealignArgs a, const PipeBufs p, const int round)
{
    __shared__ DcTask s_pend[4][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int wb4 = 4 * (a.scratch.max_band + 4);
    const IArr<32> gbase{a.scratch.base + (long long)gwarp * 32 * a.scratch.stride + lane};
    const IArr<32> bands = gbase;
    const IArr<32> rowsb = gbase + wb4;
    DcFrame st[kDcFrames];

    auto geometry = [&](int idx, const uint8_t** read, int* M, const uint8_t** win, int* N, int* lo, int* hi) -> bool {
        ReadCtx c; uint32_t zs1 = 0, e1 = 0, zs2 = 0, e2 = 0; int low = 0;
        const bool active = pipe_task(a, p, round, idx, &c, &zs1, &e1, &zs2, &e2, &low);
        *N = (int)(e1 - zs1); *M = (int)(e2 - zs2);
        *lo = max(-*M, low); *hi = min(*N, low + a.P.g);                   // localalign.c:70-71
        *read = a.reads + c.roff + zs2; *win = a.ref.raw + c.cbase + zs1;
        return active;
    };
    auto phase0 = [&](int idx, bool valid, BandLocal& L) -> bool {
        bool mine = false;
        if (valid && a.P.g + 1 >= kWarpBandMin) {
            const uint8_t* read; const uint8_t* win; int M, N, lo, hi;
            const bool active = geometry(idx, &read, &M, &win, &N, &lo, &hi);
            const int band = hi - lo + 1;
            mine = active && band >= kWarpBandMin && band <= kWarpBandMax && a.P.G >= 0 && a.P.H >= 0 &&
                   2 * band <= a.scratch.max_band && M <= a.scratch.max_rows && 2 * M + band + 4 <= p.cig_stride;
        }
        if (a.P.g + 1 >= kWarpBandMin)                                      // warp-uniform: narrow-band runs skip the pass
            warp_serve_wide_bands(a.P, idx, mine, geometry, L);
        return mine;
    };
    auto phase1 = [&](int idx, DcTask& t, const BandLocal* pre) -> bool {
        ReadCtx c; uint32_t zs1 = 0, e1 = 0, zs2 = 0, e2 = 0; int low = 0;
        bool active = pipe_task(a, p, round, idx, &c, &zs1, &e1, &zs2, &e2, &low);
        const int N = (int)(e1 - zs1), M = (int)(e2 - zs2);
        const int lo = max(-M, low), hi = min(N, low + a.P.g);            // localalign.c:70-71
        const int band = hi - lo + 1;
        if (active && (band < 1 || 2 * band > a.scratch.max_band || M > a.scratch.max_rows || 2 * M + band + 4 > p.cig_stride)) {
            atomicExch(a.error_flag, 1);                                  // cannot happen for N, M >= 1; kept as a guard
            active = false;
        }
        Aln r;
        memset(&r, 0, sizeof(r));
        bool need = false;
        if (active) {
            const uint8_t* read = a.reads + c.roff + zs2;
            const uint8_t* win = a.ref.raw + c.cbase + zs1;
            const BandLocal L = pre ? *pre : band_local<32>(a.P, bands, a.scratch.max_band, read, M, win, N, lo, hi);
            r.low = low; r.up = low + a.P.g;
            r.cells_fwd = L.cf; r.cells_rev = L.cr;
            if (!L.none) {                                                // :385-388
                r.score = L.best;
                r.q1 = L.starti + (int)zs2 - 1; r.r1 = L.startj + (int)zs1 - 1;
                r.q2 = L.endi + (int)zs2;       r.r2 = L.endj + (int)zs1;
                uint32_t* cig = p.cig + ((int64_t)round * a.n + idx) * p.cig_stride;
                int n = 0, cells = 0;
                if (band_unique_diagonal(a.P, read, M, win, lo, hi, L, cig, &n, &cells)) { r.n = n; r.cells_glob = cells; }
                else {
                    need = true;
                    t.best = L.best; t.endi = L.endi; t.endj = L.endj; t.starti = L.starti; t.startj = L.startj;
                }
            }
        }
        if (active || round == 0) p.aln[(int64_t)round * a.n + idx] = r;
        if (round == 0) { Aln z; memset(&z, 0, sizeof(z)); p.aln[(int64_t)a.n + idx] = z; }
        return need;
    };
    auto phase2 = [&](const DcTask& t) {
        const int idx = t.idx;
        ReadCtx c; uint32_t zs1 = 0, e1 = 0, zs2 = 0, e2 = 0; int low = 0;
        pipe_task(a, p, round, idx, &c, &zs1, &e1, &zs2, &e2, &low);
        const int N = (int)(e1 - zs1), M = (int)(e2 - zs2);
        const int lo = max(-M, low), hi = min(N, low + a.P.g);
        BandLocal L;
        L.best = t.best; L.endi = t.endi; L.endj = t.endj; L.starti = t.starti; L.startj = t.startj; L.cf = 0; L.cr = 0; L.none = false;
        int n = 0, cells = 0, ns = 0;
        band_global<32>(a.P, bands, rowsb, a.scratch.max_band, a.scratch.max_rows, st, a.reads + c.roff + zs2, M,
                        a.ref.raw + c.cbase + zs1, lo, hi, L, p.cig + ((int64_t)round * a.n + idx) * p.cig_stride, &n, &cells, &ns);
        Aln* r = p.aln + (int64_t)round * a.n + idx;
        r->n = n; r->cells_glob = cells;
    };
    banded_two_phase_loop(a.n, s_pend[warp], phase0, phase1, phase2);
}

// ---- warp per read: combine + results ----------------------------------------------------------
__global__ void __launch_bounds__(256)
pipe_combine_kernel(const __grid_constant__ RealignArgs a, const PipeBufs p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = (4 * p.cig_stride + 8) * 4 + 256;                 // cig1, cig2, segs (2x), misc
    unsigned char* wb = smem + (size_t)warp * per_warp;
    Cta S;
    S.keys = nullptr; S.vals = nullptr; S.hist = nullptr; S.read = nullptr; S.bits = nullptr; S.psum = nullptr;
    S.cig1 = reinterpret_cast<uint32_t*>(wb);
    S.cig2 = S.cig1 + p.cig_stride;
    S.segs = S.cig2 + p.cig_stride;
    int* misc = reinterpret_cast<int*>(wb + (size_t)(4 * p.cig_stride + 8) * 4);
    Aln* s_aln = reinterpret_cast<Aln*>(misc);                             // 2 x 11 ints
    Plan* s_plan = reinterpret_cast<Plan*>(misc + 24);
    int* s_final = misc + 40;
    unsigned long long cells_f = 0, cells_r = 0, cells_g = 0, alg_bytes = 0;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
#pragma unroll 1
    for (int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < a.n; idx += nwarps) {
        const ReadCtx c = load_read_ctx(a, idx);
        if (c.bad) {
            if (lane == 0) {
                a.status[idx] = ST_ASSERT; a.nseg[idx] = 0; a.rstart[idx] = 0; a.seg_off[idx] = 0;
                if (a.detail) memset(&a.detail[idx], 0, sizeof(indelgpu_detail));
                atomicExch(a.error_flag, 1);
            }
            continue;
        }
        const int flags = p.flags[idx];
        __syncwarp();
        if (lane == 0) {
            s_aln[0] = p.aln[idx]; s_aln[1] = p.aln[(int64_t)a.n + idx];
            if (flags & PF_VOTE1_OK) *s_plan = p.plan[idx]; else { s_plan->go = 0; s_plan->status = ST_ASSERT; }
            s_final[0] = 0; s_final[1] = 0; s_final[2] = 0; s_final[3] = -1;
            if ((flags & PF_GO) && !(flags & PF_VOTE2_OK)) s_final[0] = ST_ASSERT;
        }
        __syncwarp();
        const int n1 = s_aln[0].n, n2 = s_aln[1].n;
#pragma unroll 1
        for (int t = lane; t < n1; t += 32) S.cig1[t] = p.cig[(int64_t)idx * p.cig_stride + t];
#pragma unroll 1
        for (int t = lane; t < n2; t += 32) S.cig2[t] = p.cig[((int64_t)a.n + idx) * p.cig_stride + t];
        __syncwarp();
        const bool done2 = (flags & PF_VOTE2_OK) != 0;
        combine_read(S, c.readlen, s_aln, s_plan, s_final, done2);
        __syncwarp();
        const Aln* s_a1 = s_aln; const Aln* s_a2 = s_aln + 1;
        const int ns = s_final[1];
        long long off = 0;
        if (lane == 0) {
            if (ns > 0) off = (long long)atomicAdd(a.seg_count, (unsigned long long)ns);
            if (off + ns > a.seg_capacity) { atomicExch(a.error_flag, 2); off = -1; }
            a.status[idx] = s_final[0]; a.nseg[idx] = off < 0 ? 0 : ns;
            a.rstart[idx] = s_final[2]; a.seg_off[idx] = off < 0 ? 0 : off;
            cells_f += (unsigned long long)(s_a1->cells_fwd + s_a2->cells_fwd);
            cells_r += (unsigned long long)(s_a1->cells_rev + s_a2->cells_rev);
            cells_g += (unsigned long long)(s_a1->cells_glob + s_a2->cells_glob);
            alg_bytes += (unsigned long long)((c.right1 - c.left1) + c.readlen + 4 * (6 + s_a1->n));
            if (s_plan->go) alg_bytes += (unsigned long long)((int)(s_plan->e1 - s_plan->zs1) + (int)(s_plan->e2 - s_plan->zs2) + 4 * (6 + s_a2->n));
            if (a.detail) {
                indelgpu_detail d;
                d.low1 = s_a1->low; d.up1 = s_a1->up; d.r1 = s_a1->r1; d.r2 = s_a1->r2; d.q1 = s_a1->q1; d.q2 = s_a1->q2;
                d.n1 = s_a1->n; d.score1 = s_a1->score;
                d.low2 = s_a2->low; d.up2 = s_a2->up; d.r3 = s_a2->r1; d.r4 = s_a2->r2; d.q3 = s_a2->q1; d.q4 = s_a2->q2;
                d.n2 = s_a2->n; d.score2 = s_a2->score;
                d.index = s_final[3];
                d.cells_fwd = s_a1->cells_fwd + s_a2->cells_fwd;
                d.cells_rev = s_a1->cells_rev + s_a2->cells_rev;
                d.cells_glob = s_a1->cells_glob + s_a2->cells_glob;
                a.detail[idx] = d;
            }
        }
        off = __shfl_sync(0xFFFFFFFFu, off, 0);
        if (off >= 0) for (int t = lane; t < ns; t += 32) a.segs[off + t] = S.segs[t];
        if (a.cigar1) for (int t = lane; t < min(s_a1->n, a.cigar_stride); t += 32) a.cigar1[(int64_t)idx * a.cigar_stride + t] = S.cig1[t];
        if (a.cigar2) for (int t = lane; t < min(s_a2->n, a.cigar_stride); t += 32) a.cigar2[(int64_t)idx * a.cigar_stride + t] = S.cig2[t];
    }
    if (lane == 0 && (cells_f | cells_r | cells_g | alg_bytes)) {
        atomicAdd(a.cell_totals + 0, cells_f);
        atomicAdd(a.cell_totals + 1, cells_r);
        atomicAdd(a.cell_totals + 2, cells_g);
        atomicAdd(a.cell_totals + 4, alg_bytes);
    }
}

}  // namespace indelgpu
