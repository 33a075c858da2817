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
template <bool DIRECT, int HB>
__global__ void __launch_bounds__(256)
pipe_vote_kernel(const __grid_constant__ RealignArgs a, const PipeBufs p, const int round)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpView V;
    bind_warp(V, smem + (size_t)warp * a.L.total, a.L);
    init_warp_tables(V);
    if (lane == 0) { mbar_init(V.bar, 1); mbar_fence_init(); }
    __syncwarp();
    uint32_t phase = 0;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
#pragma unroll 1
    for (int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < a.n; idx += nwarps) {
        const ReadCtx c = load_read_ctx(a, idx);
        if (c.bad) { if (round == 0 && lane == 0) p.flags[idx] = 0; continue; }
        const int flags = round ? p.flags[idx] : 0;
        if (round == 1 && !(flags & PF_VOTE1_OK)) continue;
        uint32_t zs1 = (uint32_t)c.left1, e1 = (uint32_t)c.right1, zs2 = 0, e2 = (uint32_t)c.readlen, anc = (uint32_t)c.position;
        if (round == 1) {                                         // plan: alignment.c:568-717
            Plan* s_plan = reinterpret_cast<Plan*>(V.misc);
            if (lane == 0) {
                const Aln a1 = p.aln[idx];
                make_plan(a.P, a1, p.cig + (int64_t)idx * p.cig_stride, c.position, c.left2, c.right2, (unsigned)c.readlen, s_plan);
                p.plan[idx] = *s_plan;
            }
            __syncwarp();
            const bool go = s_plan->go != 0;
            zs1 = s_plan->zs1; e1 = s_plan->e1; zs2 = s_plan->zs2; e2 = s_plan->e2; anc = s_plan->anc;
            __syncwarp();
            if (!go) continue;
        }
        if (lane == 0) stage_read(a, V, c, 0);
        const bool landed = mbar_wait(V.bar, phase);
        phase ^= 1u;
        if (!landed) { if (lane == 0) atomicExch(a.error_flag, 3); break; }
        const uint8_t* read = read_buf(V, 0) + (int)(c.roff & 15);
        pack_read_warp(V, read, c.readlen);
        const int64_t sw0 = ((c.cbase + c.left2) & ~(int64_t)63) >> 4;
        bool ok;
        const int low = vote_band_warp<DIRECT, HB>(a.P, V, win_buf(V, 0), sw0, c.cbase + zs1, (int)(e1 - zs1), (int)zs2,
                                                   (int)(e2 - zs2), (int)(anc - zs1), &ok);
        if (lane == 0) {
            p.low[(int64_t)round * a.n + idx] = low;
            if (round == 0) p.flags[idx] = ok ? PF_VOTE1_OK : 0;
            else p.flags[idx] = flags | PF_GO | (ok ? PF_VOTE2_OK : 0);
        }
        __syncwarp();
    }
}

// ---- thread per read: attempt_band_alignment (alignment.c:343-391) ----------------------------
__global__ void __launch_bounds__(128, BAND_MIN_BLOCKS)
pipe_dp_kernel(const __grid_constant__ RealignArgs a, const PipeBufs p, const int round, const int bands_in_smem)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int wb4 = 4 * (a.scratch.max_band + 4);
    const IArr<32> gbase{a.scratch.base + (long long)gwarp * 32 * a.scratch.stride + lane};
    const IArr<32> bands = bands_in_smem ? IArr<32>{reinterpret_cast<int*>(smem) + (size_t)warp * 32 * wb4 + lane} : gbase;
    const IArr<32> rowsb = gbase + wb4;
    DcFrame st[kDcFrames];
#pragma unroll 1
    for (int idx0 = gwarp * 32; idx0 < a.n; idx0 += nwarps * 32) {
        const int idx = idx0 + lane;
        bool active = false;
        ReadCtx c; c.bad = true; c.roff = 0; c.cbase = 0; c.readlen = 0;
        uint32_t zs1 = 0, e1 = 0, zs2 = 0, e2 = 0;
        int low = 0;
        if (idx < a.n) {
            c = load_read_ctx(a, idx);
            const int flags = c.bad ? 0 : p.flags[idx];
            active = !c.bad && (round == 0 ? (flags & PF_VOTE1_OK) != 0 : (flags & PF_VOTE2_OK) != 0);
            if (active) {
                low = p.low[(int64_t)round * a.n + idx];
                if (round == 0) { zs1 = (uint32_t)c.left1; e1 = (uint32_t)c.right1; zs2 = 0; e2 = (uint32_t)c.readlen; }
                else { const Plan pl = p.plan[idx]; zs1 = pl.zs1; e1 = pl.e1; zs2 = pl.zs2; e2 = pl.e2; }
            }
        }
        const int N = (int)(e1 - zs1), M = (int)(e2 - zs2);
        const int lo = max(-M, low), hi = min(N, low + a.P.g);            // localalign.c:70-71
        const int band = hi - lo + 1;
        if (active && (band < 1 || 2 * band > a.scratch.max_band || M > a.scratch.max_rows || 2 * M + band + 4 > p.cig_stride)) {
            atomicExch(a.error_flag, 1);                                  // cannot happen for N, M >= 1; kept as a guard
            active = false;
        }
        const uint8_t* rptr = a.reads + c.roff + zs2;
        const uint8_t* wptr = a.ref.raw + c.cbase + zs1;
        if (!active) {
            if (idx < a.n && round == 0) { Aln z; memset(&z, 0, sizeof(z)); p.aln[idx] = z; p.aln[(int64_t)a.n + idx] = z; }
            continue;
        }
        int out[10];
        uint32_t* cig = p.cig + ((int64_t)round * a.n + idx) * p.cig_stride;
        align_banded_serial<32>(a.P, bands, rowsb, a.scratch.max_band, a.scratch.max_rows, st, rptr, M, wptr, N, lo, hi, cig, out);
        Aln r;
        r.low = low; r.up = low + a.P.g; r.score = out[0];
        if (out[0] <= 0) { r.r1 = r.r2 = r.q1 = r.q2 = 0; r.n = 0; }      // alignment.c:365-372
        else {
            r.q1 = out[1] + (int)zs2 - 1; r.r1 = out[2] + (int)zs1 - 1;   // :385-388
            r.q2 = out[3] + (int)zs2;     r.r2 = out[4] + (int)zs1;
            r.n = out[5];
        }
        r.cells_fwd = out[6]; r.cells_rev = out[7]; r.cells_glob = out[8];
        p.aln[(int64_t)round * a.n + idx] = r;
        if (round == 0) { Aln z; memset(&z, 0, sizeof(z)); p.aln[(int64_t)a.n + idx] = z; }
    }
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
