// The fused two-round realignment kernel (persistent CTAs, one read at a time per CTA).
#pragma once

#include "kernels.cuh"
#include "band_dp.cuh"

namespace indelgpu {

struct RealignArgs {
    DevParams P;
    RefView ref;
    int n;
    const uint8_t* reads; const int64_t* read_off;
    const int32_t* tid; const int32_t* position; const int32_t* range1;
    int32_t* status; int32_t* nseg; int32_t* rstart; int64_t* seg_off;
    uint32_t* segs; int64_t seg_capacity; unsigned long long* seg_count;
    indelgpu_detail* detail; uint32_t* cigar1; uint32_t* cigar2; int cigar_stride;
    int* work_counter;                 // zeroed before launch
    unsigned long long* cell_totals;   // 3 words: fwd, rev, glob; word [4] = algorithmic bytes
    int* error_flag;                   // set non-zero on a limit violation
    int max_read, max_numdiag;
    BandScratch scratch;               // global scratch for bands wider than one diagonal
};

// round-2 plan produced by lane 0 after round 1 (alignment.c:568-717)
struct Plan {
    int go;                  // 1 = run round 2
    int status;              // terminal status when go == 0
    uint32_t zs1, e1, anc, zs2, e2;
    int tail;                // 1: slice is the read's tail (prefix clip), 0: the head (suffix clip)
    uint32_t f_nonmatch, l_nonmatch;
};

__device__ __forceinline__ void make_plan(const DevParams& P, const Aln& A1, const uint32_t* cig1,
                                          int32_t anchor, int32_t left2, int32_t right2,
                                          unsigned readlength, Plan* pl)
{
    const int r1 = A1.r1, r2 = A1.r2, q1 = A1.q1, q2 = A1.q2, n1 = A1.n;
    const unsigned ethreshold = (unsigned)P.ethr;
    pl->go = 0;
    if (q1 == q2) { pl->status = INDELGPU_ST_UNALIGNED; return; }              // :568
    if (q1 == 0 && q2 == (int)readlength) { pl->status = INDELGPU_ST_WHOLE; return; }   // :575
    unsigned f_nonmatch, l_nonmatch;                                            // :584-599
    {
        int i, j;
        for (i = 0, j = 0; i < n1; i++) {
            const int op = cig_op(cig1[i]);
            if (i == 0 && op == OP_SOFT) continue;
            if (op != OP_EQ) break;
            j += cig_len(cig1[i]);
        }
        f_nonmatch = (unsigned)j;
        for (i = n1 - 1, j = 0; i >= 0; i--) {
            const int op = cig_op(cig1[i]);
            if (i == n1 - 1 && op == OP_SOFT) continue;
            if (op != OP_EQ) break;
            j += cig_len(cig1[i]);
        }
        l_nonmatch = (unsigned)j;
    }
    pl->f_nonmatch = f_nonmatch; pl->l_nonmatch = l_nonmatch;
    // the guards mix int and unsigned exactly as the reference does (:608-609,631-632,665-666,687-688)
    if (r1 > anchor) {
        if (q1 == 0) {
            if (((readlength - f_nonmatch) < ethreshold) || ((right2 - r1 - f_nonmatch) < ethreshold)) { pl->status = INDELGPU_ST_SHORT; return; }
            pl->zs1 = (uint32_t)r1 + f_nonmatch; pl->e1 = (uint32_t)right2; pl->anc = (uint32_t)r1;
            pl->zs2 = f_nonmatch; pl->e2 = readlength; pl->tail = 1;
        } else if (q2 == (int)readlength) {
            if (((readlength - l_nonmatch) < ethreshold) || ((r2 - l_nonmatch - anchor) < ethreshold)) { pl->status = INDELGPU_ST_SHORT; return; }
            pl->zs1 = (uint32_t)anchor; pl->e1 = (uint32_t)r2 - l_nonmatch; pl->anc = (uint32_t)r2;
            pl->zs2 = 0; pl->e2 = readlength - l_nonmatch; pl->tail = 0;
        } else { pl->status = INDELGPU_ST_NOBRANCH; return; }
    } else if (r1 < anchor) {
        if (r2 >= anchor) { pl->status = INDELGPU_ST_NOBRANCH; return; }
        if (q1 == 0) {
            if (((readlength - f_nonmatch) < ethreshold) || ((anchor - r1 - f_nonmatch) < ethreshold)) { pl->status = INDELGPU_ST_SHORT; return; }
            pl->zs1 = (uint32_t)r1 + f_nonmatch; pl->e1 = (uint32_t)anchor; pl->anc = (uint32_t)r1;
            pl->zs2 = f_nonmatch; pl->e2 = readlength; pl->tail = 1;
        } else if (q2 == (int)readlength) {
            if (((readlength - l_nonmatch) < ethreshold) || ((r2 - l_nonmatch - left2) < ethreshold)) { pl->status = INDELGPU_ST_SHORT; return; }
            pl->zs1 = (uint32_t)left2; pl->e1 = (uint32_t)r2 - l_nonmatch; pl->anc = (uint32_t)r2;
            pl->zs2 = 0; pl->e2 = readlength - l_nonmatch; pl->tail = 0;
        } else { pl->status = INDELGPU_ST_NOBRANCH; return; }
    } else { pl->status = INDELGPU_ST_NOBRANCH; return; }
    pl->go = 1;
}

// attempt_band_alignment (alignment.c:343-391) by warp 0: local_align + fetch_cigar + coordinate shift.
// BANDED = false is the default-flag build (-g 0): every band is one diagonal, so the banded DP
// (and its registers) is compiled out.
template <bool BANDED>
__device__ void band_alignment_warp(const RealignArgs& a, Cta& S, int64_t cbase,
                                    uint32_t zs1, uint32_t e1, uint32_t zs2, uint32_t e2,
                                    int low, int up, uint32_t* cig, Aln* out, int* s_tmp)
{
    const int N = (int)(e1 - zs1), M = (int)(e2 - zs2);
    const uint8_t* win = a.ref.raw + cbase + zs1;
    const int lo = max(-M, low), hi = min(N, up);                // localalign.c:70-71
    if (!BANDED || hi - lo + 1 == 1) align_diag1(a.P, S, win, N, (int)zs2, M, lo, cig, s_tmp);
    else align_banded(a.P, a.scratch, S.read + zs2, M, win, N, lo, hi, cig, S.L.ops_cap, s_tmp);
    if ((threadIdx.x & 31) == 0) {
        const int score = s_tmp[0];
        out->low = low; out->up = up; out->score = score;
        if (score <= 0) { out->r1 = out->r2 = out->q1 = out->q2 = 0; out->n = 0; }   // :365-372
        else {
            out->q1 = s_tmp[1] + (int)zs2 - 1; out->r1 = s_tmp[2] + (int)zs1 - 1;    // :385-388
            out->q2 = s_tmp[3] + (int)zs2;     out->r2 = s_tmp[4] + (int)zs1;
            out->n = s_tmp[5];
        }
        out->cells_fwd = s_tmp[6]; out->cells_rev = s_tmp[7]; out->cells_glob = s_tmp[8];
    }
    __syncwarp();
}

template <bool BANDED>
__global__ void __launch_bounds__(kThreads)
realign_kernel(const __grid_constant__ RealignArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_idx;
    __shared__ unsigned long long s_red[kWarps];
    __shared__ Aln s_a1, s_a2;
    __shared__ Plan s_plan;
    __shared__ int s_tmp[16];
    __shared__ int s_final[4];     // status, nseg, rstart, index
    __shared__ long long s_segoff;

    Cta S;
    S.L = make_layout(a.max_read, a.max_numdiag);
    S.keys = reinterpret_cast<uint32_t*>(smem + S.L.off_keys);
    S.vals = reinterpret_cast<uint32_t*>(smem + S.L.off_vals);
    S.hist = reinterpret_cast<uint32_t*>(smem + S.L.off_hist);
    S.read = smem + S.L.off_read;
    S.bits = reinterpret_cast<uint32_t*>(smem + S.L.off_bits);
    S.psum = reinterpret_cast<int*>(smem + S.L.off_psum);
    S.cig1 = reinterpret_cast<uint32_t*>(smem + S.L.off_cig1);
    S.cig2 = reinterpret_cast<uint32_t*>(smem + S.L.off_cig2);
    S.segs = reinterpret_cast<uint32_t*>(smem + S.L.off_segs);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    for (int s = tid; s < S.L.hist_words; s += kThreads) S.hist[s] = 0;
    unsigned long long cells[4] = {0, 0, 0, 0};

    while (true) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(a.work_counter, 1);
        __syncthreads();
        const int idx = s_idx;
        if (idx >= a.n) break;

        const int64_t roff = a.read_off[idx];
        const int readlen = (int)(a.read_off[idx + 1] - roff);
        const int32_t ctg = a.tid[idx];
        const int32_t position = a.position[idx];
        const int32_t range1 = a.range1[idx];
        bool bad = readlen <= 0 || readlen > a.max_read || ctg < 0 || ctg >= a.ref.ncontigs || position < 0;
        int64_t cbase = 0; int32_t reflength = 0;
        if (!bad) {
            cbase = a.ref.contig_off[ctg];
            const int64_t cl = a.ref.contig_len[ctg];
            reflength = (int32_t)cl;
            bad = cl > 0x7FFFFFFF;
        }
        // windows: alignment.c:774-783
        int32_t distance = range1;
        const int32_t left1  = position >= distance ? position - distance : 0;
        const int32_t right1 = reflength < (position + distance) ? reflength : position + distance;
        distance = (int32_t)((unsigned)range1 + (unsigned)a.P.maxdel);
        const int32_t left2  = position >= distance ? position - distance : 0;
        const int32_t right2 = reflength < (position + distance) ? reflength : position + distance;
        const int32_t anchor = position;
        // forceasserts of alignment.c:548-553
        bad = bad || !(anchor >= left1 && anchor >= left2 && anchor <= right1 && anchor <= right2 && right2 > 0);
        bad = bad || (right2 - left2) + readlen + 2 > a.max_numdiag;
        if (bad) {
            if (tid == 0) {
                a.status[idx] = ST_ASSERT; a.nseg[idx] = 0; a.rstart[idx] = 0; a.seg_off[idx] = 0;
                if (a.detail) memset(&a.detail[idx], 0, sizeof(indelgpu_detail));
                atomicExch(a.error_flag, 1);
            }
            continue;
        }
        for (int t = tid; t < readlen; t += kThreads) S.read[t] = a.reads[roff + t];
        if (tid == 0) {
            s_a2.low = s_a2.up = s_a2.score = s_a2.r1 = s_a2.r2 = s_a2.q1 = s_a2.q2 = s_a2.n = 0;
            s_a2.cells_fwd = s_a2.cells_rev = s_a2.cells_glob = 0;
            s_final[0] = 0; s_final[1] = 0; s_final[2] = 0; s_final[3] = -1;
        }
        __syncthreads();

        // ---------------- round 1 (alignment.c:555-566)
        bool ok;
        const int low1 = vote_band(a.P, S, a.ref.packed, cbase + left1, right1 - left1, 0, readlen,
                                   (int)((uint32_t)anchor - (uint32_t)left1), &ok, s_red);
        if (warp == 0) {
            if (!ok) {
                if (lane == 0) { s_plan.go = 0; s_plan.status = ST_ASSERT; s_a1 = s_a2; }
            } else {
                band_alignment_warp<BANDED>(a, S, cbase, (uint32_t)left1, (uint32_t)right1, 0, (uint32_t)readlen,
                                    low1, low1 + a.P.g, S.cig1, &s_a1, s_tmp);
                if (lane == 0) make_plan(a.P, s_a1, S.cig1, anchor, left2, right2, (unsigned)readlen, &s_plan);
            }
        }
        __syncthreads();

        // ---------------- round 2 (alignment.c:601-717)
        if (s_plan.go) {
            const Plan pl = s_plan;
            const int low2 = vote_band(a.P, S, a.ref.packed, cbase + pl.zs1, (int)(pl.e1 - pl.zs1),
                                       (int)pl.zs2, (int)(pl.e2 - pl.zs2), (int)(pl.anc - pl.zs1), &ok, s_red);
            if (warp == 0) {
                if (!ok) { if (lane == 0) s_final[0] = ST_ASSERT; }
                else {
                    band_alignment_warp<BANDED>(a, S, cbase, pl.zs1, pl.e1, pl.zs2, pl.e2,
                                        low2, low2 + a.P.g, S.cig2, &s_a2, s_tmp);
                    const int q1 = s_a1.q1, q2 = s_a1.q2, r1 = s_a1.r1, r2 = s_a1.r2, n1 = s_a1.n;
                    const int q3 = s_a2.q1, q4 = s_a2.q2, r3 = s_a2.r1, r4 = s_a2.r2;
                    int n2 = s_a2.n;
                    bool fail = pl.tail ? (q4 != readlen || q3 == q4) : (q3 != 0 || q3 == q4);
                    if (fail) { if (lane == 0) s_final[0] = INDELGPU_ST_R2FAIL; }
                    else {
                        if (lane == 0) {                 // add_prefix/suffix_soft_clip (:478-532)
                            if (pl.tail && pl.f_nonmatch) {
                                if (cig_op(S.cig2[0]) == OP_SOFT) S.cig2[0] = ((uint32_t)(cig_len(S.cig2[0]) + (int)pl.f_nonmatch) << 4) | OP_SOFT;
                                else { for (int t = n2; t > 0; t--) S.cig2[t] = S.cig2[t - 1]; S.cig2[0] = (pl.f_nonmatch << 4) | OP_SOFT; n2++; }
                            } else if (!pl.tail && pl.l_nonmatch) {
                                if (cig_op(S.cig2[n2 - 1]) == OP_SOFT) S.cig2[n2 - 1] = ((uint32_t)(cig_len(S.cig2[n2 - 1]) + (int)pl.l_nonmatch) << 4) | OP_SOFT;
                                else { S.cig2[n2] = (pl.l_nonmatch << 4) | OP_SOFT; n2++; }
                            }
                            s_a2.n = n2;
                        }
                        n2 = __shfl_sync(0xFFFFFFFFu, n2, 0);
                        __syncwarp();
                        // combine (:719-758)
                        int mode = 0, index = -1;
                        if (q1 > q3 && q1 <= q4)      { mode = 1; index = best_junction_warp(q3, q4, S.cig2, n2, q1, q2, S.cig1, n1); }
                        else if (q3 > q1 && q3 <= q2) { mode = 2; index = best_junction_warp(q1, q2, S.cig1, n1, q3, q4, S.cig2, n2); }
                        else if (q1 > q4 && r1 == r4) { mode = 3; index = q4; }
                        else if (q3 > q2 && r2 == r3) { mode = 4; index = q2; }
                        if (lane == 0) {
                            if (mode == 0) s_final[0] = INDELGPU_ST_NOCOMBINE;
                            else {
                                int ns;
                                if (mode == 1 || mode == 3) { ns = stitch_segments(S.segs, r3, S.cig2, n2, index, q1, r1, S.cig1, n1); s_final[2] = r3; }
                                else                        { ns = stitch_segments(S.segs, r1, S.cig1, n1, index, q3, r3, S.cig2, n2); s_final[2] = r1; }
                                s_final[0] = INDELGPU_ST_SPLIT; s_final[1] = ns; s_final[3] = index;
                            }
                        }
                    }
                }
            }
        } else if (warp == 0 && lane == 0) {
            s_final[0] = s_plan.status;
            if (s_plan.status == INDELGPU_ST_WHOLE) {            // :575-582
                s_final[1] = stitch_segments(S.segs, s_a1.r1, S.cig1, s_a1.n, readlen, 0, -1, nullptr, 0);
                s_final[2] = s_a1.r1; s_final[3] = readlen;
            }
        }

        // ---------------- results
        if (warp == 0) {
            __syncwarp();
            const int ns = s_final[1];
            if (lane == 0) {
                long long off = 0;
                if (ns > 0) off = (long long)atomicAdd(a.seg_count, (unsigned long long)ns);
                if (off + ns > a.seg_capacity) { atomicExch(a.error_flag, 2); off = -1; }
                s_segoff = off;
                a.status[idx] = s_final[0]; a.nseg[idx] = off < 0 ? 0 : ns;
                a.rstart[idx] = s_final[2]; a.seg_off[idx] = off < 0 ? 0 : off;
                cells[0] += (unsigned long long)(s_a1.cells_fwd + s_a2.cells_fwd);
                cells[1] += (unsigned long long)(s_a1.cells_rev + s_a2.cells_rev);
                cells[2] += (unsigned long long)(s_a1.cells_glob + s_a2.cells_glob);
                // algorithmic bytes: N + M in, 4 * (6 + ncigar) out, per alignment
                cells[3] += (unsigned long long)((right1 - left1) + readlen + 4 * (6 + s_a1.n));
                if (s_plan.go) cells[3] += (unsigned long long)((int)(s_plan.e1 - s_plan.zs1) + (int)(s_plan.e2 - s_plan.zs2) + 4 * (6 + s_a2.n));
                if (a.detail) {
                    indelgpu_detail d;
                    d.low1 = s_a1.low; d.up1 = s_a1.up; d.r1 = s_a1.r1; d.r2 = s_a1.r2; d.q1 = s_a1.q1; d.q2 = s_a1.q2;
                    d.n1 = s_a1.n; d.score1 = s_a1.score;
                    d.low2 = s_a2.low; d.up2 = s_a2.up; d.r3 = s_a2.r1; d.r4 = s_a2.r2; d.q3 = s_a2.q1; d.q4 = s_a2.q2;
                    d.n2 = s_a2.n; d.score2 = s_a2.score;
                    d.index = s_final[3];
                    d.cells_fwd = s_a1.cells_fwd + s_a2.cells_fwd;
                    d.cells_rev = s_a1.cells_rev + s_a2.cells_rev;
                    d.cells_glob = s_a1.cells_glob + s_a2.cells_glob;
                    a.detail[idx] = d;
                }
            }
            __syncwarp();
            const long long off = s_segoff;
            if (off >= 0) for (int t = lane; t < ns; t += 32) a.segs[off + t] = S.segs[t];
            if (a.cigar1) for (int t = lane; t < min(s_a1.n, a.cigar_stride); t += 32) a.cigar1[(int64_t)idx * a.cigar_stride + t] = S.cig1[t];
            if (a.cigar2) for (int t = lane; t < min(s_a2.n, a.cigar_stride); t += 32) a.cigar2[(int64_t)idx * a.cigar_stride + t] = S.cig2[t];
        }
    }
    if (tid == 0 && (cells[0] | cells[1] | cells[2] | cells[3])) {
        atomicAdd(a.cell_totals + 0, cells[0]);
        atomicAdd(a.cell_totals + 1, cells[1]);
        atomicAdd(a.cell_totals + 2, cells[2]);
        atomicAdd(a.cell_totals + 4, cells[3]);
    }
}

}  // namespace indelgpu
