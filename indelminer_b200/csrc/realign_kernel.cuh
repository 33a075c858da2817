// The fused two-round realignment kernel: persistent warps, ONE WARP PER READ.
//
// Every warp owns a slice of shared memory (WarpLayout) and loops over reads taken from a global
// counter.  While read i is being realigned, lane 0 has already issued the TMA bulk copies
// (cp.async.bulk -> mbarrier) that stage read i+1's bytes and its packed reference window
// [left2, right2) -- a superset of every window either round scans -- into the other half of the
// warp's double buffer.  No CTA-wide barrier exists after start-up.
//
//   round 1  vote (k-mer diagonal histogram)  ->  align on the voted band  ->  CIGAR
//   plan     which window / read slice round 2 uses (attempt_diagonal_alignments' branches)
//   round 2  vote -> align -> CIGAR
//   combine  junction choice + segment stitching (update_readsegs)
#pragma once

#include "kernels.cuh"
#include "band_dp.cuh"
#include "warp_vote.cuh"
#include "kmer_index.cuh"

namespace indelgpu {

struct RealignArgs {
    DevParams P;
    RefView ref;
    KmerIndex idx;                     // resident k-mer index of the reference (used when L.indexed)
    int n;
    const uint8_t* reads; const int64_t* read_off;
    const int32_t* read_len;           // optional: lengths when the reads are not back to back (4-bit batches)
    const uint8_t* seq4; const int64_t* byte_off; const uint8_t* rflags;    // optional: the batch in the BAM's 4-bit form (fused kernel)
    const int32_t* tid; const int32_t* position; const int32_t* range1;
    int32_t* status; int32_t* nseg; int32_t* rstart; int64_t* seg_off;
    uint32_t* segs; int64_t seg_capacity; unsigned long long* seg_count;
    indelgpu_detail* detail; uint32_t* cigar1; uint32_t* cigar2; int cigar_stride;
    int* work_counter;                 // zeroed before launch
    unsigned long long* cell_totals;   // 3 words: fwd, rev, glob; word [4] = algorithmic bytes
    int* error_flag;                   // set non-zero on a limit violation
    int max_read, max_numdiag;
    WarpLayout L;                      // per-warp shared-memory slice, computed by the host
    BandScratch scratch;               // lane-interleaved work arrays of the banded DP (pipe_dp_kernel)
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
        #pragma unroll 1
        for (i = 0, j = 0; i < n1; i++) {
            const int op = cig_op(cig1[i]);
            if (i == 0 && op == OP_SOFT) continue;
            if (op != OP_EQ) break;
            j += cig_len(cig1[i]);
        }
        f_nonmatch = (unsigned)j;
        #pragma unroll 1
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

// attempt_band_alignment (alignment.c:343-391) by one warp for a band of ONE diagonal (-g 0):
// local_align + fetch_cigar + coordinate shift.  Wider bands go through realign_pipeline.cuh.
__device__ __forceinline__ void band_alignment_warp(const RealignArgs& a, Cta& S, int64_t cbase,
                                    uint32_t zs1, uint32_t e1, uint32_t zs2, uint32_t e2,
                                    int low, int up, uint32_t* cig, Aln* out, int* s_tmp)
{
    const int N = (int)(e1 - zs1), M = (int)(e2 - zs2);
    const uint8_t* win = a.ref.raw + cbase + zs1;
    const int lo = max(-M, low), hi = min(N, up);                // localalign.c:70-71
    (void)hi;
    align_diag1(a.P, S, win, N, (int)zs2, M, lo, cig, s_tmp);
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

// Second half of attempt_diagonal_alignments (alignment.c:719-758) plus the whole-read exit (:575-582),
// by one warp: soft-clip bookkeeping of round 2, junction choice, segment stitching into S.segs.
// s_final = {status, number of segment words, reference start, junction index}.
__device__ __forceinline__ void combine_read(Cta& S, int readlen, Aln* s_aln, const Plan* s_plan, int* s_final, bool done2)
{
    const int lane = threadIdx.x & 31;
    const Aln* s_a1 = s_aln; const Aln* s_a2 = s_aln + 1;
    if (done2) {
        const Plan pl = *s_plan;
        const int q1 = s_a1->q1, q2 = s_a1->q2, r1 = s_a1->r1, r2 = s_a1->r2, n1 = s_a1->n;
        const int q3 = s_a2->q1, q4 = s_a2->q2, r3 = s_a2->r1, r4 = s_a2->r2;
        int n2 = s_a2->n;
        const bool fail = pl.tail ? (q4 != readlen || q3 == q4) : (q3 != 0 || q3 == q4);
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
                s_aln[1].n = n2;
            }
            n2 = __shfl_sync(0xFFFFFFFFu, n2, 0);
            __syncwarp();
            // combine (:719-758).  "first" = the segment earlier on the read.
            int mode = 0;
            if (q1 > q3 && q1 <= q4)      mode = 1;
            else if (q3 > q1 && q3 <= q2) mode = 2;
            else if (q1 > q4 && r1 == r4) mode = 3;
            else if (q3 > q2 && r2 == r3) mode = 4;
            if (mode == 0) { if (lane == 0) s_final[0] = INDELGPU_ST_NOCOMBINE; }
            else {
                const bool second_first = (mode == 1 || mode == 3);      // round-2 segment precedes round-1's
                const uint32_t* cA = second_first ? S.cig2 : S.cig1; const int nA = second_first ? n2 : n1;
                const uint32_t* cB = second_first ? S.cig1 : S.cig2; const int nB = second_first ? n1 : n2;
                const int qA1 = second_first ? q3 : q1, qA2 = second_first ? q4 : q2, rA1 = second_first ? r3 : r1;
                const int qB1 = second_first ? q1 : q3, qB2 = second_first ? q2 : q4, rB1 = second_first ? r1 : r3;
                int index = qA2;                                         // abutting on the reference (:740-749)
                if (mode <= 2) index = best_junction_warp(qA1, qA2, cA, nA, qB1, qB2, cB, nB);
                if (lane == 0) {
                    s_final[1] = stitch_segments(S.segs, rA1, cA, nA, index, qB1, rB1, cB, nB);
                    s_final[0] = INDELGPU_ST_SPLIT; s_final[2] = rA1; s_final[3] = index;
                }
            }
        }
    } else if (lane == 0 && s_final[0] != ST_ASSERT) {
        s_final[0] = s_plan->status;
        if (s_plan->status == INDELGPU_ST_WHOLE) {            // :575-582
            s_final[1] = stitch_segments(S.segs, s_a1->r1, S.cig1, s_a1->n, readlen, 0, -1, nullptr, 0);
            s_final[2] = s_a1->r1; s_final[3] = readlen;
        }
    }
}

// everything the kernel derives from one batch entry (alignment.c:764-783 + the forceasserts :548-553)
struct ReadCtx {
    int64_t roff; int readlen;
    int64_t cbase;
    int32_t position, left1, right1, left2, right2;
    bool bad, rc;
};

template <bool PACKED = false>
__device__ __forceinline__ ReadCtx load_read_ctx(const RealignArgs& a, int idx)
{
    ReadCtx c;
    c.rc = false;
    if (PACKED) { c.roff = a.byte_off[idx]; c.readlen = a.read_len[idx]; c.rc = (a.rflags[idx] & 1u) != 0u; }
    else { c.roff = a.read_off[idx]; c.readlen = a.read_len ? a.read_len[idx] : (int)(a.read_off[idx + 1] - c.roff); }
    const int32_t ctg = a.tid[idx];
    c.position = a.position[idx];
    const int32_t range1 = a.range1[idx];
    bool bad = c.readlen <= 0 || c.readlen > a.max_read || ctg < 0 || ctg >= a.ref.ncontigs || c.position < 0 || range1 < 0;
    c.cbase = 0; int32_t reflength = 0;
    if (!bad) {
        c.cbase = a.ref.contig_off[ctg];
        const int64_t cl = a.ref.contig_len[ctg];
        reflength = (int32_t)cl;
        bad = cl > 0x7FFFFFFF;
    }
    const int32_t position = c.position;
    int32_t distance = range1;                                   // windows: alignment.c:774-783
    c.left1  = position >= distance ? position - distance : 0;
    c.right1 = reflength < (position + distance) ? reflength : position + distance;
    distance = (int32_t)((unsigned)range1 + (unsigned)a.P.maxdel);
    c.left2  = position >= distance ? position - distance : 0;
    c.right2 = reflength < (position + distance) ? reflength : position + distance;
    bad = bad || !(position >= c.left1 && position >= c.left2 && position <= c.right1 && position <= c.right2 && c.right2 > 0);
    bad = bad || (c.right2 - c.left2) + c.readlen + 2 > a.max_numdiag;
    c.bad = bad;
    return c;
}

// lane 0: register the expected bytes and issue the two bulk copies of one batch entry
template <bool PACKED = false>
__device__ __forceinline__ void stage_read(const RealignArgs& a, WarpView& V, const ReadCtx& c, int buf)
{
    uint64_t* bar = V.bar + buf;
    if (c.bad) { mbar_arrive(bar); return; }
    const int64_t r0 = c.roff & ~(int64_t)15;
    const uint8_t* rsrc = PACKED ? a.seq4 : a.reads;
    const int nstage = PACKED ? (c.readlen + 1) / 2 : c.readlen;
    const uint32_t rbytes = (uint32_t)(((c.roff + nstage - r0) + 15) / 16 * 16);
    int64_t sw0;
    const uint32_t wbytes = window_span_bytes(c.cbase + c.left2, c.cbase + c.right2, &sw0);
    if (a.L.indexed) {                                           // the vote reads the resident index: only the read is staged
        mbar_arrive_expect_tx(bar, rbytes);
        bulk_g2s(read_buf(V, buf), rsrc + r0, rbytes, bar);
        return;
    }
    mbar_arrive_expect_tx(bar, rbytes + wbytes);
    bulk_g2s(read_buf(V, buf), rsrc + r0, rbytes, bar);
    bulk_g2s(win_buf(V, buf), a.ref.packed + sw0, wbytes, bar);
}

constexpr int kWorkChunk = 2;     // batch entries a warp takes per atomic: larger values lengthen every launch's tail (8 cost the chunked host path 8 %)

#ifndef REALIGN_MIN_BLOCKS
#define REALIGN_MIN_BLOCKS 3      // caps the kernel at 85 registers so that three CTAs of 7 warps fit an SM
#endif
#ifndef REALIGN_MIN_BLOCKS_INDEXED
#define REALIGN_MIN_BLOCKS_INDEXED 4   // index vote: ~5.5 KB of shared memory per warp; 64 registers -> four CTAs of 8 warps
#endif

// The -g 0 kernel.  DIRECT: direct-address k-mer table (k <= 6); HB: bits per histogram counter and
// table entry (8 when a slice has at most 255 k-mers).
// The two rounds are ONE loop body so that the vote and the alignment exist once in the instruction
// stream: warps of a CTA are in different phases and the kernel has to stay instruction-cache friendly.
// PACKED: the batch holds the reads as the BAM does (4 bits per base); they are expanded -- and reversed where flagged -- here.
template <bool DIRECT, int HB, bool INDEXED, bool PACKED = false>
__global__ void __launch_bounds__(256, INDEXED ? REALIGN_MIN_BLOCKS_INDEXED : REALIGN_MIN_BLOCKS)
realign_kernel(const __grid_constant__ RealignArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpView V;
    bind_warp(V, smem + (size_t)warp * a.L.total, a.L);
    Cta& S = V.S;
    init_warp_tables(V);
    if (lane == 0) { mbar_init(V.bar + 0, 1); mbar_init(V.bar + 1, 1); mbar_fence_init(); }
    __syncwarp();

    Aln*  s_aln  = reinterpret_cast<Aln*>(V.misc);               // 2 x 12 ints (round 1, round 2)
    Plan* s_plan = reinterpret_cast<Plan*>(V.misc + 24);         // 10 ints
    int*  s_tmp  = V.misc + 36;                                  // 16 ints
    int*  s_final = V.misc + 52;                                 // status, nseg, rstart, index
    unsigned long long cells_f = 0, cells_r = 0, cells_g = 0, alg_bytes = 0;
    uint32_t phase = 0;                                          // bit b = parity of buffer b's barrier

    int cur = -1, it = -1;
    int wnext = 0, wend = 0;                                     // lane 0: the chunk of batch entries this warp owns
    long long pend_off = 0; int pend_ns = 0, pend_idx = -1;      // segment words of the previous read, not yet written out
    ReadCtx c;
    c.bad = true;
    #pragma unroll 1
    while (true) {
        // ---- fetch the next read and start staging it into the other buffer
        int nxt = 0;
        if (lane == 0) {                                         // reads are taken kWorkChunk at a time
            if (wnext == wend) { wnext = atomicAdd(a.work_counter, kWorkChunk); wend = wnext + kWorkChunk; }
            nxt = wnext++;
        }
        nxt = __shfl_sync(0xFFFFFFFFu, nxt, 0);
        ReadCtx cn;
        cn.bad = true;
        if (nxt < a.n) {
            cn = load_read_ctx<PACKED>(a, nxt);
            if (lane == 0) stage_read<PACKED>(a, V, cn, (it + 1) & 1);
        }
        if (cur >= 0) {
            const int idx = cur, buf = it & 1;
            if (!mbar_wait(V.bar + buf, (phase >> buf) & 1u)) { if (lane == 0) atomicExch(a.error_flag, 3); break; }
            phase ^= 1u << buf;
            if (c.bad) {
                if (lane == 0) {
                    a.status[idx] = ST_ASSERT; a.nseg[idx] = 0; a.rstart[idx] = 0; a.seg_off[idx] = 0;
                    if (a.detail) memset(&a.detail[idx], 0, sizeof(indelgpu_detail));
                    atomicExch(a.error_flag, 1);
                }
            } else {
                const int readlen = c.readlen;
                const int32_t anchor = c.position, left2 = c.left2, right2 = c.right2;
                const int64_t cbase = c.cbase;
                S.read = read_buf(V, buf) + (int)(c.roff & 15);
                if (PACKED) {                                        // the batch holds the read as the BAM does: expand it here
                    if (!expand_read4_warp(V.ascii, S.read, readlen, c.rc) && lane == 0) atomicExch(a.error_flag, 1);
                    S.read = V.ascii;
                }
                const uint32_t* swin = win_buf(V, buf);
                const int64_t sw0 = ((cbase + left2) & ~(int64_t)63) >> 4;
                pack_read_warp(V, S.read, readlen);
                if (lane < 24) V.misc[lane] = 0;                     // both Aln records
                if (lane == 0) { s_final[0] = 0; s_final[1] = 0; s_final[2] = 0; s_final[3] = -1; s_plan->go = 0; s_plan->status = ST_ASSERT; }
                __syncwarp();

                // The previous read's segment words are still in S.segs; the offset its atomicAdd returned has
                // had a whole read's worth of time to arrive.  Write them out before anything stitches again.
                if (pend_idx >= 0) {
                    const long long poff = __shfl_sync(0xFFFFFFFFu, pend_off, 0);
                    if (poff + pend_ns <= a.seg_capacity) {
                        uint32_t* dst = a.segs + poff;               // nearly always fewer than 32 words: one predicated store
                        if (lane < pend_ns) dst[lane] = S.segs[lane];
                        #pragma unroll 1
                        for (int t = lane + 32; t < pend_ns; t += 32) dst[t] = S.segs[t];
                    } else if (lane == 0) atomicExch(a.error_flag, 2);
                    if (lane == 0) a.seg_off[pend_idx] = poff;
                    pend_idx = -1;
                    __syncwarp();
                }

                // ---------------- round 1 (alignment.c:555-566), round 2 (:601-717)
                uint32_t zs1 = (uint32_t)c.left1, e1 = (uint32_t)c.right1, zs2 = 0, e2 = (uint32_t)readlen, anc = (uint32_t)anchor;
                bool done2 = false;
#pragma unroll 1
                for (int round = 0; round < 2; round++) {
                    bool ok;
                    const int low = INDEXED
                        ? vote_band_index<HB>(a.P, V, a.idx, cbase + zs1, (int)(e1 - zs1), (int)zs2, (int)(e2 - zs2), (int)(anc - zs1), &ok)
                        : vote_band_warp<DIRECT, HB>(a.P, V, swin, sw0, cbase + zs1, (int)(e1 - zs1), (int)zs2,
                                                     (int)(e2 - zs2), (int)(anc - zs1), &ok);
                    if (!ok) {                                       // numdiagonals <= numgaps: the reference aborts (alignment.c:405)
                        if (lane == 0) { s_final[0] = ST_ASSERT; s_plan->status = ST_ASSERT; atomicExch(a.error_flag, 1); }
                        break;
                    }
                    band_alignment_warp(a, S, cbase, zs1, e1, zs2, e2, low, low + a.P.g,
                                                round ? S.cig2 : S.cig1, s_aln + round, s_tmp);
                    if (round == 1) { done2 = true; break; }
                    if (lane == 0) make_plan(a.P, s_aln[0], S.cig1, anchor, left2, right2, (unsigned)readlen, s_plan);
                    __syncwarp();
                    if (!s_plan->go) break;
                    zs1 = s_plan->zs1; e1 = s_plan->e1; zs2 = s_plan->zs2; e2 = s_plan->e2; anc = s_plan->anc;
                }
                const Aln* s_a1 = s_aln; const Aln* s_a2 = s_aln + 1;

                combine_read(S, readlen, s_aln, s_plan, s_final, done2);
                __syncwarp();

                // ---------------- results
                const int ns = s_final[1];
                if (ns > 0) {
                    pend_idx = idx; pend_ns = ns;
                    if (lane == 0) pend_off = (long long)atomicAdd(a.seg_count, (unsigned long long)ns);
                }
                if (lane == 0) {
                    a.status[idx] = s_final[0]; a.nseg[idx] = ns;
                    a.rstart[idx] = s_final[2]; if (ns == 0) a.seg_off[idx] = 0;
                    cells_f += (unsigned long long)(s_a1->cells_fwd + s_a2->cells_fwd);
                    cells_r += (unsigned long long)(s_a1->cells_rev + s_a2->cells_rev);
                    cells_g += (unsigned long long)(s_a1->cells_glob + s_a2->cells_glob);
                    // algorithmic bytes: N + M in, 4 * (6 + ncigar) out, per alignment
                    alg_bytes += (unsigned long long)((c.right1 - c.left1) + readlen + 4 * (6 + s_a1->n));
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
                if (a.cigar1) for (int t = lane; t < min(s_a1->n, a.cigar_stride); t += 32) a.cigar1[(int64_t)idx * a.cigar_stride + t] = S.cig1[t];
                if (a.cigar2) for (int t = lane; t < min(s_a2->n, a.cigar_stride); t += 32) a.cigar2[(int64_t)idx * a.cigar_stride + t] = S.cig2[t];
                __syncwarp();
            }
        }
        if (nxt >= a.n) break;
        cur = nxt; c = cn; it++;
    }
    if (pend_idx >= 0) {
        const long long poff = __shfl_sync(0xFFFFFFFFu, pend_off, 0);
        if (poff + pend_ns <= a.seg_capacity) for (int t = lane; t < pend_ns; t += 32) a.segs[poff + t] = S.segs[t];
        else if (lane == 0) atomicExch(a.error_flag, 2);
        if (lane == 0) a.seg_off[pend_idx] = poff;
    }
    if (lane == 0 && (cells_f | cells_r | cells_g | alg_bytes)) {
        atomicAdd(a.cell_totals + 0, cells_f);
        atomicAdd(a.cell_totals + 1, cells_r);
        atomicAdd(a.cell_totals + 2, cells_g);
        atomicAdd(a.cell_totals + 4, alg_bytes);
    }
}

}  // namespace indelgpu
