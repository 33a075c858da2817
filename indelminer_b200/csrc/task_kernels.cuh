// Kernel-level entry points over independent (read, window) tasks: the band-sweep micro-bench
// (SURVEY.md 8d, D1), the reference-prototype shims and the kernel parity tests.
#pragma once

#include "kernels.cuh"
#include "band_dp.cuh"
#include "warp_vote.cuh"

namespace indelgpu {

struct TaskArgs {
    DevParams P;
    int n;
    const uint8_t* reads; const int64_t* read_off;
    const uint8_t* refs;  const int64_t* ref_off;
    const uint32_t* packed;            // 2-bit copy of refs (vote only)
    const int32_t* anchor_rel;         // vote only
    int32_t* low; int32_t* up;         // vote: out; align: in
    int32_t* score; int32_t* ends; int32_t* ncigar; uint32_t* cigar; int cigar_stride;
    int32_t* script; int script_stride;
    int* work_counter;
    unsigned long long* cell_totals;
    int* error_flag;
    int max_read, max_numdiag;
    WarpLayout L;                      // vote kernel only
    BandScratch scratch;
};

__device__ __forceinline__ void bind_smem(Cta& S, unsigned char* smem, int max_read, int max_numdiag)
{
    S.L = make_layout(max_read, max_numdiag);
    S.keys = reinterpret_cast<uint32_t*>(smem + S.L.off_keys);
    S.vals = reinterpret_cast<uint32_t*>(smem + S.L.off_vals);
    S.hist = reinterpret_cast<uint32_t*>(smem + S.L.off_hist);
    S.read = smem + S.L.off_read;
    S.bits = reinterpret_cast<uint32_t*>(smem + S.L.off_bits);
    S.psum = reinterpret_cast<int*>(smem + S.L.off_psum);
    S.cig1 = reinterpret_cast<uint32_t*>(smem + S.L.off_cig1);
    S.cig2 = reinterpret_cast<uint32_t*>(smem + S.L.off_cig2);
    S.segs = reinterpret_cast<uint32_t*>(smem + S.L.off_segs);
}

// find_best_band over n tasks (alignment.c:393-447 with zstart1 = zstart2 = 0), one warp per task;
// the packed window is staged by a TMA bulk copy like in the fused kernel
template <bool DIRECT, int HB>
__global__ void __launch_bounds__(256)
vote_tasks_kernel(const __grid_constant__ TaskArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpView V;
    bind_warp(V, smem + (size_t)warp * a.L.total, a.L);
    init_warp_tables(V);
    if (lane == 0) { mbar_init(V.bar, 1); mbar_fence_init(); }
    __syncwarp();
    uint32_t phase = 0;
    while (true) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(a.work_counter, 1);
        idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
        if (idx >= a.n) break;
        const int64_t roff = a.read_off[idx], woff = a.ref_off[idx];
        const int M = (int)(a.read_off[idx + 1] - roff), N = (int)(a.ref_off[idx + 1] - woff);
        if (M <= 0 || M > a.max_read || N <= 0 || N + M + 2 > a.max_numdiag) {
            if (lane == 0) { a.low[idx] = 0; a.up[idx] = 0; atomicExch(a.error_flag, 1); }
            continue;
        }
        int64_t sw0;
        const uint32_t wbytes = window_span_bytes(woff, woff + N, &sw0);
        if (lane == 0) { mbar_arrive_expect_tx(V.bar, wbytes); bulk_g2s(V.win0, a.packed + sw0, wbytes, V.bar); }
        uint8_t* rd = V.rbuf0;
        for (int t = lane; t < M; t += 32) rd[t] = a.reads[roff + t];
        __syncwarp();
        pack_read_warp(V, rd, M);
        if (!mbar_wait(V.bar, phase)) { if (lane == 0) atomicExch(a.error_flag, 3); break; }
        phase ^= 1u;
        bool ok;
        const int low = vote_band_warp<DIRECT, HB>(a.P, V, V.win0, sw0, woff, N, 0, M, a.anchor_rel[idx], &ok);
        if (lane == 0) {
            if (!ok) { a.low[idx] = 0; a.up[idx] = 0; atomicExch(a.error_flag, 1); }
            else { a.low[idx] = low; a.up[idx] = low + (M < a.P.k ? 0 : a.P.g); }
        }
        __syncwarp();
    }
}

// local_align + ALIGN + fetch_cigar over n tasks with bands of ONE diagonal, one warp per task
__global__ void __launch_bounds__(32)
align_tasks_kernel(const __grid_constant__ TaskArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_idx;
    __shared__ int s_tmp[16];
    Cta S;
    bind_smem(S, smem, a.max_read, 0);
    const int lane = threadIdx.x;
    unsigned long long cells[3] = {0, 0, 0};
    while (true) {
        __syncwarp();
        if (lane == 0) s_idx = atomicAdd(a.work_counter, 1);
        __syncwarp();
        const int idx = s_idx;
        if (idx >= a.n) break;
        const int64_t roff = a.read_off[idx], woff = a.ref_off[idx];
        const int M = (int)(a.read_off[idx + 1] - roff), N = (int)(a.ref_off[idx + 1] - woff);
        const int lo = max(-M, a.low[idx]), hi = min(N, a.up[idx]);       // localalign.c:70-71
        const int band = hi - lo + 1;
        const bool bad = M <= 0 || N <= 0 || M > a.max_read || band != 1;   // wider bands: band_tasks_kernel
        if (bad) {
            if (lane == 0) { a.score[idx] = 0; a.ncigar[idx] = 0; for (int t = 0; t < 4; t++) a.ends[4 * idx + t] = 0; atomicExch(a.error_flag, 1); }
            continue;
        }
        for (int t = lane; t < M; t += 32) S.read[t] = a.reads[roff + t];
        __syncwarp();
        const uint8_t* win = a.refs + woff;
        align_diag1(a.P, S, win, N, 0, M, lo, S.cig1, s_tmp);
        const int score = s_tmp[0], n = s_tmp[5];
        if (lane == 0) {
            a.score[idx] = score;
            for (int t = 0; t < 4; t++) a.ends[4 * idx + t] = score > 0 ? s_tmp[1 + t] : 0;   // q1 r1 q2 r2
            a.ncigar[idx] = score > 0 ? n : 0;
            cells[0] += (unsigned long long)s_tmp[6]; cells[1] += (unsigned long long)s_tmp[7]; cells[2] += (unsigned long long)s_tmp[8];
        }
        if (score > 0 && a.cigar) for (int t = lane; t < min(n, a.cigar_stride); t += 32) a.cigar[(int64_t)idx * a.cigar_stride + t] = S.cig1[t];
        if (score > 0 && a.script) {
            int32_t* out = a.script + (int64_t)idx * a.script_stride;
            const int len = s_tmp[3] - s_tmp[1] + 1;                     // all-REP script (globalalign.c:358-365)
            for (int t = lane; t < min(len, a.script_stride); t += 32) out[t] = 0;
            if (lane == 0 && len < a.script_stride) out[len] = 0x7FFFFFFF;
        }
    }
    if (lane == 0 && (cells[0] | cells[1] | cells[2])) {
        atomicAdd(a.cell_totals + 0, cells[0]);
        atomicAdd(a.cell_totals + 1, cells[1]);
        atomicAdd(a.cell_totals + 2, cells[2]);
    }
}

// local_align + ALIGN + fetch_cigar over n tasks, ONE THREAD PER TASK (inter-task parallelism).
// At the band widths the caller produces (numgaps + 1 diagonals, typically <= 17) one anti-diagonal
// of a band holds at most band/2 independent cells, so a wavefront inside one alignment would leave
// most of a warp idle; 32 independent alignments per warp keep every lane busy.
//
// Every lane first runs the two sweeps of local_align (registers for bands <= 40) and the
// unique-diagonal shortcut; the few tasks that really need ALIGN's divide and conquer wait in a per-warp
// buffer until 32 of them can run it together (banded_two_phase_loop) on lane-interleaved scratch
// (IArr<32>: one 128-byte line per access, L1-resident).  Doing both in one pass per task would make every warp wait
// for its two or three gapped alignments at 10 % lane utilisation.
__global__ void __launch_bounds__(128, BAND_MIN_BLOCKS)
band_tasks_kernel(const __grid_constant__ TaskArgs a)
{
    __shared__ DcTask s_pend[4][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int wb4 = 4 * (a.scratch.max_band + 4);
    const IArr<32> gbase{a.scratch.base + (long long)gwarp * 32 * a.scratch.stride + lane};
    const IArr<32> bands = gbase;            // L1-resident; shared memory measured slower (it costs resident warps)
    const IArr<32> rowsb = gbase + wb4;
    unsigned long long cf = 0, cr = 0, cg = 0, cskip = 0;          // cskip: ALIGN cells counted but not swept (unique-diagonal shortcut)
    DcFrame st[kDcFrames];

    auto geometry = [&](int idx, const uint8_t** read, int* M, const uint8_t** win, int* N, int* lo, int* hi) {
        const int64_t roff = a.read_off[idx], woff = a.ref_off[idx];
        *M = (int)(a.read_off[idx + 1] - roff); *N = (int)(a.ref_off[idx + 1] - woff);
        *lo = max(-*M, a.low[idx]); *hi = min(*N, a.up[idx]);                // localalign.c:70-71
        *read = a.reads + roff; *win = a.refs + woff;
    };
    auto phase0 = [&](int idx, bool valid, BandLocal& L) -> bool {
        bool mine = false;
        if (valid) {
            const uint8_t* read; const uint8_t* win; int M, N, lo, hi;
            geometry(idx, &read, &M, &win, &N, &lo, &hi);
            const int band = hi - lo + 1;
            mine = M > 0 && N > 0 && band >= kWarpBandMin && band <= kWarpBandMax && a.P.G >= 0 && a.P.H >= 0 &&
                   2 * band <= a.scratch.max_band && M <= a.scratch.max_rows && 2 * M + band + 4 <= a.cigar_stride;
        }
        warp_serve_wide_bands(a.P, idx, mine, geometry, L);
        return mine;
    };
    auto phase1 = [&](int idx, DcTask& t, const BandLocal* pre) -> bool {
        const int64_t roff = a.read_off[idx], woff = a.ref_off[idx];
        const int M = (int)(a.read_off[idx + 1] - roff), N = (int)(a.ref_off[idx + 1] - woff);
        const int lo = max(-M, a.low[idx]), hi = min(N, a.up[idx]);       // localalign.c:70-71
        const int band = hi - lo + 1;
        const bool bad = M <= 0 || N <= 0 || band < 1 || 2 * band > a.scratch.max_band || M > a.scratch.max_rows ||
                         2 * M + band + 4 > a.cigar_stride;
        if (bad) {
            a.score[idx] = 0; a.ncigar[idx] = 0;
            for (int u = 0; u < 4; u++) a.ends[4 * idx + u] = 0;
            atomicExch(a.error_flag, 1);
            return false;
        }
        const uint8_t* read = a.reads + roff;
        const uint8_t* win = a.refs + woff;
        const BandLocal L = pre ? *pre : band_local<32>(a.P, bands, a.scratch.max_band, read, M, win, N, lo, hi);
        const int score = L.none ? 0 : L.best;
        a.score[idx] = score;
        a.ends[4 * idx + 0] = score > 0 ? L.starti : 0; a.ends[4 * idx + 1] = score > 0 ? L.startj : 0;
        a.ends[4 * idx + 2] = score > 0 ? L.endi : 0;   a.ends[4 * idx + 3] = score > 0 ? L.endj : 0;
        cf += (unsigned long long)L.cf; cr += (unsigned long long)L.cr;
        int n = 0, cells = 0;
        bool need = false;
        if (!L.none) {
            if (band_unique_diagonal(a.P, read, M, win, lo, hi, L, a.cigar + (int64_t)idx * a.cigar_stride, &n, &cells)) {
                cg += (unsigned long long)cells; cskip += (unsigned long long)cells;
                if (a.script) {
                    int32_t* so = a.script + (int64_t)idx * a.script_stride;
                    const int len = L.endi - L.starti + 1;
                    for (int u = 0; u < min(len, a.script_stride); u++) so[u] = 0;
                    if (len < a.script_stride) so[len] = 0x7FFFFFFF;
                }
            } else {
                need = true;
                t.best = L.best; t.endi = L.endi; t.endj = L.endj; t.starti = L.starti; t.startj = L.startj;
            }
        }
        a.ncigar[idx] = n;
        return need;
    };
    auto phase2 = [&](const DcTask& t) {
        const int idx = t.idx;
        const int64_t roff = a.read_off[idx], woff = a.ref_off[idx];
        const int M = (int)(a.read_off[idx + 1] - roff), N = (int)(a.ref_off[idx + 1] - woff);
        const int lo = max(-M, a.low[idx]), hi = min(N, a.up[idx]);
        BandLocal L;
        L.best = t.best; L.endi = t.endi; L.endj = t.endj; L.starti = t.starti; L.startj = t.startj; L.cf = 0; L.cr = 0; L.none = false;
        int n = 0, cells = 0, ns = 0;
        band_global<32>(a.P, bands, rowsb, a.scratch.max_band, a.scratch.max_rows, st, a.reads + roff, M, a.refs + woff, lo, hi, L,
                        a.cigar + (int64_t)idx * a.cigar_stride, &n, &cells, &ns);
        a.ncigar[idx] = n;
        cg += (unsigned long long)cells;
        if (a.script) {
            int32_t* so = a.script + (int64_t)idx * a.script_stride;
            const IArr<32> S = rowsb + 4 * (a.scratch.max_rows + 2);
            for (int u = 0; u < min(ns, a.script_stride); u++) so[u] = S[u];
            if (ns < a.script_stride) so[ns] = 0x7FFFFFFF;
        }
    };
    banded_two_phase_loop(a.n, s_pend[warp], phase0, phase1, phase2);

    __syncwarp();
    cf = warp_sum_u64(cf);
    cr = warp_sum_u64(cr);
    cg = warp_sum_u64(cg);
    cskip = warp_sum_u64(cskip);
    if (lane == 0 && (cf | cr | cg)) {
        atomicAdd(a.cell_totals + 0, cf);
        atomicAdd(a.cell_totals + 1, cr);
        atomicAdd(a.cell_totals + 2, cg);
        atomicAdd(a.cell_totals + 5, cskip);          // counters byte 56
    }
}

// ALIGN (globalalign.c:333-401) for one pair; the score is the re-scored script
// (CHECK_SCORE, globalalign.c:311-330, which the reference asserts equal to align()'s value)
__global__ void global_align_one_kernel(DevParams P, BandScratch scr, const uint8_t* A, const uint8_t* B,
                                        int M, int N, int low, int up, int* out_script, int* out_meta)
{
    if (threadIdx.x != 0) return;
    const IArr<1> base{scr.base};
    const int wb = scr.max_band + 4, wr = scr.max_rows + 2;
    DcCtx<1> x;
    x.P = &P; x.A = A; x.B = B; x.cells = 0;
    x.cc = base; x.dd = base + wb; x.cp = base + 2 * wb; x.dp = base + 3 * wb;
    const IArr<1> rows = base + 4 * wb;
    x.rec[0] = rows; x.rec[1] = rows + wr; x.rec[2] = rows + 2 * wr; x.fl = rows + 3 * wr;
    x.S = IArr<1>{out_script};
    DcFrame st[kDcFrames];
    const int ns = global_align_script(x, st, M, N, low, up);
    int score = 0, i = 0, j = 0;
    for (int t = 0; t < ns; t++) {
        const int op = x.S[t];
        if (op == 0) { score += (A[i] == B[j]) ? P.match : P.mismatch; i++; j++; }
        else if (op > 0) { score -= P.G + op * P.H; j += op; }
        else { score -= P.G - op * P.H; i -= op; }
    }
    out_meta[0] = score; out_meta[1] = ns; out_meta[2] = x.cells;
}

// fetch_cigar (globalalign.c:507-604) for one script
__global__ void fetch_cigar_one_kernel(const uint8_t* A, const uint8_t* B, int M, int N, const int* S,
                                       int AP, int readlength, uint32_t* cig, int* out_meta)
{
    if (threadIdx.x != 0) return;
    const int n = script_to_cigar(A, B, M, N, S, AP, readlength, cig);
    int mm = 0;
    for (int t = 0; t < n; t++) if ((cig[t] & 15u) == OP_X) mm += (int)(cig[t] >> 4);
    out_meta[0] = n; out_meta[1] = mm;
}

}  // namespace indelgpu
