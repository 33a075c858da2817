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

// The band of ONE diagonal, serial: what align_diag1 (kernels.cuh) computes with a warp, by one thread.  local_align with
// low == up is a maximum-segment scan (SURVEY.md 8a'): the forward sweep keeps run = max(0, run + w) and its first strict
// maximum, the reverse sweep walks down from the end row until the suffix sum reaches the optimum (localalign.c:100-176),
// ALIGN takes its all-REP exit (globalalign.c:358-365) and fetch_cigar (:507-604) is the run-length code of the match
// mask.  out9 as align_diag1's s_out; cig may be null (the operations are then only counted).
// unaligned little-endian 32-bit read of bytes p[0..3] from two aligned words (the sequences of a batch are packed back
// to back; the buffers are padded, so the word holding p[3] exists)
__host__ __device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const unsigned sh = 8u * (unsigned)(a & 3);
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (w[1] << (32u - sh));
}

constexpr int kDiag1MaskWords = 5;                 // rows held as a match mask in registers: 160

__host__ __device__ inline void align_diag1_serial(const DevParams& P, const uint8_t* __restrict__ read, int M,
                                                   const uint8_t* __restrict__ win, int N, int d, uint32_t* cig, int cig_cap, int* out9)
{
    const int si = d < 0 ? -d : 0, ei = M < N - d ? M : N - d;                 // localalign.c:86-87
    const uint8_t* w0 = win + d;                                                // row i compares read[i-1] with w0[i-1]
    const int rows = ei - si;
    int best = 0, endi = si, starti = 0;
    uint32_t mask[kDiag1MaskWords];                                             // bit t: row si + 1 + t matches
    const bool small = rows <= 32 * kDiag1MaskWords;
    if (small) {
        // one pass over memory, four rows per step; everything after it works on the mask
#pragma unroll
        for (int k = 0; k < kDiag1MaskWords; k++) {
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int t = 32 * k + 4 * q;
                if (t < rows) {
                    const uint32_t x = load_u32_unaligned(read + si + t) ^ load_u32_unaligned(w0 + si + t);
                    uint32_t nib = 0;
                    if ((x & 0x000000FFu) == 0u) nib |= 1u;
                    if ((x & 0x0000FF00u) == 0u) nib |= 2u;
                    if ((x & 0x00FF0000u) == 0u) nib |= 4u;
                    if ((x & 0xFF000000u) == 0u) nib |= 8u;
                    m |= nib << (4 * q);
                }
            }
            const int left = rows - 32 * k;                                     // rows past the end match nothing
            mask[k] = left >= 32 ? m : (left <= 0 ? 0u : (m & ((1u << left) - 1u)));
        }
        int run = 0;
#pragma unroll
        for (int k = 0; k < kDiag1MaskWords; k++) {
            const uint32_t m = mask[k];
            const int nb = rows - 32 * k < 32 ? rows - 32 * k : 32;
            for (int b = 0; b < nb; b++) {
                run += ((m >> b) & 1u) ? P.match : P.mismatch;
                if (run < 0) run = 0;
                if (run > best) { best = run; endi = si + 32 * k + b + 1; }   // strict: the first maximum
            }
        }
        if (best > 0) {
            int s = 0;
            const int tend = endi - 1 - si;                                     // bit of the end row
            bool found = false;
#pragma unroll
            for (int k = kDiag1MaskWords - 1; k >= 0; k--) {
                if (!found && k <= (tend >> 5)) {
                    const uint32_t m = mask[k];
                    for (int b = (k == (tend >> 5)) ? (tend & 31) : 31; b >= 0; b--) {
                        s += ((m >> b) & 1u) ? P.match : P.mismatch;
                        if (s == best) { starti = si + 32 * k + b + 1; found = true; break; }
                    }
                }
            }
        }
    } else {
        int run = 0;
        for (int i = si + 1; i <= ei; i++) {
            run += (read[i - 1] == w0[i - 1]) ? P.match : P.mismatch;
            if (run < 0) run = 0;
            if (run > best) { best = run; endi = i; }
        }
        if (best > 0) {
            int s = 0;
            for (int i = endi; i > si; i--) {
                s += (read[i - 1] == w0[i - 1]) ? P.match : P.mismatch;
                if (s == best) { starti = i; break; }
            }
        }
    }
    const bool none = best <= 0 || starti == 0 || endi == starti;              // localalign.c:191-193
    int n = 0;
    if (!none) {
        if (starti - 1 > 0) { if (cig && n < cig_cap) cig[n] = ((uint32_t)(starti - 1) << 4) | OP_SOFT; n++; }
        if (small) {
            // runs of equal bits of the mask between the start and the end row, a word at a time
            const int tlast = endi - 1 - si;
            int t = starti - 1 - si;
            while (t <= tlast) {
                uint32_t m = mask[0];
#pragma unroll
                for (int k = 1; k < kDiag1MaskWords; k++) if ((t >> 5) == k) m = mask[k];
                const bool eq = ((m >> (t & 31)) & 1u) != 0u;
                int q = t;                                                      // first position after the run
                for (;;) {
                    uint32_t mq = mask[0];
#pragma unroll
                    for (int k = 1; k < kDiag1MaskWords; k++) if ((q >> 5) == k) mq = mask[k];
                    const uint32_t diff = (eq ? ~mq : mq) >> (q & 31);          // 1 where the run is broken
                    if (diff) {
#ifdef __CUDA_ARCH__
                        q += __ffs((int)diff) - 1;
#else
                        q += __builtin_ctz(diff);
#endif
                        break;
                    }
                    q += 32 - (q & 31);
                    if (q > tlast) break;
                }
                if (q > tlast + 1) q = tlast + 1;
                if (cig && n < cig_cap) cig[n] = ((uint32_t)(q - t) << 4) | (eq ? OP_EQ : OP_X);
                n++;
                t = q;
            }
        } else {
            int i = starti;
            while (i <= endi) {
                const bool eq = read[i - 1] == w0[i - 1];
                int q = i + 1;
                while (q <= endi && (read[q - 1] == w0[q - 1]) == eq) q++;
                if (cig && n < cig_cap) cig[n] = ((uint32_t)(q - i) << 4) | (eq ? OP_EQ : OP_X);
                n++;
                i = q;
            }
        }
        if (M - endi > 0) { if (cig && n < cig_cap) cig[n] = ((uint32_t)(M - endi) << 4) | OP_SOFT; n++; }
    }
    out9[0] = none ? 0 : best;
    out9[1] = none ? 0 : starti;         // q1 (1-based inclusive)
    out9[2] = none ? 0 : starti + d;     // r1
    out9[3] = none ? 0 : endi;           // q2
    out9[4] = none ? 0 : endi + d;       // r2
    out9[5] = n;
    out9[6] = ei - si;                                   // forward cells
    out9[7] = best > 0 ? endi - starti + 1 : 0;          // reverse cells until the hit
    out9[8] = 0;                                         // ALIGN exits before any sweep (band <= 1)
}

#ifdef __CUDACC__
// local_align + ALIGN + fetch_cigar over n tasks with bands of ONE diagonal, one THREAD per task: the warp-per-task kernel
// above spends two scans and a serial CIGAR loop of lane 0 on 150 rows; 32 independent tasks per warp do the same work
// with every lane busy.  A thread reading its own task's bytes would make every load of the warp touch 32 cache lines
// (measured: 0.77 ms per 2^20 tasks, bound by the load unit), so the warp first copies the rows each of its 32 tasks
// needs -- M bytes of the read, M bytes along the window's diagonal -- into shared memory with coalesced loads, one
// task after the other, and every lane then works on its own row of that tile (stride 41 words: conflict-free).
constexpr int kAlign1RowWords = 41;               // 160 rows + 3 bytes of misalignment, rounded up; odd
__global__ void __launch_bounds__(128)
align1_tasks_kernel(const __grid_constant__ TaskArgs a)
{
    __shared__ uint32_t s_stage[4][2][32 * kAlign1RowWords];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* const sr = s_stage[warp][0];
    uint32_t* const sw = s_stage[warp][1];
    unsigned long long cf = 0, cr = 0;
    const long long gwarp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = gwarp * 32; base < a.n; base += nwarps * 32) {
        const long long idx = base + lane;
        const bool have = idx < a.n;
        int64_t roff = 0, woff = 0; int M = 0, N = 0, lo = 0, rows = 0, si = 0;
        bool valid = false;
        if (have) {
            roff = a.read_off[idx]; woff = a.ref_off[idx];
            M = (int)(a.read_off[idx + 1] - roff); N = (int)(a.ref_off[idx + 1] - woff);
            lo = max(-M, a.low[idx]);
            const int hi = min(N, a.up[idx]);                            // localalign.c:70-71
            valid = !(M <= 0 || N <= 0 || M > a.max_read || hi - lo + 1 != 1);
            si = lo < 0 ? -lo : 0;
            rows = min(M, N - lo) - si;
        }
        const bool fast = valid && rows > 0 && rows <= 32 * kDiag1MaskWords;
        const uint8_t* rsrc = a.reads + roff + si;                       // first byte of the rows this task compares
        const uint8_t* wsrc = a.refs + woff + lo + si;
        // stage: task q's bytes, whole aligned words, by all lanes
#pragma unroll 4
        for (int q = 0; q < 32; q++) {                                  // unrolled: the loads of four tasks are in flight together
            const bool fq = __shfl_sync(FULL, (int)fast, q) != 0;
            const unsigned long long rq = __shfl_sync(FULL, (unsigned long long)reinterpret_cast<uintptr_t>(rsrc), q);
            const unsigned long long wq = __shfl_sync(FULL, (unsigned long long)reinterpret_cast<uintptr_t>(wsrc), q);
            const int nrow = __shfl_sync(FULL, rows, q);
            const uint32_t* rb = reinterpret_cast<const uint32_t*>(rq & ~3ull);
            const uint32_t* wb = reinterpret_cast<const uint32_t*>(wq & ~3ull);
            const int nr = min((nrow + (int)(rq & 3) + 3) / 4 + 1, (int)kAlign1RowWords);     // + 1: the word an unaligned read runs into
            const int nw = min((nrow + (int)(wq & 3) + 3) / 4 + 1, (int)kAlign1RowWords);
            uint32_t r0 = 0, r1 = 0, w0v = 0, w1v = 0;
            if (fq) {
                if (lane < nr) r0 = __ldg(rb + lane);
                if (lane + 32 < nr) r1 = __ldg(rb + lane + 32);
                if (lane < nw) w0v = __ldg(wb + lane);
                if (lane + 32 < nw) w1v = __ldg(wb + lane + 32);
                if (lane < nr) sr[q * kAlign1RowWords + lane] = r0;
                if (lane + 32 < nr) sr[q * kAlign1RowWords + lane + 32] = r1;
                if (lane < nw) sw[q * kAlign1RowWords + lane] = w0v;
                if (lane + 32 < nw) sw[q * kAlign1RowWords + lane + 32] = w1v;
            }
        }
        __syncwarp();
        if (have) {
            if (!valid) {
                a.score[idx] = 0; a.ncigar[idx] = 0; for (int t = 0; t < 4; t++) a.ends[4 * idx + t] = 0; atomicExch(a.error_flag, 1);
            } else {
                // pointers with which read[si + t] and (win + lo)[si + t] land on the staged bytes
                const uint8_t* rp = a.reads + roff; const uint8_t* wp = a.refs + woff;
                if (fast) {
                    rp = reinterpret_cast<const uint8_t*>(sr + lane * kAlign1RowWords) + (reinterpret_cast<uintptr_t>(rsrc) & 3) - si;
                    wp = reinterpret_cast<const uint8_t*>(sw + lane * kAlign1RowWords) + (reinterpret_cast<uintptr_t>(wsrc) & 3) - lo - si;
                }
                int out[9];
                align_diag1_serial(a.P, rp, M, wp, N, lo, a.cigar ? a.cigar + idx * a.cigar_stride : nullptr, a.cigar_stride, out);
                const int score = out[0];
                a.score[idx] = score;
                for (int t = 0; t < 4; t++) a.ends[4 * idx + t] = score > 0 ? out[1 + t] : 0;   // q1 r1 q2 r2
                a.ncigar[idx] = score > 0 ? out[5] : 0;
                cf += (unsigned long long)out[6]; cr += (unsigned long long)out[7];
                if (score > 0 && a.script) {
                    int32_t* so = a.script + idx * a.script_stride;
                    const int len = out[3] - out[1] + 1;                             // all-REP script (globalalign.c:358-365)
                    for (int t = 0; t < min(len, a.script_stride); t++) so[t] = 0;
                    if (len < a.script_stride) so[len] = 0x7FFFFFFF;
                }
            }
        }
        __syncwarp();                                                    // the next group overwrites the tile
    }
    for (int o = 16; o > 0; o >>= 1) { cf += __shfl_xor_sync(FULL, cf, o); cr += __shfl_xor_sync(FULL, cr, o); }
    if (lane == 0 && (cf | cr)) { atomicAdd(a.cell_totals + 0, cf); atomicAdd(a.cell_totals + 1, cr); }
}
#endif

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
