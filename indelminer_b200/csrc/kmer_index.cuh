// Resident k-mer index of the reference, and find_best_band (src/alignment.c:393-447 with :29-181) as a
// warp-wide vote that reads it.
//
// The window-scan vote (warp_vote.cuh) looks up ALL N ~ 1.4-3.4 k window positions in a table of the
// read's k-mers, for every read and both rounds, although candidate reads lie ~64 bases apart and scan
// nearly the same windows.  That work is per reference position, not per read, so it is done ONCE, when
// the reference is uploaded: the reference is cut into blocks of kIdxBlock = 4096 bases and every block
// gets a CSR of its k-mer starts --
//     idx_rec[block * 4^k + code] = 8 bytes: count, and the block offsets of up to three occurrences INLINE
//                                   (count <= 3, 98 % of the buckets at k = 6: one 64-bit load answers the look-up);
//                                   for count > 3 the first slot is an index into
//     idx_pos[block * 4096 + first .. + count)                   = offsets inside the block (uint16), bucket by bucket
// 8 * 4^k + 2 * 4096 bytes per block = 10 bytes per base at k = 6 (0.64 GB for a 64 Mb contig, 31 GB for a
// 3.1 Gb genome: HBM is 180 GB).  The k-mer of a position is taken from the same 2-bit packed copy the
// window scan reads (non-ACGT -> the code of A, base2bits alignment.c:11-24), so both votes see the same
// k-mers.
//
// A vote then turns the loop of bin_diagonals (alignment.c:70-128) around: for each of the read slice's
// M-k+1 k-mers that is unique in the slice (:97-98) fetch its bucket in the one or two blocks the window
// overlaps -- about one position per bucket -- keep the positions inside the window, and vote for
// diagonal (j - i + M - k + 1).  The records of the next 32 k-mers are in flight while the current 32 are
// voted on, so a vote exposes about one L2 round trip.  ~150 bucket look-ups per vote instead of ~1 500 table look-ups, no
// per-warp table or staged window in shared memory (10.4 KB -> 5.5 KB per warp: more resident warps).
#pragma once

#include "warp_vote.cuh"

namespace indelgpu {

constexpr int kIdxBlockLog = 12;
constexpr int kIdxBlock = 1 << kIdxBlockLog;          // reference bases per index block
constexpr int kIdxMaxK = 6;                            // 4^k bucket words per block
constexpr int kIdxListHits = 160;                     // hit list: < 32 carried + at most 96 per (32 k-mers, block) pair + 32 per big-bucket step

struct KmerIndex {
    const uint2* rec;         // [nblocks << 2k]  x = count | slot0 << 16, y = slot1 | slot2 << 16
    const uint16_t* pos;      // [nblocks * kIdxBlock]
    int64_t nblocks;
    int k;                    // the k it was built for (0 = none)
};

#ifdef __CUDACC__
// One CTA of 256 threads per block of the reference: thread t owns packed word t of the block (16 k-mer
// starts).  Count, exclusive scan, scatter -- all in shared memory -- then one coalesced write-out.
__global__ void __launch_bounds__(256)
build_kmer_index_kernel(const uint32_t* __restrict__ packed, int64_t nblocks, int k, uint32_t kmask,
                        uint2* __restrict__ idx_rec, uint16_t* __restrict__ idx_pos)
{
    __shared__ uint32_t s_cnt[1 << (2 * kIdxMaxK)];     // counts, then write cursors
    __shared__ uint32_t s_start[1 << (2 * kIdxMaxK)];
    __shared__ uint16_t s_pos[kIdxBlock];
    __shared__ uint32_t s_warp[8];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nb = 1 << (2 * k);
    #pragma unroll 1
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        for (int c = t; c < nb; c += 256) s_cnt[c] = 0;
        __syncthreads();
        const int64_t w = blk * (kIdxBlock / 16) + t;
        const uint32_t lo = packed[w], hi = packed[w + 1];
        uint32_t codes[16];
#pragma unroll
        for (int p = 0; p < 16; p++) { codes[p] = __funnelshift_r(lo, hi, 2 * p) & kmask; atomicAdd(&s_cnt[codes[p]], 1u); }
        __syncthreads();
        // exclusive scan of the nb counters: `per` consecutive counters per thread
        const int per = nb >= 256 ? nb / 256 : 1;
        uint32_t local = 0;
        if (t * per < nb) for (int j = 0; j < per; j++) local += s_cnt[t * per + j];
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int q = 0; q < warp; q++) base += s_warp[q];
        uint32_t run = base + incl - local;
        if (t * per < nb)
            for (int j = 0; j < per; j++) {
                const int c = t * per + j;
                const uint32_t n = s_cnt[c];
                s_start[c] = run | (n << 16);
                s_cnt[c] = run;                          // cursor
                run += n;
            }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < 16; p++) {
            const uint32_t slot = atomicAdd(&s_cnt[codes[p]], 1u);
            s_pos[slot] = (uint16_t)(t * 16 + p);
        }
        __syncthreads();
        for (int c = t; c < nb; c += 256) {
            const uint32_t st = s_start[c], first = st & 0xFFFFu, n = st >> 16;
            uint32_t s0 = first, s1 = 0, s2 = 0;                 // n > 3: slot 0 points into the block's position list
            if (n <= 3) { s0 = n > 0 ? s_pos[first] : 0u; s1 = n > 1 ? s_pos[first + 1] : 0u; s2 = n > 2 ? s_pos[first + 2] : 0u; }
            idx_rec[blk * nb + c] = make_uint2(n | (s0 << 16), s1 | (s2 << 16));
        }
        for (int q = t; q < kIdxBlock / 2; q += 256)
            reinterpret_cast<uint32_t*>(idx_pos + blk * kIdxBlock)[q] = reinterpret_cast<const uint32_t*>(s_pos)[q];
        __syncthreads();
    }
}

// find_best_band for one (window, read slice) pair by one warp, reading the index.
//   g0, N       global base coordinate (contig offset + zstart1) and length of the window
//   zs2, M      read slice inside the packed read V.pk
// Same contract as vote_band_warp: returns low (up = low + g); *ok false when the reference would abort.
template <int HB>
__device__ __forceinline__ int vote_band_index(const DevParams& P, WarpView& V, const KmerIndex& X,
                                               int64_t g0, int N, int zs2, int M, int anchor_rel, bool* ok)
{
    const int lane = threadIdx.x & 31;
    const int k = P.k, g = P.g;
    const int numdiag = (N - (k - 1)) + (M - (k - 1));            // alignment.c:403-404
    *ok = numdiag > g && numdiag <= V.L.hist_cap;
    if (!*ok) return 0;
    if (M < k) return numdiag - 1;                                // alignment.c:408-412
    const uint32_t kmask = P.kmask;
    const int nk = M - k + 1;
    uint32_t* seen = reinterpret_cast<uint32_t*>(V.tab16);
    uint32_t* dup = seen + (((1 << (2 * k)) + 31) >> 5);

    // 1. which k-mers occur more than once in the slice (alignment.c:97-98): two bitmaps, no table
    #pragma unroll 1
    for (int i = lane; i < nk; i += 32) {
        const uint32_t c = kmer_at(V.pk, zs2 + i, kmask);
        const uint32_t bit = 1u << (c & 31);
        const uint32_t old = atomicOr(&seen[c >> 5], bit);
        if (old & bit) atomicOr(&dup[c >> 5], bit);
    }
    __syncwarp();

    int a = anchor_rel;
    a = a < -1 ? -1 : (a > numdiag ? numdiag : a);
    uint32_t lbest = 0, lkey = 0xFFFFFFFFu;
    if (N >= k) {
        typedef typename HitIdx<HB>::type hit_t;
        hit_t* list = reinterpret_cast<hit_t*>(V.list);
        const int shiftM = M - k + 1;
        const uint32_t jmax = (uint32_t)(N - k);                  // last k-mer start inside the window
        const int64_t b0 = g0 >> kIdxBlockLog, b1 = (g0 + N - k) >> kIdxBlockLog;
        const int nbl = 2 * k;
        const uint32_t lt = (1u << lane) - 1u;
        int fill = 0;                                             // hits waiting in the list (warp-uniform)
        const int nblk = (int)(b1 - b0) + 1;
        // this lane's record for k-mer offset ci + lane in block b0 + bb (count 0 when it has none); requested one
        // (32 k-mers, block) pair ahead of its use, so the look-ups of a vote expose about one L2 round trip in all
        auto request = [&](int ci, int bb) -> uint2 {
            const int i = ci + lane;
            if (i >= nk) return make_uint2(0u, 0u);
            const uint32_t c = kmer_at(V.pk, zs2 + i, kmask);
            if ((dup[c >> 5] >> (c & 31)) & 1u) return make_uint2(0u, 0u);      // not unique in the slice (alignment.c:97-98)
            return __ldg(X.rec + ((b0 + bb) << nbl) + c);
        };
        // pass C on full chunks of 32; what is left (< 32) moves to the front
        auto drain = [&]() {
            if (fill < 32) return;
            int head = 0;
            #pragma unroll 1
            for (; fill - head >= 32; head += 32) vote_hits_chunk<HB>(V, list + head, 32, a, lbest, lkey);
            const int rem = fill - head;
            const hit_t v = (lane < rem) ? list[head + lane] : (hit_t)0;
            __syncwarp();
            if (lane < rem) list[lane] = v;
            fill = rem;
            __syncwarp();
        };
        uint2 rnext = request(0, 0);
        #pragma unroll 1
        for (int ci = 0, bb = 0; ci < nk;) {
            const uint2 r = rnext;
            int nci = ci, nbb = bb + 1;                           // pairs in order: block fastest
            if (nbb == nblk) { nbb = 0; nci += 32; }
            rnext = request(nci, nbb);
            const int i = ci + lane;
            const int jrel = (int)(((b0 + bb) << kIdxBlockLog) - g0);          // window offset of the block's first base
            const int rel = jrel - i + shiftM;                    // its diagonal index for this lane's k-mer
            const uint32_t cnt = r.x & 0xFFFFu;
            const int s0 = (int)(r.x >> 16), s1 = (int)(r.y & 0xFFFFu), s2 = (int)(r.y >> 16);
            const bool big = cnt > 3u;                            // 2 % of the buckets at k = 6: walked below
            const bool h0 = !big && cnt > 0u && (uint32_t)(jrel + s0) <= jmax;   // 0 <= j <= N - k in one unsigned compare
            const bool h1 = !big && cnt > 1u && (uint32_t)(jrel + s1) <= jmax;
            const bool h2 = !big && cnt > 2u && (uint32_t)(jrel + s2) <= jmax;
            const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, h0), m1 = __ballot_sync(0xFFFFFFFFu, h1),
                           m2 = __ballot_sync(0xFFFFFFFFu, h2);
            uint32_t mb = __ballot_sync(0xFFFFFFFFu, big);
            // a lane's hits go to consecutive slots, lanes in k-mer order: the votes of the true diagonal are neighbours
            int slot = fill + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt);
            if (h0) list[slot++] = (hit_t)(rel + s0);
            if (h1) list[slot++] = (hit_t)(rel + s1);
            if (h2) list[slot] = (hit_t)(rel + s2);
            fill += __popc(m0) + __popc(m1) + __popc(m2);
            __syncwarp();
            // a bucket with more than three positions is read by the whole warp, 32 positions per step
            #pragma unroll 1
            while (mb) {
                const int src = __ffs(mb) - 1;
                mb &= mb - 1;
                const int bcnt = (int)__shfl_sync(0xFFFFFFFFu, cnt, src);
                const int bfirst = __shfl_sync(0xFFFFFFFFu, s0, src);
                const int brel = rel + lane - src;                // the owner's k-mer offset is ci + src
                const uint16_t* bp = X.pos + ((b0 + bb) << kIdxBlockLog) + bfirst;
                #pragma unroll 1
                for (int t0 = 0; t0 < bcnt; t0 += 32) {
                    bool hit = false; int p = 0;
                    if (t0 + lane < bcnt) { p = (int)__ldg(bp + t0 + lane); hit = (uint32_t)(jrel + p) <= jmax; }
                    const uint32_t mh = __ballot_sync(0xFFFFFFFFu, hit);
                    if (hit) list[fill + __popc(mh & lt)] = (hit_t)(brel + p);
                    fill += __popc(mh);
                    __syncwarp();
                    drain();
                }
            }
            drain();
            ci = nci; bb = nbb;
        }
        if (fill > 0) vote_hits_chunk<HB>(V, list, fill, a, lbest, lkey);
    }
    __syncwarp();

    // 3. bin_bands + select_band (alignment.c:130-181) and zeroing of the histogram
    const int idx = select_band_warp<HB>(V, numdiag, g, a, lbest, lkey);

    // 4. leave the bitmaps clean for the next vote
    #pragma unroll 1
    for (int i = lane; i < nk; i += 32) {
        const uint32_t c = kmer_at(V.pk, zs2 + i, kmask);
        seen[c >> 5] = 0u; dup[c >> 5] = 0u;
    }
    __syncwarp();
    return idx - (M - k + 1);                                     // alignment.c:438
}
#endif  // __CUDACC__

}  // namespace indelgpu
