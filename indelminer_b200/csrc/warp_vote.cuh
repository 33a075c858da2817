// Warp-level building blocks of the realignment kernels (sm_100a):
//   * per-warp shared-memory layout (one read is realigned by one warp, no CTA barriers)
//   * TMA bulk copies (cp.async.bulk + mbarrier) that stage the packed reference window and the
//     read bytes of the NEXT read while the current one is being realigned
//   * find_best_band (src/alignment.c:393-447 with :29-181) as a warp-wide k-mer diagonal vote
#pragma once

#include "kernels.cuh"

namespace indelgpu {

// ---------------------------------------------------------------------------------------
// per-warp shared-memory slice, computed identically on host and device
// ---------------------------------------------------------------------------------------
struct WarpLayout {
    int indexed;         // 1: the vote reads the resident k-mer index of the reference (kmer_index.cuh): no table, no staged window
    int direct;          // 1: direct-address table of 4^k uint16 entries; 0: open-addressing hash
    int hash_slots;      // power of two (hash only)
    int hist_bits;       // 8 or 16 bits per diagonal counter
    int hist_words;      // uint32 words of the histogram (multiple of 4)
    int hist_cap;        // diagonals the histogram can hold
    int win_bytes;       // bytes of one staged packed window (multiple of 16)
    int read_bytes;      // bytes of one staged read (multiple of 16, includes 16 bytes of misalignment)
    int ascii_bytes;     // 4-bit batches: the read expanded to ASCII (0 when the batch is ASCII already)
    int pk_words;        // packed read words
    int ops_cap;         // words per CIGAR buffer
    int off_tab, off_hist, off_win0, off_win1, off_read0, off_read1, off_pk, off_psum, off_bits,
        off_cig1, off_cig2, off_segs, off_bar, off_misc, off_list, off_ascii;
    int total;           // bytes per warp (multiple of 128)
};

__host__ __device__ inline int cigar_cap(const DevParams& P, int max_read, int banded)
{
    // A positive-scoring ungapped local segment with x mismatches has x * |mismatch| < (M - x) * match,
    // so at most 2x + 1 runs; + 2 soft clips + 1 for the clip round 2 may insert (alignment.c:478-532).
    if (!banded && P.match > 0 && P.mismatch < 0) {
        const long long x = ((long long)max_read * P.match) / ((long long)P.match - P.mismatch);
        const long long cap = 2 * x + 8;
        if (cap < max_read + 4) return (int)cap;
    }
    return max_read + 4;
}

// max_numdiag bounds the STAGED window (window 2, alignment.c:780-783) plus the read; max_votediag bounds
// the diagonals one vote can address.  The two differ: round 1 votes on window 1 (2 * range1 bases) and
// every round-2 window lies on one side of the anchor (alignment.c:606-706: at most range1 + maxdelsize).
__host__ __device__ inline WarpLayout make_warp_layout(const DevParams& P, int max_read, int max_numdiag, int max_votediag, int banded,
                                                       int indexed = 0, int packed4 = 0)
{
    WarpLayout L;
    L.indexed = indexed && P.k <= 6;
    L.direct = P.k <= 6;
    int hs = 256;
    #pragma unroll 1
    while (hs < 4 * max_read) hs <<= 1;
    L.hash_slots = L.direct ? 0 : hs;
    L.hist_bits = (max_read - P.k + 1 <= 255 && max_votediag <= 65535) ? 8 : 16;
    L.hist_cap = max_votediag;
    // indexed: two bitmaps over the 4^k codes ("seen in the slice", "seen more than once")
    const int tab_bytes = L.indexed ? 2 * (((1 << (2 * P.k)) + 31) / 32 * 4) : L.direct ? ((L.hist_bits / 8) << (2 * P.k)) : hs * 8;
    L.hist_words = round_up((max_votediag + 4) * (L.hist_bits / 8), 16) / 4;
    L.win_bytes = L.indexed ? 0 : round_up(max_numdiag / 4 + 64, 16);   // window <= max_numdiag bases, 64-base aligned start, hi word
    L.read_bytes = round_up(max_read + 32, 16);
    L.ascii_bytes = packed4 ? round_up(max_read + 16, 16) : 0;
    L.pk_words = max_read / 16 + 3;
    L.ops_cap = cigar_cap(P, max_read, banded);
    int o = 0;
    L.off_tab = o;   o += round_up(tab_bytes < 16 ? 16 : tab_bytes, 16);
    L.off_hist = o;  o += L.hist_words * 4;
    L.off_win0 = o;  o += L.win_bytes;
    L.off_win1 = o;  o += L.win_bytes;
    L.off_read0 = o; o += L.read_bytes;
    L.off_read1 = o; o += L.read_bytes;
    L.off_ascii = o; o += L.ascii_bytes;
    L.off_pk = o;    o += round_up(L.pk_words * 4, 16);
    L.off_cig1 = o;  o += round_up(L.ops_cap * 4, 16);
    L.off_cig2 = o;  o += round_up(L.ops_cap * 4, 16);
    L.off_bar = o;   o += 16;                                      // two mbarriers
    L.off_misc = o;  o += 256;                                     // Aln x 2, Plan, scalars
    // the vote's hit list and the alignment's prefix sums / match bits are never live together: one region
    {
        // kHitListCap entries (window scan) or kIdxListHits entries (kmer_index.cuh)
        const int list_bytes = round_up((L.indexed ? 160 : 32 + 256 + 32) * (L.hist_bits == 8 ? 2 : 4), 16);
        const int psum_bytes = round_up((max_read + 2) * 4, 16);
        const int bits_bytes = round_up(((max_read + 127) / 128 * 4 + 4) * 4, 16);
        // ... and the stitched segment words are written after the last alignment and flushed before the next read's
        // first vote (realign_kernel's pend_* block): the same region again
        const int segs_bytes = round_up((2 * L.ops_cap + 4) * 4, 16);
        L.off_list = o; L.off_psum = o; L.off_bits = o + psum_bytes; L.off_segs = o;
        int region = list_bytes > psum_bytes + bits_bytes ? list_bytes : psum_bytes + bits_bytes;
        if (segs_bytes > region) region = segs_bytes;
        o += region;
    }
    L.total = round_up(o, 128);
    return L;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded wait: a copy that never lands is a bug, not something to hang the GPU on
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    #pragma unroll 1
    for (int spin = 0; spin < (1 << 24); spin++) {
        uint32_t ok;
        // the time hint lets the hardware suspend the warp until the phase completes instead of spinning on issue slots
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
        if (ok) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// per-warp views
// ---------------------------------------------------------------------------------------
struct WarpView;
__device__ __forceinline__ uint32_t* win_buf(const WarpView& V, int b);
__device__ __forceinline__ uint8_t* read_buf(const WarpView& V, int b);

struct WarpView {
    WarpLayout L;
    uint16_t* tab16; uint32_t* keys; uint32_t* vals;
    uint32_t* hist;
    uint8_t* ascii;                      // 4-bit batches: the current read as ASCII
    uint32_t* win0; uint8_t* rbuf0;      // double buffers: buffer b at win0 + b * win_bytes / rbuf0 + b * read_bytes
    uint32_t* pk;
    uint64_t* bar;        // [2]
    int* misc;
    void* list;           // hit list of the vote (HitIdx<HB>::type[kHitListCap])
    Cta S;                // psum / bits / cig1 / cig2 / segs views used by the scalar pieces (kernels.cuh)
};

__device__ __forceinline__ uint32_t* win_buf(const WarpView& V, int b) { return V.win0 + b * (V.L.win_bytes / 4); }
__device__ __forceinline__ uint8_t* read_buf(const WarpView& V, int b) { return V.rbuf0 + b * V.L.read_bytes; }

__device__ __forceinline__ void bind_warp(WarpView& V, unsigned char* base, const WarpLayout& L)
{
    V.L = L;
    V.tab16 = reinterpret_cast<uint16_t*>(base + L.off_tab);
    V.keys = reinterpret_cast<uint32_t*>(base + L.off_tab);
    V.vals = V.keys + L.hash_slots;
    V.hist = reinterpret_cast<uint32_t*>(base + L.off_hist);
    V.win0 = reinterpret_cast<uint32_t*>(base + L.off_win0);
    V.rbuf0 = base + L.off_read0;
    V.ascii = base + L.off_ascii;
    V.pk = reinterpret_cast<uint32_t*>(base + L.off_pk);
    V.bar = reinterpret_cast<uint64_t*>(base + L.off_bar);
    V.misc = reinterpret_cast<int*>(base + L.off_misc);
    V.list = base + L.off_list;
    V.S.keys = nullptr; V.S.vals = nullptr; V.S.hist = nullptr;
    V.S.read = nullptr;
    V.S.bits = reinterpret_cast<uint32_t*>(base + L.off_bits);
    V.S.psum = reinterpret_cast<int*>(base + L.off_psum);
    V.S.cig1 = reinterpret_cast<uint32_t*>(base + L.off_cig1);
    V.S.cig2 = reinterpret_cast<uint32_t*>(base + L.off_cig2);
    V.S.segs = reinterpret_cast<uint32_t*>(base + L.off_segs);
    V.S.L.ops_cap = L.ops_cap;
}

// zero the table and the histogram once; every vote leaves both clean again
__device__ __forceinline__ void init_warp_tables(WarpView& V)
{
    const int lane = threadIdx.x & 31;
    const int tabw = (V.L.off_hist - V.L.off_tab) / 4;
    uint32_t* t = reinterpret_cast<uint32_t*>(V.tab16);
    #pragma unroll 1
    for (int s = lane; s < tabw; s += 32) t[s] = V.L.direct ? 0u : kEmptyKey;
    if (!V.L.direct) for (int s = lane; s < V.L.hash_slots; s += 32) V.vals[s] = 0;
    #pragma unroll 1
    for (int s = lane; s < V.L.hist_words; s += 32) V.hist[s] = 0;
    __syncwarp();
}

// 2-bit pack of the staged read (codes of base_code; bases 16 per word, base j of a word in bits 2j..2j+1)
__device__ __forceinline__ void pack_read_warp(WarpView& V, const uint8_t* read, int M)
{
    const int lane = threadIdx.x & 31;
    const int chunks = (M + 31) >> 5;
    #pragma unroll 1
    for (int c = 0; c < chunks; c++) {
        const int i = (c << 5) + lane;
        const uint32_t q = i < M ? base_code(read[i]) : 0u;
        const uint32_t v = q << (2 * (lane & 15));
        const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? v : 0u);
        const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? 0u : v);
        if (lane == 0) { V.pk[2 * c] = lo; V.pk[2 * c + 1] = hi; }
    }
    if (lane == 0) { V.pk[2 * chunks] = 0; V.pk[2 * chunks + 1] = 0; }
    __syncwarp();
}

// BAM 4-bit bases (bam1_seqi: high nibble first) -> ASCII as bit2char does (readaln.c:4-17: 1 A, 2 C, 4 G, 8 T, 15 N),
// written back to front and complemented when the batch entry asks for the reverse complement
// (reverse_complement_string, sequences.c:204-220, on A C G T N).  Returns false on a code bit2char stops the program on.
__device__ __forceinline__ bool expand_read4_warp(uint8_t* out, const uint8_t* nib, int L, bool rc)
{
    const int lane = threadIdx.x & 31;
    // byte c of the table = the base of code c (0 = invalid)
    const unsigned long long f_lo = 0x0000004700434100ULL, f_hi = 0x4E00000000000054ULL;     // . A C . G . . .   T . . . . . . N
    const unsigned long long r_lo = 0x0000004300475400ULL, r_hi = 0x4E00000000000041ULL;     // . T G . C . . .   A . . . . . . N
    bool ok = true;
    #pragma unroll 1
    for (int j = lane; 2 * j < L; j += 32) {
        const uint32_t byte = nib[j];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int t = 2 * j + h;
            const uint32_t code = h == 0 ? (byte >> 4) : (byte & 15u);
            const unsigned long long tab = rc ? (code < 8u ? r_lo : r_hi) : (code < 8u ? f_lo : f_hi);
            uint32_t ch = (uint32_t)(tab >> (8 * (code & 7u))) & 0xFFu;
            if (t < L) {
                if (ch == 0u) { ok = false; ch = 'N'; }
                out[rc ? L - 1 - t : t] = (uint8_t)ch;
            }
        }
    }
    __syncwarp();
    return __all_sync(0xFFFFFFFFu, ok);
}

__device__ __forceinline__ uint32_t kmer_at(const uint32_t* pk, int i, uint32_t kmask)
{
    const int w = i >> 4;
    return __funnelshift_r(pk[w], pk[w + 1], 2 * (i & 15)) & kmask;
}

// element type of the per-warp hit list: diagonal indices fit 16 bits whenever HB == 8 (make_warp_layout)
template <int HB> struct HitIdx { typedef uint32_t type; };
template <> struct HitIdx<8> { typedef uint16_t type; };
constexpr int kHitListCap = 32 + 256 + 32;     // carried remainder + at most 256 hits appended at a time (a step with more is split in two)

// table entry: offset + 1 of the read k-mer if it occurs exactly once in the slice, else 0.
// Direct tables hold one entry per possible k-mer, HB/8 bytes wide (HB == 8 iff the slice has at most
// 255 k-mers, so the same width serves the table and the histogram counters).
template <bool DIRECT, int HB>
__device__ __forceinline__ uint32_t kmer_lookup(const WarpView& V, uint32_t code)
{
    if (DIRECT) {
        if (HB == 8) return reinterpret_cast<const uint8_t*>(V.tab16)[code];
        return V.tab16[code];
    }
    const int hm = V.L.hash_slots - 1;
    uint32_t slot = hash_slot(code, hm);
    #pragma unroll 1
    while (true) {
        const uint32_t kk = V.keys[slot];
        if (kk == code) { const uint32_t v = V.vals[slot]; return (v >> 16) == 1u ? (v & 0xFFFFu) : 0u; }
        if (kk == kEmptyKey) return 0u;
        slot = (slot + 1) & hm;
    }
}

template <int HB>
__device__ __forceinline__ void tab_store(WarpView& V, uint32_t code, uint32_t v)
{
    if (HB == 8) reinterpret_cast<uint8_t*>(V.tab16)[code] = (uint8_t)v;
    else V.tab16[code] = (uint16_t)v;
}

// ---------------------------------------------------------------------------------------
// pieces shared by the two votes (window scan below, reference index in kmer_index.cuh)
// ---------------------------------------------------------------------------------------

// Pass C: n (<= 32) entries of the hit list, one per lane: equal neighbours are merged by ballot and the leader
// of a run issues ONE shared-memory atomic for the whole run.  For one-diagonal bands select_band
// (alignment.c:142-181) rides along: the atomic that returns the largest count ever seen is the LAST vote of
// its diagonal, so every lane keeps the largest count its own atomics produced and, among those, the smallest
// tie key 2 * |a - i| + (i > a) (nearest to a, the lower index on equal distance).
template <int HB>
__device__ __forceinline__ void vote_hits_chunk(WarpView& V, const typename HitIdx<HB>::type* list, int n, int a,
                                                uint32_t& lbest, uint32_t& lkey)
{
    const int lane = threadIdx.x & 31;
    const bool valid = lane < n;
    const int idx = valid ? (int)list[lane] : -1;
    const int prev = __shfl_up_sync(0xFFFFFFFFu, idx, 1);
    const bool leader = valid && (lane == 0 || idx != prev);
    const uint32_t leaders = __ballot_sync(0xFFFFFFFFu, leader);
    if (leader) {
        const uint32_t rest = (lane == 31) ? 0u : (leaders >> (lane + 1));
        const uint32_t cnt = (uint32_t)(rest ? __ffs(rest) : n - lane);
        constexpr int LG = (HB == 8) ? 2 : 1;
        const int sh = (idx & ((1 << LG) - 1)) * HB;
        const uint32_t old = atomicAdd(&V.hist[idx >> LG], cnt << sh);
        const uint32_t now = ((old >> sh) & ((1u << HB) - 1u)) + cnt;
        const uint32_t key = 2u * (uint32_t)(a > idx ? a - idx : idx - a) + (idx > a ? 1u : 0u);
        if (now > lbest) { lbest = now; lkey = key; }
        else if (now == lbest) lkey = min(lkey, key);
    }
}

// bin_bands + select_band (alignment.c:130-181) on the finished histogram, which is left zeroed.
// a = anchor_rel clamped to [-1, numdiag]; lbest / lkey from vote_hits_chunk (used when g == 0).
template <int HB>
__device__ __forceinline__ int select_band_warp(WarpView& V, int numdiag, int g, int a, uint32_t lbest, uint32_t lkey)
{
    const int lane = threadIdx.x & 31;
    int idx;
    constexpr int PER = 32 / HB;                                  // counters per word
    if (g == 0) {
        const uint32_t cmax = __reduce_max_sync(0xFFFFFFFFu, lbest);
        if (cmax == 0u) {                                         // no vote at all: the index nearest to a
            idx = a < 0 ? 0 : (a > numdiag - 1 ? numdiag - 1 : a);
        } else {
            const uint32_t bestkey = __reduce_min_sync(0xFFFFFFFFu, lbest == cmax ? lkey : 0xFFFFFFFFu);
            idx = (bestkey & 1u) ? a + (int)(bestkey >> 1) : a - (int)(bestkey >> 1);
            const int nq = ((numdiag + PER - 1) / PER + 3) / 4;   // uint4 chunks
            uint4* h4 = reinterpret_cast<uint4*>(V.hist);
            #pragma unroll 1
            for (int q = lane; q < nq; q += 32) h4[q] = make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
        unsigned long long best = 0;
        #pragma unroll 1
        for (int i = lane; i < numdiag; i += 32) {
            uint32_t b = 0;
            if (i < numdiag - g)
                #pragma unroll 1
                for (int j = 0; j <= g; j++) {
                    const int t = i + j;
                    b += (HB == 8) ? ((V.hist[t >> 2] >> ((t & 3) * 8)) & 0xFFu) : ((V.hist[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu);
                }
            const uint32_t dist = (uint32_t)(a > i ? a - i : i - a);
            const unsigned long long key = ((unsigned long long)b << 42) |
                                           ((unsigned long long)(0x1FFFFFu - dist) << 21) |
                                           (unsigned long long)(0x1FFFFFu - (uint32_t)i);
            best = key > best ? key : best;
        }
        best = warp_max_u64(best);
        __syncwarp();
        #pragma unroll 1
        for (int s = lane; s < (numdiag + PER - 1) / PER + 1 && s < V.L.hist_words; s += 32) V.hist[s] = 0;
        idx = (int)(0x1FFFFFu - (uint32_t)(best & 0x1FFFFFu));
    }
    return idx;
}

// ---------------------------------------------------------------------------------------
// find_best_band for one (window, read slice) pair, executed by one warp.
//   swin / sw0  staged packed window: swin[w - sw0] is packed word w of the reference
//   wabs, N     absolute base offset (in packed-reference coordinates) and length of the window
//   zs2, M      read slice inside the packed read V.pk
// Restated set-wise (SURVEY.md 8a''): for every window offset j whose k-mer equals a k-mer that occurs
// exactly once in the slice (at offset i): diag[j - i + (M-k+1)]++ ; then the arg-max band with the
// reference's tie rule (count desc, |anchor_rel - index| asc, index asc).
// Returns low (up = low + g); *ok false when the reference would have aborted (alignment.c:405).
// ---------------------------------------------------------------------------------------
template <bool DIRECT, int HB>
__device__ __forceinline__ int vote_band_warp(const DevParams& P, WarpView& V, const uint32_t* swin, int64_t sw0,
                              int64_t wabs, int N, int zs2, int M, int anchor_rel, bool* ok)
{
    const int lane = threadIdx.x & 31;
    const int k = P.k, g = P.g;
    const int numdiag = (N - (k - 1)) + (M - (k - 1));            // alignment.c:403-404
    *ok = numdiag > g && numdiag <= V.L.hist_cap;                 // the second clause is this kernel's limit (never hit for
    if (!*ok) return 0;                                           // windows derived from a batch entry; see make_warp_layout)
    if (M < k) return numdiag - 1;                                // alignment.c:408-412
    const uint32_t kmask = P.kmask;
    const int nk = M - k + 1;

    // 1. index the k-mers of the slice.  Direct table, no atomics: everybody stores its offset, then whoever
    //    does not find its own offset back knows the code is shared and zeroes the entry.
    if (DIRECT) {
        #pragma unroll 1
        for (int i = lane; i < nk; i += 32) tab_store<HB>(V, kmer_at(V.pk, zs2 + i, kmask), (uint32_t)(i + 1));
        __syncwarp();
        #pragma unroll 1
        for (int i = lane; i < nk; i += 32) {
            const uint32_t c = kmer_at(V.pk, zs2 + i, kmask);
            if (kmer_lookup<true, HB>(V, c) != (uint32_t)(i + 1)) tab_store<HB>(V, c, 0u);
        }
    } else {
        const int hm = V.L.hash_slots - 1;
        #pragma unroll 1
        for (int i = lane; i < nk; i += 32) {
            const uint32_t code = kmer_at(V.pk, zs2 + i, kmask);
            uint32_t slot = hash_slot(code, hm);
            #pragma unroll 1
            while (true) {
                const uint32_t prev = atomicCAS(&V.keys[slot], kEmptyKey, code);
                if (prev == kEmptyKey || prev == code) { atomicAdd(&V.vals[slot], (1u << 16) | (uint32_t)(i + 1)); break; }
                slot = (slot + 1) & hm;
            }
        }
    }
    __syncwarp();

    // 2. scan the window: one packed word (16 k-mer starts) per lane per step.
    //    Pass A is branch-free: which of the 16 positions hit a unique read k-mer.  Pass B appends the
    //    diagonal index of every hit to a per-warp list (a lane's hits go to consecutive slots, lanes in
    //    window order, so the votes of one diagonal -- the true alignment -- are neighbours).  Pass C
    //    drains the list 32 entries at a time with all lanes busy: equal neighbours are merged by ballot
    //    and their leader issues one shared-memory atomic for the run.
    //    For one-diagonal bands (g == 0) select_band (alignment.c:142-181) rides along: an atomic that
    //    returns the largest count ever seen is the LAST vote of its diagonal, so every lane keeps the
    //    largest count its own atomics produced and, among those, the smallest tie key
    //    2 * |a - i| + (i > a)   (nearest to a, the lower index on equal distance).
    int a = anchor_rel;
    a = a < -1 ? -1 : (a > numdiag ? numdiag : a);                // clamping keeps every comparison
    uint32_t lbest = 0, lkey = 0xFFFFFFFFu;
    if (N >= k) {
        typedef typename HitIdx<HB>::type hit_t;
        hit_t* list = reinterpret_cast<hit_t*>(V.list);
        const int fr = (int)(wabs - (sw0 << 4));                  // first k-mer start, relative to the staged buffer
        const int lr = fr + N - k;                                // last k-mer start, inclusive
        const int w0 = fr >> 4, w1 = lr >> 4;
        const int shiftM = M - k + 1;
        int fill = 0;                                             // entries waiting in the list (warp-uniform)
#pragma unroll 1
        for (int wb = w0; wb <= w1 || fill > 0; wb += 32) {
            const int wi = wb + lane;
            if (wb <= w1) {
                uint32_t hits = 0, lo = 0, hi = 0;
                const int rel0 = (wi << 4) - fr;                  // window offset of position 0 of this word
                if (wi <= w1) {
                    lo = swin[wi]; hi = swin[wi + 1];
#pragma unroll
                    for (int p = 0; p < 16; p++) {
                        const uint32_t code = __funnelshift_r(lo, hi, 2 * p) & kmask;
                        if (kmer_lookup<DIRECT, HB>(V, code) != 0u) hits |= 1u << p;
                    }
                    const int plo = rel0 < 0 ? -rel0 : 0;
                    const int phi = (wi == w1) ? (lr - (wi << 4)) : 15;
                    hits &= (0xFFFFu << plo) & (0xFFFFu >> (15 - phi));
                }
                // pass B, then pass C on the full chunks.  A step with more than 256 hits (repeat-rich windows only) is
                // appended in two halves -- positions 0-7 of every lane, then 8-15 -- so that the list never holds more
                // than 31 + 256 entries
                uint32_t rest = hits; bool low_half = false;
#pragma unroll 1
                do {
                    const uint32_t take = low_half ? (rest & 0x00FFu) : rest;
                    const int mine = __popc(take);
                    int pre = mine;                               // inclusive scan of the lanes' hit counts
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, pre, o); if (lane >= o) pre += t; }
                    const int total = __shfl_sync(0xFFFFFFFFu, pre, 31);
                    if (total > 256) { low_half = true; continue; }
                    int slot = fill + pre - mine;
                    uint32_t h = take;
#pragma unroll 1
                    while (h) {
                        const int p = __ffs(h) - 1;
                        h &= h - 1;
                        const uint32_t off = kmer_lookup<DIRECT, HB>(V, __funnelshift_r(lo, hi, 2 * p) & kmask);
                        list[slot++] = (hit_t)(rel0 + p - (int)(off - 1u) + shiftM);
                    }
                    fill += total;
                    rest &= ~take; low_half = false;
                    __syncwarp();
                    if (__any_sync(0xFFFFFFFFu, rest != 0u)) {       // more to come: make room first
                        int dn = 0;
#pragma unroll 1
                        for (; fill - dn >= 32; dn += 32) vote_hits_chunk<HB>(V, list + dn, 32, a, lbest, lkey);
                        const int rem = fill - dn;
                        const hit_t v = (lane < rem) ? list[dn + lane] : (hit_t)0;
                        __syncwarp();
                        if (lane < rem) list[lane] = v;
                        fill = rem;
                        __syncwarp();
                    }
                } while (__any_sync(0xFFFFFFFFu, rest != 0u));
            }
            // pass C: full chunks of 32; the last, partial chunk once the window is exhausted
            int done = 0;
#pragma unroll 1
            while (fill - done >= 32 || (wb + 32 > w1 && fill - done > 0)) {
                const int n = min(32, fill - done);
                vote_hits_chunk<HB>(V, list + done, n, a, lbest, lkey);
                done += n;
            }
            if (done) {                                           // carry the remainder (< 32 entries) to the front
                const int rem = fill - done;
                const hit_t v = (lane < rem) ? list[done + lane] : (hit_t)0;
                __syncwarp();
                if (lane < rem) list[lane] = v;
                fill = rem;
                __syncwarp();
            }
        }
    }
    __syncwarp();

    // 3. bin_bands + select_band (alignment.c:130-181) and zeroing of the histogram
    const int idx = select_band_warp<HB>(V, numdiag, g, a, lbest, lkey);

    // 4. leave the table clean for the next vote
    if (DIRECT) {
        #pragma unroll 1
        for (int i = lane; i < nk; i += 32) tab_store<HB>(V, kmer_at(V.pk, zs2 + i, kmask), 0u);
    } else {
        #pragma unroll 1
        for (int s = lane; s < V.L.hash_slots; s += 32) { V.keys[s] = kEmptyKey; V.vals[s] = 0; }
    }
    __syncwarp();
    return idx - (M - k + 1);                                     // alignment.c:438
}

// Stage `bytes` (multiple of 16) of a packed window starting at packed word sw0 into smem with one
// bulk copy; returns the byte count it registered on the barrier.
__device__ __forceinline__ uint32_t window_span_bytes(int64_t first_base, int64_t end_base, int64_t* sw0)
{
    const int64_t a0 = first_base & ~(int64_t)63;                 // 64 bases = 16 bytes of packed reference
    *sw0 = a0 >> 4;
    const int64_t lastw = (end_base >> 4) + 1;                    // hi word of the last k-mer start
    const int64_t words = lastw - *sw0 + 1;
    return (uint32_t)(((words * 4 + 15) / 16) * 16);
}
#endif  // __CUDACC__

}  // namespace indelgpu
