// Banded affine local/global alignment for bands wider than one diagonal (-g N > 0):
//   local_align   src/localalign.c:15-196   forward sweep (end point), reverse sweep (start point)
//   ALIGN/align   src/globalalign.c:66-401  linear-space divide-and-conquer script
//   fetch_cigar   src/globalalign.c:507-604
// The script of ALIGN differs from a plain traceback whenever the plain traceback meets a tie
// (SURVEY.md 0.5), so the divide-and-conquer procedure itself is reproduced: same sweep, same
// crossing-point bookkeeping, same block order.  Recursion is an explicit frame stack; the work
// arrays are shared by all levels exactly as in the reference (a child only touches rows below
// the ones its parent still reads, globalalign.c:280).
#pragma once
#include <type_traits>

#include "kernels.cuh"

// IG_HD: the serial pieces are plain C++ and are also compiled for the host so that
// tests/host_harness can exercise this exact source without a GPU (test only; the library
// never calls them on the host).
#define IG_HD __host__ __device__

// resident CTAs per SM the thread-per-alignment kernels are compiled for (register cap 65536 / (128 * N))
#ifndef BAND_MIN_BLOCKS
#define BAND_MIN_BLOCKS 4
#endif

namespace indelgpu {

IG_HD inline int ig_min(int a, int b) { return a < b ? a : b; }
IG_HD inline int ig_max(int a, int b) { return a > b ? a : b; }

// global-memory scratch of the thread-per-alignment kernels: per warp 32 * stride ints, lane-interleaved
struct BandScratch {
    int* base;
    long long stride;        // ints per thread (alignment)
    int max_band;            // widest band the slice was sized for (after ALIGN's widening)
    int max_rows;            // longest read
};

__host__ __device__ inline long long band_scratch_ints(int max_band, int max_rows)
{
    // band region: cc dd cp dp (ALIGN), aliased by the two rolling (H, D) rows of local_align
    // rows region: rec[3] (crossing pointer + type per row and state) | fl (forward link + type) | script
    return 4LL * (max_band + 4) + 4LL * (max_rows + 2) + (2LL * max_rows + max_band + 16);
}


// int array view with an element stride: 1 = a plain array (one aligning lane per warp),
// 32 = lane-interleaved scratch (element i of lane l at base[i * 32 + l]) so that the 32 alignments a
// warp runs side by side -- one per thread -- touch one 128-byte line per access.
template <int STRIDE>
struct IArr {
    int* p;
    IG_HD int& operator[](int i) const { return p[i * STRIDE]; }           // |i * STRIDE| < 2^31 for every supported size
    IG_HD IArr operator+(long long k) const { return IArr{p + k * STRIDE}; }
};

struct DcFrame {
    int a, b;                // offsets of the "1-based" views into read / window
    int M, N, low, up;
    int tb, te;
    int stage, k, l, kt, rmid, t2, t3, pad;
};

template <int STRIDE>
struct DcCtx {
    const DevParams* P;
    const uint8_t* A;        // read slice, 0-based
    const uint8_t* B;        // window, 0-based
    IArr<STRIDE> cc, dd, cp, dp;
    // crossing records of align() (MP/MT, FP/FT of globalalign.c:33-36), a row index (>= -1) and a path type
    // (0..2) packed in one int each: the D&C chases these pointers through memory, so bytes are latency
    IArr<STRIDE> rec[3], fl;
    IArr<STRIDE> S; int ns; int last;
    int cells;
};
IG_HD inline int dc_pack(int ptr, int type) { return ((ptr + 1) << 2) | type; }
IG_HD inline int dc_ptr(int packed) { return (packed >> 2) - 1; }
IG_HD inline int dc_type(int packed) { return packed & 3; }

template <int STRIDE>
IG_HD inline void put_del(DcCtx<STRIDE>& x, int k)      // globalalign.c:40-46
{
    if (x.last < 0) { x.S[x.ns - 1] -= k; x.last = x.S[x.ns - 1]; }
    else { x.S[x.ns++] = -k; x.last = -k; }
}
template <int STRIDE>
IG_HD inline void put_ins(DcCtx<STRIDE>& x, int k)      // globalalign.c:48-54
{
    if (x.last > 0) { x.S[x.ns - 1] += k; x.last = x.S[x.ns - 1]; }
    else { x.S[x.ns++] = k; x.last = k; }
}
template <int STRIDE>
IG_HD inline void put_rep(DcCtx<STRIDE>& x) { x.S[x.ns++] = 0; x.last = 0; }

// One sweep of align() (globalalign.c:95-258): fills the crossing list, returns through f the
// first crossing row (k = r, or -1), its successor l and type kt.  a1/b1 index so that a1[1] is
// the first symbol.
template <int STRIDE>
IG_HD inline void dc_sweep(DcCtx<STRIDE>& x, DcFrame& f)
{
    const int g = x.P->G, h = x.P->H, m = g + h;
    const uint8_t* a1 = x.A + f.a - 1;
    const uint8_t* b1 = x.B + f.b - 1;
    const int M = f.M, N = f.N, low = f.low, up = f.up, tb = f.tb, te = f.te;
    const IArr<STRIDE> CC = x.cc, DD = x.dd, CP = x.cp, DP = x.dp;
    const int band = up - low + 1;
    const int midd = band / 2 + 1;
    const int rmid = low + midd - 1;
    int leftd = 1 - low, rightd = band;
    int IP = 0;
    if (leftd < midd) {
        for (int j = 0; j < midd; j++) CP[j] = DP[j] = -1;
        for (int j = midd; j <= rightd; j++) CP[j] = DP[j] = 0;
        x.rec[0][0] = x.rec[1][0] = x.rec[2][0] = dc_pack(-1, 0);
    } else if (leftd > midd) {
        const int fr = leftd - midd;
        for (int j = 0; j <= midd; j++) CP[j] = DP[j] = fr;
        for (int j = midd + 1; j <= rightd; j++) CP[j] = DP[j] = -1;
        x.rec[0][fr] = x.rec[1][fr] = x.rec[2][fr] = dc_pack(-1, 0);
    } else {
        for (int j = 0; j <= rightd; j++) CP[j] = DP[j] = 0;
        x.rec[0][0] = x.rec[1][0] = x.rec[2][0] = dc_pack(-1, 0);
    }
    CC[leftd] = 0;
    {
        int t = (tb == 2) ? 0 : -g;
        for (int j = leftd + 1; j <= rightd; j++) { t -= h; CC[j] = t; DD[j] = t - g; }
    }
    CC[rightd + 1] = DD[rightd + 1] = kNeg;
    DD[leftd] = (tb == 1) ? 0 : -g;
    CC[leftd - 1] = kNeg;

    int c = 0, d = 0, e = 0;
    for (int i = 1; i <= M; i++) {
        if (i > N - up) rightd--;
        if (leftd > 1) leftd--;
        const uint8_t ai = a1[i];
        {
            const int open = CC[leftd + 1] - m, ext = DD[leftd + 1] - h;
            if (open > ext) { d = open; DP[leftd] = CP[leftd + 1]; }
            else            { d = ext;  DP[leftd] = DP[leftd + 1]; }
            const int ib = leftd + low - 1 + i;
            c = open;
            if (ib > 0) c = CC[leftd] + (ai == b1[ib] ? x.P->match : x.P->mismatch);
            if (d > c || ib <= 0) { c = d; CP[leftd] = DP[leftd]; }
            e = c - g;
            DD[leftd] = d; CC[leftd] = c;
            IP = CP[leftd];
            if (leftd == midd) CP[leftd] = DP[leftd] = IP = i;
        }
        x.cells += rightd - leftd + 1;
        for (int q = leftd + 1; q <= rightd; q++) {
            const int sub = (ai == b1[q + low - 1 + i]) ? x.P->match : x.P->mismatch;
            if (q != midd) {
                int open = c - m; e -= h;
                if (open > e) { e = open; IP = CP[q - 1]; }
                open = CC[q + 1] - m; d = DD[q + 1] - h;
                if (open > d) { d = open; DP[q] = CP[q + 1]; }
                else          { DP[q] = DP[q + 1]; }
                c = CC[q] + sub;
                if (c < d || c < e) {
                    if (e > d) { c = e; CP[q] = IP; }
                    else       { c = d; CP[q] = DP[q]; }
                }
                CC[q] = c; DD[q] = d;
            } else {
                int open = c - m; e -= h;
                int mp0, mt0, mp1, mt1 = 2, mp2, mt2 = 1;
                if (open > e) { e = open; mp1 = CP[q - 1]; }
                else          { mp1 = IP; }
                open = CC[q + 1] - m; d = DD[q + 1] - h;
                if (open > d) { d = open; mp2 = CP[q + 1]; }
                else          { mp2 = DP[q + 1]; }
                c = CC[q] + sub;
                if (c < d || c < e) {
                    if (e > d) { c = e; mp0 = mp1; mt0 = 2; }
                    else       { c = d; mp0 = mp2; mt0 = 1; }
                } else { mp0 = i - 1; mt0 = 0; }
                if (c - g > e) { mp1 = mp0; mt1 = mt0; }
                if (c - g > d) { mp2 = mp0; mt2 = mt0; }
                x.rec[0][i] = dc_pack(mp0, mt0); x.rec[1][i] = dc_pack(mp1, mt1); x.rec[2][i] = dc_pack(mp2, mt2);
                CP[q] = DP[q] = IP = i;
                CC[q] = c; DD[q] = d;
            }
        }
    }
    int k, l;
    if (te == 1 && d + g > c)      { k = DP[rightd]; l = 2; }
    else if (te == 2 && e + g > c) { k = IP;         l = 1; }
    else                           { k = CP[rightd]; l = 0; }
    if (rmid > N - M) l = 2; else if (rmid < N - M) l = 1;
    int r = -1;
    while (k > -1) {
        x.fl[k] = dc_pack(r, l);
        r = k;
        const int nxt = x.rec[l][r];
        k = dc_ptr(nxt); l = dc_type(nxt);
    }
    f.rmid = rmid;
    f.k = r;
    if (r != -1) { const int v = x.fl[r]; f.l = dc_ptr(v); f.kt = dc_type(v); }
    f.t2 = up - rmid - 1; f.t3 = low - rmid + 1;
}

// dc_sweep for bands of at most W2 diagonals with the four band-wide arrays (CC, DD, CP, DP) in
// registers and the loop over the band unrolled.  One cell body serves every diagonal: the first cell
// of a row (globalalign.c:150-166) is the general cell with the horizontal inputs at -infinity (then
// e and IP come out of the next cell's "open" branch with exactly the values the reference's separate
// first-cell code assigns), and the mid-diagonal cell (:189-232) is the general cell plus its
// crossing records.  Cells outside [leftd, rightd] leave every register untouched.
template <int W2, int STRIDE>
IG_HD inline void dc_sweep_reg(DcCtx<STRIDE>& x, DcFrame& f)
{
    const int g = x.P->G, h = x.P->H, m = g + h;
    const uint8_t* a1 = x.A + f.a - 1;
    const uint8_t* b1 = x.B + f.b - 1;
    const int M = f.M, N = f.N, low = f.low, up = f.up, tb = f.tb, te = f.te;
    const int band = up - low + 1;
    const int midd = band / 2 + 1;
    const int rmid = low + midd - 1;
    int leftd = 1 - low, rightd = band;
    int CC[W2 + 2], DD[W2 + 2], CP[W2 + 2], DP[W2 + 2];
    int IP = 0;
    {
        const int fr = leftd - midd;
        const int t0 = (tb == 2) ? 0 : -g;
#pragma unroll
        for (int j = 0; j <= W2 + 1; j++) {
            int p;
            if (leftd < midd)      p = (j < midd) ? -1 : 0;
            else if (leftd > midd) p = (j <= midd) ? fr : -1;
            else                   p = 0;
            CP[j] = DP[j] = p;
            int cc = kNeg, dd = kNeg;
            if (j == leftd) { cc = 0; dd = (tb == 1) ? 0 : -g; }
            else if (j > leftd && j <= rightd) { cc = t0 - h * (j - leftd); dd = cc - g; }
            CC[j] = cc; DD[j] = dd;
        }
        if (leftd < midd || leftd == midd) x.rec[0][0] = x.rec[1][0] = x.rec[2][0] = dc_pack(-1, 0);
        else x.rec[0][fr] = x.rec[1][fr] = x.rec[2][fr] = dc_pack(-1, 0);
    }
    int c = 0, d = 0, e = 0;                                     // values of the last cell of the last row
    // the row's W2 bytes of B in a byte-shifted register window: byte q - 1 = B[q + low - 1 + i] (0 outside 1..N)
    constexpr int NW = (W2 + 3) / 4;
    uint32_t wr[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) wr[k] = 0;
#pragma unroll
    for (int q = 1; q <= W2; q++) {
        const int ib = q + low;
        if (ib >= 1 && ib <= N) wr[(q - 1) >> 2] |= (uint32_t)b1[ib] << (8 * ((q - 1) & 3));
    }
    uint32_t anext = (M >= 1) ? a1[1] : 0u;                      // fetched a row ahead
    for (int i = 1; i <= M; i++) {
        if (i > N - up) rightd--;
        if (leftd > 1) leftd--;
        const uint32_t ai = anext;
        if (i < M) anext = a1[i + 1];
        const int xb = W2 + low + i;                             // new top byte for row i + 1, fetched now, merged below
        const uint32_t wnew = (xb >= 1 && xb <= N) ? (uint32_t)b1[xb] : 0u;
        x.cells += rightd - leftd + 1;
        int cl = kNeg, el = kNeg;                                // horizontal inputs of the row's first cell
#pragma unroll
        for (int q = 1; q <= W2; q++) {
            if (q >= leftd && q <= rightd) {
                const bool first = q == leftd;
                const int ib = q + low - 1 + i;
                const uint32_t bq = (wr[(q - 1) >> 2] >> (8 * ((q - 1) & 3))) & 0xFFu;
                // horizontal
                int openh = cl - m, en = el - h;
                const bool eopen = openh > en;
                if (eopen) en = openh;
                const int ipn = eopen ? CP[q - 1] : IP;          // pointer of the best path ending in an insert here
                // vertical
                const int openv = CC[q + 1] - m;
                int dn = DD[q + 1] - h;
                const bool dopen = openv > dn;
                if (dopen) dn = openv;
                const int dpn = dopen ? CP[q + 1] : DP[q + 1];
                // diagonal (column 0 and below has none: CC[q] is -infinity there)
                int cn = CC[q] + ((ib > 0 && ai == bq) ? x.P->match : x.P->mismatch);
                if (ib <= 0) cn = kNeg;
                int cpn = CP[q], kind = 0;                       // 0 diagonal, 1 came from d, 2 came from e
                if (cn < dn || cn < en) {
                    if (en > dn) { cn = en; cpn = ipn; kind = 2; }
                    else         { cn = dn; cpn = dpn; kind = 1; }
                }
                if (!first && q == midd) {                       // crossing records (:189-232)
                    int mp1 = ipn, mt1 = 2, mp2 = dpn, mt2 = 1, mp0, mt0;
                    if (kind == 2)      { mp0 = mp1; mt0 = 2; }
                    else if (kind == 1) { mp0 = mp2; mt0 = 1; }
                    else                { mp0 = i - 1; mt0 = 0; }
                    if (cn - g > en) { mp1 = mp0; mt1 = mt0; }
                    if (cn - g > dn) { mp2 = mp0; mt2 = mt0; }
                    x.rec[0][i] = dc_pack(mp0, mt0); x.rec[1][i] = dc_pack(mp1, mt1); x.rec[2][i] = dc_pack(mp2, mt2);
                }
                IP = first ? cpn : ipn;                          // :164 IP = CP[leftd] ; else the carried insert pointer
                if (q == midd) { cpn = i; IP = i; DP[q] = i; } else DP[q] = dpn;
                CP[q] = cpn; CC[q] = cn; DD[q] = dn;
                cl = cn; el = first ? cn - g : en;               // :162 e = c - g after the first cell
                c = cn; d = dn; e = el;
            }
        }
#pragma unroll
        for (int k = 0; k < NW; k++) wr[k] = (wr[k] >> 8) | ((k + 1 < NW) ? (wr[k + 1] << 24) : 0u);
        wr[(W2 - 1) >> 2] |= wnew << (8 * ((W2 - 1) & 3));
    }
    int k, l;
    int dpr = DP[1], cpr = CP[1];
#pragma unroll
    for (int q = 2; q <= W2; q++) if (q == rightd) { dpr = DP[q]; cpr = CP[q]; }
    if (te == 1 && d + g > c)      { k = dpr; l = 2; }
    else if (te == 2 && e + g > c) { k = IP;  l = 1; }
    else                           { k = cpr; l = 0; }
    if (rmid > N - M) l = 2; else if (rmid < N - M) l = 1;
    int r = -1;
    while (k > -1) {
        x.fl[k] = dc_pack(r, l);
        r = k;
        const int nxt = x.rec[l][r];
        k = dc_ptr(nxt); l = dc_type(nxt);
    }
    f.rmid = rmid;
    f.k = r;
    if (r != -1) { const int v = x.fl[r]; f.l = dc_ptr(v); f.kt = dc_type(v); }
    f.t2 = up - rmid - 1; f.t3 = low - rmid + 1;
}

IG_HD inline void dc_push(DcFrame* st, int& sp, int a, int b, int M, int N,
                                        int low, int up, int tb, int te)
{
    DcFrame& f = st[sp++];
    f.a = a; f.b = b; f.M = M; f.N = N; f.low = low; f.up = up; f.tb = tb; f.te = te; f.stage = 0;
}

// align() of globalalign.c:66-307 with the recursion unrolled into frames.
template <int STRIDE>
IG_HD inline void dc_align(DcCtx<STRIDE>& x, DcFrame* st, int a0, int b0, int M0, int N0, int low0, int up0)
{
    int sp = 0;
    dc_push(st, sp, a0, b0, M0, N0, low0, up0, 0, 0);
    while (sp > 0) {
        DcFrame& f = st[sp - 1];
        switch (f.stage) {
        case 0: {
            if (f.N <= 0) { if (f.M > 0) put_del(x, f.M); sp--; break; }
            if (f.M <= 0) { put_ins(x, f.N); sp--; break; }
            if (f.up - f.low + 1 <= 1) { for (int i = 0; i < f.M; i++) put_rep(x); sp--; break; }
            {
                const int bw = f.up - f.low + 1;
                // one instantiation per width class: a narrower one skips fewer cells of its unrolled band
                // (measured: 2 classes 308, 5 classes 364, 7 classes 372 GCUPS at band 17)
                if (bw <= 3)       dc_sweep_reg<3>(x, f);
                else if (bw <= 5)  dc_sweep_reg<5>(x, f);
                else if (bw <= 8)  dc_sweep_reg<8>(x, f);
                else if (bw <= 10) dc_sweep_reg<10>(x, f);
                else if (bw <= 12) dc_sweep_reg<12>(x, f);
                else if (bw <= 16) dc_sweep_reg<16>(x, f);
                else if (bw <= 20) dc_sweep_reg<20>(x, f);
                else               dc_sweep(x, f);
            }
            const int r = f.k, rmid = f.rmid;
            if (r == -1) {                                   // :260-262
                f.stage = 6;
                if (rmid < 0) dc_push(st, sp, f.a, f.b, f.M, f.N, rmid + 1, f.up, f.tb, f.te);
                else          dc_push(st, sp, f.a, f.b, f.M, f.N, f.low, rmid - 1, f.tb, f.te);
                break;
            }
            if (rmid < 0)      { f.stage = 1; dc_push(st, sp, f.a, f.b, r - 1, r + rmid, rmid + 1, ig_min(f.up, r + rmid), f.tb, 1); }
            else if (rmid > 0) { f.stage = 2; dc_push(st, sp, f.a, f.b, r, r + rmid - 1, ig_max(-r, f.low), rmid - 1, f.tb, 2); }
            else f.stage = 3;
            break;
        }
        case 1: put_del(x, 1); f.stage = 3; break;           // :271
        case 2: put_ins(x, 1); f.stage = 3; break;           // :274
        case 3: {
            if (f.l > -1) {                                  // :280-293
                const int t1 = f.l - f.k - 1;
                if (f.kt == 0) { put_rep(x); f.k = f.l; { const int v = x.fl[f.k]; f.l = dc_ptr(v); f.kt = dc_type(v); } }
                else if (f.kt == 1) {
                    put_ins(x, 1); f.stage = 4;
                    dc_push(st, sp, f.a + f.k, f.b + f.k + f.rmid + 1, t1, t1, 0, ig_min(t1, f.t2), 2, 1);
                } else {
                    put_del(x, 1); f.stage = 5;
                    dc_push(st, sp, f.a + f.k + 1, f.b + f.k + f.rmid, t1, t1, ig_max(-t1, f.t3), 0, 1, 2);
                }
            } else {                                         // :296-304
                const int k = f.k, rmid = f.rmid;
                if (f.N - f.M > rmid) {
                    put_ins(x, 1); f.stage = 6;
                    const int t1 = k + rmid + 1;
                    dc_push(st, sp, f.a + k, f.b + t1, f.M - k, f.N - t1, 0, ig_min(f.N - t1, f.t2), 2, f.te);
                } else if (f.N - f.M < rmid) {
                    put_del(x, 1); f.stage = 6;
                    const int t1 = f.M - (k + 1);
                    dc_push(st, sp, f.a + k + 1, f.b + k + rmid, t1, f.N - (k + rmid), ig_max(-t1, f.t3), 0, 1, f.te);
                } else sp--;
            }
            break;
        }
        case 4: put_del(x, 1); f.k = f.l; { const int v = x.fl[f.k]; f.l = dc_ptr(v); f.kt = dc_type(v); } f.stage = 3; break;   // :286
        case 5: put_ins(x, 1); f.k = f.l; { const int v = x.fl[f.k]; f.l = dc_ptr(v); f.kt = dc_type(v); } f.stage = 3; break;   // :291
        default: sp--; break;
        }
    }
}

// Cells the reference's align() sweeps (globalalign.c:147-234, all recursion levels) when the optimal
// path of an M x M problem is the main diagonal and is unique: a level whose mid-diagonal is not
// diagonal 0 finds no crossing and recurses on the half band that holds the path (:260-262); the level
// whose mid-diagonal is diagonal 0 crosses on every row and emits its REPs without recursing.
IG_HD inline int dc_cells_of_diagonal_path(int M, int low, int up)
{
    low = ig_min(ig_max(-M, low), 0);                                  // ALIGN's clamps with N == M (:347-348)
    up  = ig_max(ig_min(M, up), 0);
    int cells = 0;
    while (up - low + 1 > 1) {
        const int band = up - low + 1;
        int leftd = 1 - low, rightd = band;
        for (int i = 1; i <= M; i++) {
            if (i > M - up) rightd--;
            if (leftd > 1) leftd--;
            cells += rightd - leftd + 1;
        }
        const int rmid = low + band / 2;                               // low + midd - 1, midd = band / 2 + 1
        if (rmid == 0) break;
        if (rmid < 0) low = rmid + 1; else up = rmid - 1;
    }
    return cells;
}

// ALIGN (globalalign.c:333-401): A, B 0-based first symbols.  Returns the number of script entries.
template <int STRIDE>
IG_HD inline int global_align_script(DcCtx<STRIDE>& x, DcFrame* st, int M, int N, int low, int up)
{
    x.ns = 0; x.last = 0;
    low = ig_min(ig_max(-M, low), ig_min(N - M, 0));                  // :347-348
    up  = ig_max(ig_min(N, up), ig_max(N - M, 0));
    if (N <= 0) { if (M > 0) put_del(x, M); }
    else if (M <= 0) put_ins(x, N);
    else if (up - low + 1 <= 1) { for (int i = 0; i < M; i++) put_rep(x); }
    else dc_align(x, st, 0, 0, M, N, low, up);           // views start one before the first symbol
    return x.ns;
}

// the script of an alignment without gaps: every entry is a replacement
struct ZeroScript { IG_HD int operator[](int) const { return 0; } };

// fetch_cigar (globalalign.c:507-604): A, B 0-based first ALIGNED symbols
template <class Script>
IG_HD inline int script_to_cigar(const uint8_t* A, const uint8_t* B, int M, int N, const Script& S,
                                 int AP, int readlength, uint32_t* cig, int ns = 0)
{
    int n = 0, i = 0, j = 0, k = 0;
    const int clip = AP - 1;
    if (clip > 0) cig[n++] = ((uint32_t)clip << 4) | OP_SOFT;
    int run_op = -1, run_len = 0, total = clip, pending = 0;
    auto emit = [&](int op) {
        if (run_op != -1 && run_op != op) { cig[n++] = ((uint32_t)run_len << 4) | (uint32_t)run_op; total += run_len; run_len = 0; }
        run_op = op; run_len++;
    };
    if (std::is_same<Script, ZeroScript>::value) {
        // gap-free script (every entry a replacement, so M == N): nothing to read but the two sequences;
        // four columns per step with the eight loads issued together, instead of a load-compare chain per base
        const int L = ig_min(M, N);
        for (; i + 4 <= L; i += 4) {
            const uint8_t a0 = A[i], a1 = A[i + 1], a2 = A[i + 2], a3 = A[i + 3];
            const uint8_t b0 = B[i], b1 = B[i + 1], b2 = B[i + 2], b3 = B[i + 3];
            emit(a0 == b0 ? OP_EQ : OP_X); emit(a1 == b1 ? OP_EQ : OP_X);
            emit(a2 == b2 ? OP_EQ : OP_X); emit(a3 == b3 ? OP_EQ : OP_X);
        }
        for (; i < L; i++) emit(A[i] == B[i] ? OP_EQ : OP_X);
        j = i;
    }
    while (i < M || j < N) {
        // four replacements at a time when the script says so (`ns` = its length, 0 = unknown): the script
        // lives in memory the D&C just wrote, and one round trip for twelve loads beats twelve round trips
        if (pending == 0 && k + 4 <= ns && i + 4 <= M && j + 4 <= N) {
            const int s0 = S[k], s1 = S[k + 1], s2 = S[k + 2], s3 = S[k + 3];
            const uint8_t a0 = A[i], a1 = A[i + 1], a2 = A[i + 2], a3 = A[i + 3];
            const uint8_t b0 = B[j], b1 = B[j + 1], b2 = B[j + 2], b3 = B[j + 3];
            if ((s0 | s1 | s2 | s3) == 0) {
                emit(a0 == b0 ? OP_EQ : OP_X); emit(a1 == b1 ? OP_EQ : OP_X);
                emit(a2 == b2 ? OP_EQ : OP_X); emit(a3 == b3 ? OP_EQ : OP_X);
                k += 4; i += 4; j += 4;
                continue;
            }
        }
        int op;
        if (pending == 0 && S[k] == 0) { k++; op = (A[i] == B[j]) ? OP_EQ : OP_X; i++; j++; }
        else {
            if (pending == 0) pending = S[k++];
            if (pending > 0) { pending--; j++; op = OP_DEL; }
            else             { pending++; i++; op = OP_INS; }
        }
        emit(op);
    }
    if (run_op != -1 && run_len > 0) { cig[n++] = ((uint32_t)run_len << 4) | (uint32_t)run_op; total += run_len; }
    if (total < readlength) cig[n++] = ((uint32_t)(readlength - total) << 4) | OP_SOFT;
    return n;
}

// ---------------------------------------------------------------------------------------
// local_align's two sweeps (localalign.c:82-176) for bands of at most W diagonals with ALL DP state in
// registers: one (H, D) pair per diagonal, updated in place (ascending t forward, descending t in
// reverse, so every cell still sees the previous row's neighbours), the band's W reference bytes in a
// byte-shifted register window, every loop over the band fully unrolled.  Cells outside the band or
// the window keep -infinity, which reproduces the reference's boundary cases (t == tlo, j == 0,
// the column right of the end cell) without branches: with E, D or the diagonal at -infinity each
// special case of the serial code is the general recurrence.
// Same results and cell counts as the memory-resident sweeps below (tests/test_host_harness.py).
// ---------------------------------------------------------------------------------------
// WMIN = the narrowest band routed to this instantiation: diagonals t < WMIN lie inside every band it sees.
// Rows whose band is not cut by the window (the great majority) take a branch-free body: no per-cell validity
// selects, and the position of the row's first maximum rides in the low bits of one max() per cell
// (key = score * 64 + 63 - t: the largest key is the largest score at the smallest t) instead of a compare and
// three selects per cell; rows at the window's edges keep the general body.
template <int W, int WMIN>
IG_HD inline void local_sweeps_reg(const DevParams& P, const uint8_t* read, int M, const uint8_t* win, int N,
                                   int low, int up, int& best_o, int& endi_o, int& endj_o,
                                   int& starti_o, int& startj_o, bool& found_o, int& cf_o, int& cr_o)
{
    constexpr int NW = (W + 3) / 4;
    const int G = P.G, Hh = P.H, m = G + Hh;
    const int band = up - low + 1;
    const int si = ig_max(0, -up), ei = ig_min(M, N - low);
    int Hr[W], Dr[W];
    uint32_t wr[NW];
#pragma unroll
    for (int t = 0; t < W; t++) {
        const int j = si + low + t;
        const bool v = t < band && j >= 0 && j <= N;
        Hr[t] = v ? 0 : kNeg; Dr[t] = v ? -G : kNeg;
    }
#pragma unroll
    for (int k = 0; k < NW; k++) wr[k] = 0;
#pragma unroll
    for (int t = 0; t < W; t++) {                                  // byte t = win[i + low + t - 1] for row i = si + 1
        const int x = si + low + t;
        if (x >= 0 && x < N) wr[t >> 2] |= (uint32_t)win[x] << (8 * (t & 3));
    }
    int best = 0, endi = si, endt = 0, cf = 0;
    uint32_t anext = (si + 1 <= ei) ? read[si] : 0u;              // the next row's read base is fetched a row ahead
    for (int i = si + 1; i <= ei; i++) {
        const int tlo = ig_max(0, -i - low), thi = ig_min(band - 1, N - i - low);
        const uint32_t ai = anext;
        if (i < ei) anext = read[i];
        const int xn = i + low + W - 1;                             // new top byte for row i + 1, fetched now, merged below
        const uint32_t wnew = (xn >= 0 && xn < N) ? (uint32_t)win[xn] : 0u;
        int e = kNeg, left = kNeg;
        if (tlo == 0 && thi == band - 1) {
            // the band lies inside the window on this row
            int key = -1;
#pragma unroll
            for (int t = 0; t < W; t++) {
                const int hup = (t + 1 < W) ? Hr[t + 1] : kNeg, dup = (t + 1 < W) ? Dr[t + 1] : kNeg;
                const int d = ig_max(hup - m, dup - Hh);
                const bool eq = ((wr[t >> 2] >> (8 * (t & 3))) & 0xFFu) == ai;
                e = ig_max(left - m, e - Hh);
                const int c = ig_max(ig_max(Hr[t] + (eq ? P.match : P.mismatch), e), ig_max(d, 0));
                if (t < WMIN) {
                    Hr[t] = c; Dr[t] = d; left = c;
                    key = ig_max(key, c * 64 + (63 - t));
                } else if (t < band) {
                    Hr[t] = c; Dr[t] = d; left = c;
                    key = ig_max(key, c * 64 + (63 - t));
                } else { left = kNeg; e = kNeg; }
            }
            if ((key >> 6) > best) { best = key >> 6; endi = i; endt = 63 - (key & 63); }
        } else {
#pragma unroll
            for (int t = 0; t < W; t++) {
                const int hup = (t + 1 < W) ? Hr[t + 1] : kNeg, dup = (t + 1 < W) ? Dr[t + 1] : kNeg;
                const int d = ig_max(hup - m, dup - Hh);
                const bool eq = ((wr[t >> 2] >> (8 * (t & 3))) & 0xFFu) == ai;
                e = ig_max(left - m, e - Hh);
                int c = ig_max(ig_max(Hr[t] + (eq ? P.match : P.mismatch), e), ig_max(d, 0));
                const bool v = t >= tlo && t <= thi;
                if (v) {
                    Hr[t] = c; Dr[t] = d; left = c;
                    if (c > best) { best = c; endi = i; endt = t; }
                } else { left = kNeg; e = kNeg; }
            }
        }
        cf += thi - tlo + 1;
#pragma unroll
        for (int k = 0; k < NW; k++) wr[k] = (wr[k] >> 8) | ((k + 1 < NW) ? (wr[k + 1] << 24) : 0u);
        wr[(W - 1) >> 2] |= wnew << (8 * ((W - 1) & 3));
    }
    const int endj = endi + low + endt;
    best_o = best; endi_o = endi; endj_o = (best > 0) ? endj : si + low; cf_o = cf;
    starti_o = 0; startj_o = 0; found_o = false; cr_o = 0;
    if (best <= 0) return;

    // reverse (localalign.c:132-176)
    const int tend = (endj - endi) - low;
    {
        const int tl = ig_max(0, -endi - low);
#pragma unroll
        for (int t = 0; t < W; t++) {
            const bool v = t <= tend && t >= tl;
            const int acc = -(G + Hh * (tend - t));
            Hr[t] = (t == tend) ? 0 : (v ? acc : kNeg);
            Dr[t] = (t == tend) ? -G : (v ? acc - G : kNeg);
        }
    }
#pragma unroll
    for (int k = 0; k < NW; k++) wr[k] = 0;
#pragma unroll
    for (int t = 0; t < W; t++) {                                  // byte t = win[i + low + t - 1] for row i = endi
        const int x = endi + low + t - 1;
        if (x >= 0 && x < N) wr[t >> 2] |= (uint32_t)win[x] << (8 * (t & 3));
    }
    int starti = 0, startt = 0, cr = 0; bool found = false;
    anext = (endi >= 1) ? read[endi - 1] : 0u;
    for (int i = endi; i >= 1 && !found; i--) {
        const int thi = ig_min(band - 1, tend + (endi - i) + 1);
        const int tlo = ig_max(0, 1 - i - low);
        const uint32_t ai = anext;
        if (i > 1) anext = read[i - 2];
        const int xn = i + low - 2;                                 // new bottom byte for row i - 1, fetched now, merged below
        const uint32_t wnew = (xn >= 0 && xn < N) ? (uint32_t)win[xn] : 0u;
        int e = kNeg, right = kNeg;
        if (tlo == 0 && thi == band - 1) {
            // the band lies inside the window on this row: every cell is computed, and the first cell in sweep order
            // (largest t) that reaches the optimum is picked afterwards -- what the sweep writes behind it is never read
            int tfound = -1;
#pragma unroll
            for (int t = W - 1; t >= 0; t--) {
                const int hdn = (t > 0) ? Hr[t - 1] : kNeg, ddn = (t > 0) ? Dr[t - 1] : kNeg;
                const int d = ig_max(hdn - m, ddn - Hh);
                const bool eq = ((wr[t >> 2] >> (8 * (t & 3))) & 0xFFu) == ai;
                e = ig_max(right - m, e - Hh);
                const int c = ig_max(ig_max(Hr[t] + (eq ? P.match : P.mismatch), e), d);
                if (t < WMIN || t < band) {
                    Hr[t] = c; Dr[t] = d; right = c;
                    tfound = ig_max(tfound, c == best ? t : -1);
                } else { right = kNeg; e = kNeg; }
            }
            if (tfound >= 0) { found = true; starti = i; startt = tfound; cr += thi - tfound + 1; }
        } else {
#pragma unroll
            for (int t = W - 1; t >= 0; t--) {
                const int hdn = (t > 0) ? Hr[t - 1] : kNeg, ddn = (t > 0) ? Dr[t - 1] : kNeg;
                const int d = ig_max(hdn - m, ddn - Hh);
                const bool eq = ((wr[t >> 2] >> (8 * (t & 3))) & 0xFFu) == ai;
                e = ig_max(right - m, e - Hh);
                const int c = ig_max(ig_max(Hr[t] + (eq ? P.match : P.mismatch), e), d);
                const bool v = t >= tlo && t <= thi && !found;
                if (v) {
                    Hr[t] = c; Dr[t] = d; right = c;
                    if (c == best) { found = true; starti = i; startt = t; cr += thi - t + 1; }
                } else {
                    right = kNeg; e = kNeg;
                    if (t < tlo) { Hr[t] = kNeg; Dr[t] = kNeg; }          // left the window: stale values must not be read
                }
            }
        }
        if (!found) cr += ig_max(0, thi - tlo + 1);
#pragma unroll
        for (int k = NW - 1; k >= 0; k--) wr[k] = (wr[k] << 8) | ((k > 0) ? (wr[k - 1] >> 24) : 0u);
        wr[0] |= wnew;
    }
    starti_o = starti; startj_o = starti + low + startt; found_o = found; cr_o = cr;
}

// ---------------------------------------------------------------------------------------
// One banded alignment = local_align (two sweeps) + ALIGN (divide and conquer) + fetch_cigar, executed
// by ONE thread.  It is split in two phases so that kernels can run the second phase only for the
// alignments that need it, 32 at a time per warp (banded_two_phase_loop): inside a warp the
// few gapped alignments would otherwise serialise against the many that finish after phase 1.
// `low`/`up` are already clamped (localalign.c:70-71).
//   bands : 4 * (max_band + 4) ints (shared memory when it fits, else the head of the global slice)
//   rowsb : 4 * (max_rows + 2) ints + the script
// ---------------------------------------------------------------------------------------
// an alignment waiting for its ALIGN phase (banded_two_phase_loop)
struct DcTask { int idx, best, endi, endj, starti, startj; };

struct BandLocal {
    int best, endi, endj, starti, startj, cf, cr;
    bool none;                 // no alignment (localalign.c:180-193)
};

// a script of zeros that needs no memory (all-REP alignments)

// phase 1a: the two sweeps of local_align
template <int STRIDE>
IG_HD inline BandLocal band_local(const DevParams& P, IArr<STRIDE> bands, int max_band,
                                  const uint8_t* read, int M, const uint8_t* win, int N, int low, int up)
{
    const int G = P.G, H = P.H, m = G + H;
    const int band = up - low + 1;
    const int wb = max_band + 4;
    IArr<STRIDE> Hp = bands, Dp = bands + wb, Hn = bands + 2 * wb, Dn = bands + 3 * wb;
    int best = 0, endi = 0, endj = 0, cf = 0, cr = 0, starti = 0, startj = 0; bool found = false;
    // the unrolled sweeps compute W diagonals whatever the band: pick the tightest instantiation
#define IG_SWEEP(W, WMIN) local_sweeps_reg<W, WMIN>(P, read, M, win, N, low, up, best, endi, endj, starti, startj, found, cf, cr)
    if (band <= 6)       IG_SWEEP(6, 1);
    else if (band <= 8)  IG_SWEEP(8, 7);
    else if (band <= 12) IG_SWEEP(12, 9);
    else if (band <= 16) IG_SWEEP(16, 13);
    else if (band <= 20) IG_SWEEP(20, 17);
    else if (band <= 24) IG_SWEEP(24, 21);
    else if (band <= 32) IG_SWEEP(32, 25);
    else if (band <= 40) IG_SWEEP(40, 33);
#undef IG_SWEEP
    else {
#define AT(arr, t) ((arr)[(t) + 1])
    // forward (localalign.c:82-131)
    const int si = ig_max(0, -up), ei = ig_min(M, N - low);
    // Cells outside the band read as -infinity.  Row i+1 only reads row i inside [tlo(i) - 1, thi(i) + 1]
    // (both limits move by at most one per row), so after one full fill it is enough to refresh the two
    // entries just outside each row's range.
    for (int t = -1; t <= band; t++) { AT(Hp, t) = kNeg; AT(Dp, t) = kNeg; AT(Hn, t) = kNeg; AT(Dn, t) = kNeg; }
    for (int t = 0; t < band; t++) {
        const int j = si + low + t;
        if (j >= 0 && j <= N) { AT(Hp, t) = 0; AT(Dp, t) = -G; }
    }
    best = 0; endi = si; endj = si + low; cf = 0; cr = 0;
    for (int i = si + 1; i <= ei; i++) {
        const int tlo = ig_max(0, -i - low), thi = ig_min(band - 1, N - i - low);
        AT(Hn, tlo - 1) = kNeg; AT(Dn, tlo - 1) = kNeg; AT(Hn, thi + 1) = kNeg; AT(Dn, thi + 1) = kNeg;
        int e = kNeg, left = kNeg;
        const uint8_t ai = read[i - 1];
        for (int t = tlo; t <= thi; t++) {
            const int j = i + low + t;
            const int d = ig_max(AT(Hp, t + 1) - m, AT(Dp, t + 1) - H);
            int c;
            if (j == 0) c = d;
            else {
                c = AT(Hp, t) + (ai == win[j - 1] ? P.match : P.mismatch);
                if (t > tlo) { e = ig_max(left - m, e - H); if (e > c) c = e; }
                if (d > c) c = d;
            }
            if (c < 0) c = 0;
            if (t == tlo) e = c - G;
            left = c;
            AT(Hn, t) = c; AT(Dn, t) = d;
            if (c > best) { best = c; endi = i; endj = j; }
        }
        cf += thi - tlo + 1;
        IArr<STRIDE> tmp = Hp; Hp = Hn; Hn = tmp; tmp = Dp; Dp = Dn; Dn = tmp;
    }
    // reverse (localalign.c:132-176)
    starti = 0; startj = 0; found = false;
    if (best > 0) {
        const int tend = (endj - endi) - low;
        for (int t = -1; t <= band; t++) { AT(Hp, t) = kNeg; AT(Dp, t) = kNeg; AT(Hn, t) = kNeg; AT(Dn, t) = kNeg; }
        {
            const int tl = ig_max(0, -endi - low);
            AT(Hp, tend) = 0; AT(Dp, tend) = -G;
            int acc = -G;
            for (int t = tend - 1; t >= tl; t--) { acc -= H; AT(Hp, t) = acc; AT(Dp, t) = acc - G; }
        }
        for (int i = endi; i >= 1 && !found; i--) {
            const int thi = ig_min(band - 1, tend + (endi - i) + 1);
            const int tlo = ig_max(0, 1 - i - low);
            if (tlo - 1 <= band) { AT(Hn, tlo - 1) = kNeg; AT(Dn, tlo - 1) = kNeg; }
            if (thi + 1 <= band) { AT(Hn, thi + 1) = kNeg; AT(Dn, thi + 1) = kNeg; }
            int e = kNeg, right = kNeg;
            const uint8_t ai = read[i - 1];
            for (int t = thi; t >= tlo; t--) {
                const int j = i + low + t;
                const int d = ig_max(AT(Hp, t - 1) - m, AT(Dp, t - 1) - H);
                int c;
                if (t == thi) {
                    c = (j <= N) ? AT(Hp, t) + (ai == win[j - 1] ? P.match : P.mismatch) : AT(Hp, t - 1) - m;
                    if (d > c) c = d;
                    e = c - G;
                } else {
                    e = ig_max(right - m, e - H);
                    c = AT(Hp, t) + (ai == win[j - 1] ? P.match : P.mismatch);
                    if (e > c) c = e;
                    if (d > c) c = d;
                }
                right = c;
                AT(Hn, t) = c; AT(Dn, t) = d;
                cr++;
                if (c == best) { starti = i; startj = j; found = true; break; }
            }
            IArr<STRIDE> tmp = Hp; Hp = Hn; Hn = tmp; tmp = Dp; Dp = Dn; Dn = tmp;
        }
    }
    }
#undef AT
    BandLocal L;
    L.best = best; L.endi = endi; L.endj = endj; L.starti = starti; L.startj = startj; L.cf = cf; L.cr = cr;
    L.none = best <= 0 || !found || starti > M || startj > N ||
             endi - starti == 0 || endj - startj == 0;            // localalign.c:180-193
    return L;
}

// phase 1b, exact shortcut: when the end points lie on one diagonal, the ungapped path reaches the local
// optimum and it has so few mismatches that every gapped path (which pays at least two gap opens and two
// extensions and aligns at least one pair fewer) scores strictly less, the all-REP script is the UNIQUE
// optimum, so ALIGN's divide and conquer (globalalign.c:66-307) returns it whatever its tie rules.
// Writes the CIGAR and returns true; the cell count ALIGN would have swept depends on the geometry only.
IG_HD inline bool band_unique_diagonal(const DevParams& P, const uint8_t* read, int M, const uint8_t* win,
                                       int low, int up, const BandLocal& L, uint32_t* cig, int* ncig, int* cells_glob)
{
    const int M2 = L.endi - L.starti + 1, N2 = L.endj - L.startj + 1;
    if (M2 != N2 || !(P.match > 0 && P.mismatch < P.match && P.G >= 0 && P.H >= 0)) return false;
    const uint8_t* A = read + L.starti - 1;
    const uint8_t* B = win + L.startj - 1;
    int mm = 0;
    for (int i = 0; i < M2; i++) mm += (A[i] != B[i]) ? 1 : 0;
    const long long sd = (long long)(M2 - mm) * P.match + (long long)mm * P.mismatch;
    if (sd != (long long)L.best || (long long)mm * (P.match - P.mismatch) >= (long long)P.match + 2LL * (P.G + P.H)) return false;
    *ncig = script_to_cigar(A, B, M2, N2, ZeroScript(), L.starti, M, cig);
    *cells_glob = dc_cells_of_diagonal_path(M2, low - (L.startj - L.starti), up - (L.startj - L.starti));
    return true;
}

// phase 2: ALIGN (divide and conquer) on the sub-rectangle + fetch_cigar.  Leaves the script in rowsb.
template <int STRIDE>
IG_HD inline void band_global(const DevParams& P, IArr<STRIDE> bands, IArr<STRIDE> rowsb, int max_band, int max_rows, DcFrame* st,
                              const uint8_t* read, int M, const uint8_t* win, int low, int up, const BandLocal& L,
                              uint32_t* cig, int* ncig, int* cells_glob, int* nscript)
{
    const int wb = max_band + 4, wr = max_rows + 2;
    DcCtx<STRIDE> x;
    x.P = &P; x.cells = 0; x.ns = 0; x.last = 0;
    x.cc = bands; x.dd = bands + wb; x.cp = bands + 2 * wb; x.dp = bands + 3 * wb;
    x.rec[0] = rowsb; x.rec[1] = rowsb + wr; x.rec[2] = rowsb + 2 * wr; x.fl = rowsb + 3 * wr;
    x.S = rowsb + 4 * wr;
    const int M2 = L.endi - L.starti + 1, N2 = L.endj - L.startj + 1;
    x.A = read + L.starti - 1; x.B = win + L.startj - 1;
    global_align_script(x, st, M2, N2, low - (L.startj - L.starti), up - (L.startj - L.starti));
    *ncig = script_to_cigar(x.A, x.B, M2, N2, x.S, L.starti, M, cig, x.ns);
    *cells_glob = x.cells; *nscript = x.ns;
}

// both phases back to back (host harness, single-task kernels)
// out: score, q1, r1, q2, r2 (1-based inclusive, slice/window relative), ncigar, cells fwd, rev, glob, nscript
template <int STRIDE>
IG_HD inline void align_banded_serial(const DevParams& P, IArr<STRIDE> bands, IArr<STRIDE> rowsb, int max_band, int max_rows, DcFrame* st,
                                      const uint8_t* read, int M, const uint8_t* win, int N,
                                      int low, int up, uint32_t* cig, int* out)
{
    const BandLocal L = band_local<STRIDE>(P, bands, max_band, read, M, win, N, low, up);
    int n = 0, cg = 0, ns = 0;
    if (!L.none) {
        if (band_unique_diagonal(P, read, M, win, low, up, L, cig, &n, &cg)) {
            ns = L.endi - L.starti + 1;
            const IArr<STRIDE> S = rowsb + 4 * (max_rows + 2);
            for (int i = 0; i < ns; i++) S[i] = 0;
        } else {
            band_global<STRIDE>(P, bands, rowsb, max_band, max_rows, st, read, M, win, low, up, L, cig, &n, &cg, &ns);
        }
    }
    out[0] = L.none ? 0 : L.best;     // ALIGN's score equals the local optimum
    out[1] = L.starti; out[2] = L.startj; out[3] = L.endi; out[4] = L.endj;
    out[5] = n; out[6] = L.cf; out[7] = L.cr; out[8] = L.none ? 0 : cg;
    out[9] = L.none ? 0 : ns;         // script entries, left in rowsb
}

// ---------------------------------------------------------------------------------------
// local_align's two sweeps for WIDE bands (41..160 diagonals), one alignment per WARP: lane l owns the C
// consecutive diagonals t = l*C .. l*C+C-1 and the warp does one row per step.  The row's horizontal
// recurrence  e_t = max(c_{t-1} - (G+H), e_{t-1} - H)  unrolls to a prefix maximum,
//     e_t = max_{t' < t} (h_{t'} - (G+H) + (t'+1) H) - t H ,   h = the cell value without its e term,
// (a gap opened from a cell that was itself reached by e is never better than extending that gap, G >= 0),
// so a row is: the vertical and diagonal terms in registers (the neighbour lane's edge by shuffle), one
// warp max-scan, the combine.  Same values, end points and cell counts as local_sweeps_reg (the serial
// statement of the same recurrence); 32 alignments per warp in rolling memory rows were 2.5x slower per cell.
// All arguments are warp-uniform; every lane returns the same result.
// ---------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ BandLocal local_sweeps_warp(const DevParams& P, const uint8_t* read, int M, const uint8_t* win,
                                                       int N, int low, int up)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int G = P.G, Hh = P.H, m = G + Hh;
    const int band = up - low + 1;
    const int si = ig_max(0, -up), ei = ig_min(M, N - low);
    const int t0 = lane * C;
    const int kLow = 2 * kNeg;                                    // "no candidate" of the scans
    int Hr[C], Dr[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int t = t0 + c, j = si + low + t;
        const bool v = t < band && j >= 0 && j <= N;
        Hr[c] = v ? 0 : kNeg; Dr[c] = v ? -G : kNeg;
    }
    int best = 0, endi = si, endt = 0, cf = 0;
    // Rows whose band lies inside the window (nearly all of them) take a body without per-cell bounds tests: which of a
    // lane's slots exist (t < band) does not change from row to row, the window byte of slot c is wl[c][i], the scan's
    // position terms are per-lane constants, and the position of the lane's best cell of the row rides in the low bits
    // of one maximum (key = score * 256 + 255 - t).
    bool vs[C]; const uint8_t* wl[C]; int kA[C], tH[C];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int t = t0 + c;
        vs[c] = t < band;
        wl[c] = win + (low - 1) + (vs[c] ? t : 0);                // a slot outside the band reads slot 0's byte and ignores it
        kA[c] = (t + 1) * Hh - m;
        tH[c] = t * Hh;
    }
#pragma unroll 1
    for (int i = si + 1; i <= ei; i++) {
        const int tlo = ig_max(0, -i - low), thi = ig_min(band - 1, N - i - low);
        const uint32_t ai = read[i - 1];
        int hnext = __shfl_down_sync(FULL, Hr[0], 1), dnext = __shfl_down_sync(FULL, Dr[0], 1);
        if (lane == 31) { hnext = kNeg; dnext = kNeg; }
        int hq[C], dq[C];
        if (i + low >= 1 && i + low + band - 1 <= N) {            // tlo == 0, thi == band - 1, every byte inside the window
            int laneA = kLow;
#pragma unroll
            for (int c = 0; c < C; c++) {
                const int hup = (c + 1 < C) ? Hr[c + 1] : hnext, dup = (c + 1 < C) ? Dr[c + 1] : dnext;
                const int d = ig_max(hup - m, dup - Hh);
                const uint32_t b = wl[c][i];
                hq[c] = ig_max(Hr[c] + (b == ai ? P.match : P.mismatch), ig_max(d, 0));
                dq[c] = d;
                if (vs[c]) laneA = ig_max(laneA, hq[c] + kA[c]);
            }
            int incl = laneA;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl = ig_max(incl, v); }
            int run = __shfl_up_sync(FULL, incl, 1);
            if (lane == 0) run = kLow;
            int key = -1;
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (vs[c]) {
                    const int cv = ig_max(hq[c], run - tH[c]);
                    run = ig_max(run, hq[c] + kA[c]);
                    Hr[c] = cv; Dr[c] = dq[c];
                    key = ig_max(key, cv * 256 + (255 - (t0 + c)));
                }
            }
            if ((key >> 8) > best) { best = key >> 8; endi = i; endt = 255 - (key & 255); }
            cf += band;
            continue;
        }
        int laneA = kLow;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int t = t0 + c;
            const int hup = (c + 1 < C) ? Hr[c + 1] : hnext, dup = (c + 1 < C) ? Dr[c + 1] : dnext;
            const int d = ig_max(hup - m, dup - Hh);
            const int x = i + low + t - 1;
            const uint32_t b = (x >= 0 && x < N) ? (uint32_t)win[x] : 0u;
            hq[c] = ig_max(Hr[c] + (b == ai ? P.match : P.mismatch), ig_max(d, 0));
            dq[c] = d;
            if (t >= tlo && t <= thi) laneA = ig_max(laneA, hq[c] - m + (t + 1) * Hh);
        }
        int incl = laneA;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl = ig_max(incl, v); }
        int run = __shfl_up_sync(FULL, incl, 1);
        if (lane == 0) run = kLow;
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int t = t0 + c;
            if (t >= tlo && t <= thi) {
                const int e = run - t * Hh;
                const int cv = ig_max(hq[c], e);
                run = ig_max(run, hq[c] - m + (t + 1) * Hh);
                Hr[c] = cv; Dr[c] = dq[c];
                if (cv > best) { best = cv; endi = i; endt = t; }
            }
        }
        cf += thi - tlo + 1;
    }
    {   // first maximum in row-major order: largest score, then smallest row, then smallest diagonal
        const int top = __reduce_max_sync(FULL, best);
        const unsigned pos = best == top ? ((unsigned)endi << 12) | (unsigned)endt : 0xFFFFFFFFu;
        const unsigned first = __reduce_min_sync(FULL, pos);
        best = top; endi = (int)(first >> 12); endt = (int)(first & 0xFFFu);
        if (top <= 0) { endi = si; endt = 0; }
    }
    BandLocal L;
    const int endj = endi + low + endt;
    L.best = best; L.endi = endi; L.endj = (best > 0) ? endj : si + low; L.cf = cf; L.cr = 0;
    L.starti = 0; L.startj = 0; L.none = true;
    if (best <= 0) return L;

    // reverse (localalign.c:132-176)
    const int tend = (endj - endi) - low;
    {
        const int tl = ig_max(0, -endi - low);
#pragma unroll
        for (int c = 0; c < C; c++) {
            const int t = t0 + c;
            const bool v = t <= tend && t >= tl;
            const int acc = -(G + Hh * (tend - t));
            Hr[c] = (t == tend) ? 0 : (v ? acc : kNeg);
            Dr[c] = (t == tend) ? -G : (v ? acc - G : kNeg);
        }
    }
    int starti = 0, startt = 0, cr = 0; bool found = false;
    int kB[C];
#pragma unroll
    for (int c = 0; c < C; c++) kB[c] = -m - (t0 + c - 1) * Hh;
#pragma unroll 1
    for (int i = endi; i >= 1 && !found; i--) {
        const int thi = ig_min(band - 1, tend + (endi - i) + 1);
        const int tlo = ig_max(0, 1 - i - low);
        const uint32_t ai = read[i - 1];
        int hprev = __shfl_up_sync(FULL, Hr[C - 1], 1), dprev = __shfl_up_sync(FULL, Dr[C - 1], 1);
        if (lane == 0) { hprev = kNeg; dprev = kNeg; }
        int hq[C], dq[C];
        if (thi == band - 1 && i + low >= 1 && i + low + band - 1 <= N) {   // the wedge has reached the band's width; all bytes inside
            int laneB = kLow;
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                const int hdn = (c > 0) ? Hr[c - 1] : hprev, ddn = (c > 0) ? Dr[c - 1] : dprev;
                const int d = ig_max(hdn - m, ddn - Hh);
                const uint32_t b = wl[c][i];
                hq[c] = ig_max(Hr[c] + (b == ai ? P.match : P.mismatch), d);
                dq[c] = d;
                if (vs[c]) laneB = ig_max(laneB, hq[c] + kB[c]);
            }
            int incl = laneB;                                     // suffix maximum over the lanes above
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_down_sync(FULL, incl, o); if (lane + o < 32) incl = ig_max(incl, v); }
            int run = __shfl_down_sync(FULL, incl, 1);
            if (lane == 31) run = kLow;
            int hit = -1;                                         // largest own diagonal whose value equals the optimum
#pragma unroll
            for (int c = C - 1; c >= 0; c--) {
                if (vs[c]) {
                    const int cv = ig_max(hq[c], run + tH[c]);
                    run = ig_max(run, hq[c] + kB[c]);
                    Hr[c] = cv; Dr[c] = dq[c];
                    if (cv == best && hit < 0) hit = t0 + c;
                }
            }
            const int tophit = __reduce_max_sync(FULL, hit);
            if (tophit >= 0) { found = true; starti = i; startt = tophit; cr += thi - tophit + 1; }
            else cr += band;
            continue;
        }
        int laneB = kLow;
#pragma unroll
        for (int c = C - 1; c >= 0; c--) {
            const int t = t0 + c;
            const int hdn = (c > 0) ? Hr[c - 1] : hprev, ddn = (c > 0) ? Dr[c - 1] : dprev;
            const int d = ig_max(hdn - m, ddn - Hh);
            const int x = i + low + t - 1;
            const uint32_t b = (x >= 0 && x < N) ? (uint32_t)win[x] : 0u;
            hq[c] = ig_max(Hr[c] + (b == ai ? P.match : P.mismatch), d);
            dq[c] = d;
            if (t >= tlo && t <= thi) laneB = ig_max(laneB, hq[c] - m - (t - 1) * Hh);
        }
        int incl = laneB;                                         // suffix maximum over the lanes above
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_down_sync(FULL, incl, o); if (lane + o < 32) incl = ig_max(incl, v); }
        int run = __shfl_down_sync(FULL, incl, 1);
        if (lane == 31) run = kLow;
        int hit = -1;                                             // largest own diagonal whose value equals the optimum
#pragma unroll
        for (int c = C - 1; c >= 0; c--) {
            const int t = t0 + c;
            if (t >= tlo && t <= thi) {
                const int e = run + t * Hh;
                const int cv = ig_max(hq[c], e);
                run = ig_max(run, hq[c] - m - (t - 1) * Hh);
                Hr[c] = cv; Dr[c] = dq[c];
                if (cv == best && hit < 0) hit = t;
            } else if (t < tlo) { Hr[c] = kNeg; Dr[c] = kNeg; }   // left the window: stale values must not be read
        }
        const int tophit = __reduce_max_sync(FULL, hit);
        if (tophit >= 0) { found = true; starti = i; startt = tophit; cr += thi - tophit + 1; }
        else cr += ig_max(0, thi - tlo + 1);
    }
    L.cr = cr; L.starti = starti; L.startj = starti + low + startt;
    L.none = !found || starti > M || L.startj > N || L.endi - starti == 0 || L.endj - L.startj == 0;   // localalign.c:180-193
    return L;
}

constexpr int kWarpBandMin = 41, kWarpBandMax = 160;

__device__ __forceinline__ BandLocal local_sweeps_warp_any(const DevParams& P, const uint8_t* read, int M, const uint8_t* win,
                                                           int N, int low, int up)
{
    const int band = up - low + 1;
    if (band <= 64) return local_sweeps_warp<2>(P, read, M, win, N, low, up);
    if (band <= 96) return local_sweeps_warp<3>(P, read, M, win, N, low, up);
    if (band <= 128) return local_sweeps_warp<4>(P, read, M, win, N, low, up);
    return local_sweeps_warp<5>(P, read, M, win, N, low, up);
}

// The warp-per-alignment sweeps for the tasks of one pass of banded_two_phase_loop: every lane says whether
// its task wants them (`mine`: valid task, band in [kWarpBandMin, kWarpBandMax], G >= 0); the warp then
// serves the lanes one after the other.  fetch(idx, &read, &M, &win, &N, &lo, &hi) reloads a task's
// geometry from its index (warp-uniform).
template <class Fetch>
__device__ __forceinline__ void warp_serve_wide_bands(const DevParams& P, int idx, bool mine, Fetch fetch, BandLocal& L)
{
    const int lane = threadIdx.x & 31;
    uint32_t todo = __ballot_sync(0xFFFFFFFFu, mine);
#pragma unroll 1
    while (todo) {
        const int q = __ffs(todo) - 1;
        todo &= todo - 1;
        const int idxq = __shfl_sync(0xFFFFFFFFu, idx, q);
        const uint8_t* read; const uint8_t* win; int M, N, lo, hi;
        fetch(idxq, &read, &M, &win, &N, &lo, &hi);
        const BandLocal R = local_sweeps_warp_any(P, read, M, win, N, lo, hi);
        if (lane == q) L = R;
    }
}

#ifdef __CUDACC__
// Warp-level scheduling of the two phases (used by band_tasks_kernel and pipe_dp_kernel).  Every lane
// runs phase 1 (sweeps + shortcut) on its own task; tasks that need ALIGN are collected in a per-warp
// buffer and phase 2 runs whenever 32 are waiting (and once more for the tail), so that the divide and
// conquer always executes with a full warp while other warps of the SM hide its latency.
// phase0(idx, valid, BandLocal&) -> bool "sweeps done by the warp" (called by all 32 lanes together);
// phase1(idx, DcTask&, const BandLocal* pre) -> bool "needs phase 2"; phase2(const DcTask&).
// `pend` holds 64 DcTask per warp.
template <class P0, class P1, class P2>
__device__ __forceinline__ void banded_two_phase_loop(int n, DcTask* pend, P0 phase0, P1 phase1, P2 phase2)
{
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int waiting = 0;                                             // warp-uniform
#pragma unroll 1
    for (int idx0 = gwarp * 32; idx0 < n || waiting > 0; idx0 += nwarps * 32) {
        if (idx0 < n) {
            const int idx = idx0 + lane;
            DcTask t;
            t.idx = idx;
            BandLocal pre;
            const bool have = phase0(idx, idx < n, pre);         // warp-cooperative part (wide bands); all lanes call it
            const bool q = (idx < n) ? phase1(idx, t, have ? &pre : nullptr) : false;
            const uint32_t qm = __ballot_sync(0xFFFFFFFFu, q);
            if (q) pend[waiting + __popc(qm & ((1u << lane) - 1u))] = t;
            waiting += __popc(qm);
            __syncwarp();
        }
        const bool last = idx0 + nwarps * 32 >= n;
#pragma unroll 1
        while (waiting >= 32 || (last && waiting > 0)) {
            const int take = min(32, waiting);
            if (lane < take) phase2(pend[waiting - take + lane]);
            waiting -= take;
            __syncwarp();
        }
    }
}
#endif

constexpr int kDcFrames = 24;     // recursion depth <= log2(band) + 2 (every child band is at most half its parent's)


}  // namespace indelgpu
