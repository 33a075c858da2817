/* TEST INFRASTRUCTURE ONLY -- see indel_oracle.h.
 *
 * Plain-C restatement of the split-read realignment path of indelMINER.
 * Written from the reference's behaviour (file:line cited per function), not
 * from its text: re-entrant (no file-scope state), 0-based sequence pointers,
 * k-mer voting restated set-theoretically (SURVEY.md 8a''), DP cell counters
 * added for the GCUPS denominators.  Differentially pinned against the
 * reference objects in oracle/_ref (tests/test_oracle_vs_reference.py).
 */
#include "indel_oracle.h"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NEG_SENTINEL (-9999999)   /* MININT, localalign.c:3 / globalalign.h:16 */

enum { OP_INS = 1, OP_DEL = 2, OP_SOFT = 4, OP_EQ = 7, OP_X = 8 };   /* bam.h:138-155, readaln.h:10-11 */

static void* xalloc(size_t n)
{
    void* p = calloc(n ? n : 1, 1);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); exit(2); }
    return p;
}

void orc_default_params(orc_params* p)
{
    p->klength = 6; p->numgaps = 0; p->maxdelsize = 1000; p->ethreshold = 10;
    p->match = 1; p->mismatch = -10; p->gapopen = 10; p->gapextend = 10;
}

static inline int subst(const orc_params* p, char a, char b)
{
    /* W[a][b] of localalign.c:61-67: raw byte equality, so 'N' matches 'N' */
    return a == b ? p->match : p->mismatch;
}

/* ------------------------------------------------------------------------ */
/* k-mer diagonal voting                                                     */
/* ------------------------------------------------------------------------ */

/* base2bits of alignment.c:11-24: A/a->0 C/c->1 G/g->2 T/t->3, anything else 0 */
static inline uint32_t base_code(char ch)
{
    switch (ch) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 0;
    }
}

typedef struct { uint32_t code; int pos; } kmer_rec;

static int kmer_cmp(const void* a, const void* b)
{
    const kmer_rec* x = a; const kmer_rec* y = b;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    return x->pos - y->pos;
}

static uint32_t kmer_at(const char* s, int k)
{
    uint32_t c = 0;
    for (int t = 0; t < k; t++) c = (c << 2) | base_code(s[t]);
    return c;
}

/* 1 when find_best_band's forceassert(numdiagonals > numgaps) (alignment.c:403-405, unsigned arithmetic kept) fires */
int orc_would_assert(const orc_params* p, uint32_t N, uint32_t M)
{
    const uint32_t k = (uint32_t)p->klength, g = (uint32_t)p->numgaps;
    const uint32_t numdiag = (N - (k - 1)) + (M - (k - 1));
    return !(numdiag > g);
}

void orc_find_best_band(const orc_params* p,
                        const char* refseq, uint32_t zstart1, uint32_t end1, uint32_t anchor,
                        const char* readseq, uint32_t zstart2, uint32_t end2,
                        int* plow, int* pup)
{
    const uint32_t k = (uint32_t)p->klength, g = (uint32_t)p->numgaps;
    const uint32_t N = end1 - zstart1, M = end2 - zstart2;
    /* alignment.c:403-405 (unsigned arithmetic kept) */
    const uint32_t numdiag = (N - (k - 1)) + (M - (k - 1));
    /* the reference stops the program here (forceassert, alignment.c:405).  The oracle must survive to say so: the
     * stand-alone entry point still exits, orc_realign_read reports ORC_ST_ASSERT (see orc_would_assert) */
    if (!(numdiag > g)) { fprintf(stderr, "oracle: numdiagonals <= numgaps\n"); exit(1); }
    if (M < k) {                                   /* alignment.c:408-412 */
        *plow = (int)(numdiag - 1); *pup = (int)(numdiag - 1);
        return;
    }
    const char* w = refseq + zstart1;
    const char* r = readseq + zstart2;

    /* read k-mers that occur exactly once in the read slice (alignment.c:97-98) */
    int nread = (int)(M - k + 1);
    kmer_rec* recs = xalloc(sizeof(kmer_rec) * (size_t)nread);
    for (int i = 0; i < nread; i++) { recs[i].code = kmer_at(r + i, (int)k); recs[i].pos = i; }
    qsort(recs, (size_t)nread, sizeof(kmer_rec), kmer_cmp);
    int nu = 0;
    for (int i = 0; i < nread; ) {
        int j = i + 1;
        while (j < nread && recs[j].code == recs[i].code) j++;
        if (j == i + 1) recs[nu++] = recs[i];
        i = j;
    }

    int* diag = xalloc(sizeof(int) * (size_t)numdiag);
    if (N >= k) {
        for (uint32_t j = 0; j + k <= N; j++) {      /* every window k-mer (alignment.c:49-65) */
            uint32_t code = kmer_at(w + j, (int)k);
            int lo = 0, hi = nu - 1, hit = -1;
            while (lo <= hi) {
                int mid = (lo + hi) / 2;
                if (recs[mid].code == code) { hit = mid; break; }
                if (recs[mid].code < code) lo = mid + 1; else hi = mid - 1;
            }
            if (hit >= 0) {
                /* alignment.c:102-107: ref offset - read offset + (M - k + 1) */
                uint32_t idx = j - (uint32_t)recs[hit].pos + (M - k + 1);
                if (idx < numdiag) diag[idx] += 1;
            }
        }
    }

    /* bin_bands (alignment.c:130-140) + select_band (:142-181) */
    int a = (int)(anchor - zstart1);
    int best = 0, dist = INT_MAX;
    uint32_t indx = 0;
    for (uint32_t i = 0; i < numdiag; i++) {
        int b = 0;
        if (i < numdiag - g) for (uint32_t j = i; j <= i + g; j++) b += diag[j];
        int dd = abs((int)((uint32_t)a - i));
        if (b > best) { best = b; indx = i; dist = dd; }
        else if (b == best && dd < dist) { indx = i; dist = dd; }
    }
    *plow = (int)(indx - (M - k + 1));               /* alignment.c:438-439 */
    *pup  = (int)(indx + g - (M - k + 1));
    free(diag); free(recs);
}

/* ------------------------------------------------------------------------ */
/* global alignment in a band with the divide-and-conquer script             */
/* ------------------------------------------------------------------------ */

typedef struct {
    const orc_params* p;
    int g, h, m;
    int *cc, *dd, *cp, *dp;          /* per band diagonal: best, best-ending-in-vertical-gap, crossing rows */
    int *mp[3]; signed char *mt[3];  /* per row: previous crossing row / arrival type, per arrival state */
    int *fp; signed char *ft;        /* forward links of the crossing list */
    int* S; int ns; int last;
    orc_cells* cells;
} gctx;

/* script append with merging of adjacent same-sign ops (globalalign.c:40-59) */
static void put_del(gctx* c, int k)
{
    if (c->last < 0) { c->S[c->ns - 1] -= k; c->last = c->S[c->ns - 1]; }
    else { c->S[c->ns++] = -k; c->last = -k; }
}
static void put_ins(gctx* c, int k)
{
    if (c->last > 0) { c->S[c->ns - 1] += k; c->last = c->S[c->ns - 1]; }
    else { c->S[c->ns++] = k; c->last = k; }
}
static void put_rep(gctx* c) { c->S[c->ns++] = 0; c->last = 0; }

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* a1/b1 are "1-based" views: a1[1] is the first symbol (globalalign.c:66-307) */
static int dc_align(gctx* x, const char* a1, const char* b1, int M, int N,
                    int low, int up, int tb, int te)
{
    if (N <= 0) { if (M > 0) put_del(x, M); return -1; }
    if (M <= 0) { put_ins(x, N); return -1; }
    const int band = up - low + 1;
    if (band <= 1) { for (int i = 1; i <= M; i++) put_rep(x); return -1; }

    const int g = x->g, h = x->h, m = x->m;
    int *CC = x->cc, *DD = x->dd, *CP = x->cp, *DP = x->dp;
    const int midd = band / 2 + 1;          /* band index of the mid diagonal        :98  */
    const int rmid = low + midd - 1;        /* its diagonal number                   :99  */
    int leftd = 1 - low, rightd = band;     /* band indices of diagonal 0 / up       :100 */
    int IP = 0;

    /* where did a path into each row-0 cell last touch the mid diagonal?  :102-133 */
    if (leftd < midd) {
        for (int j = 0; j < midd; j++) CP[j] = DP[j] = -1;
        for (int j = midd; j <= rightd; j++) CP[j] = DP[j] = 0;
        x->mp[0][0] = x->mp[1][0] = x->mp[2][0] = -1;
    } else if (leftd > midd) {
        const int fr = leftd - midd;
        for (int j = 0; j <= midd; j++) CP[j] = DP[j] = fr;
        for (int j = midd + 1; j <= rightd; j++) CP[j] = DP[j] = -1;
        x->mp[0][fr] = x->mp[1][fr] = x->mp[2][fr] = -1;
    } else {
        for (int j = 0; j <= rightd; j++) CP[j] = DP[j] = 0;
        x->mp[0][0] = x->mp[1][0] = x->mp[2][0] = -1;
    }

    /* row 0 (:135-146); tb waives the open penalty of a leading gap */
    CC[leftd] = 0;
    {
        int t = (tb == 2) ? 0 : -g;
        for (int j = leftd + 1; j <= rightd; j++) { t -= h; CC[j] = t; DD[j] = t - g; }
    }
    CC[rightd + 1] = DD[rightd + 1] = NEG_SENTINEL;
    DD[leftd] = (tb == 1) ? 0 : -g;
    CC[leftd - 1] = NEG_SENTINEL;

    int c = 0, d = 0, e = 0;
    for (int i = 1; i <= M; i++) {
        if (i > N - up) rightd--;
        if (leftd > 1) leftd--;
        const char ai = a1[i];
        /* leftmost cell of the row (:151-164) */
        {
            int open = CC[leftd + 1] - m, ext = DD[leftd + 1] - h;
            if (open > ext) { d = open; DP[leftd] = CP[leftd + 1]; }
            else            { d = ext;  DP[leftd] = DP[leftd + 1]; }
            int ib = leftd + low - 1 + i;
            c = open;
            if (ib > 0) c = CC[leftd] + subst(x->p, ai, b1[ib]);
            if (d > c || ib <= 0) { c = d; CP[leftd] = DP[leftd]; }
            e = c - g;
            DD[leftd] = d; CC[leftd] = c;
            IP = CP[leftd];
            if (leftd == midd) CP[leftd] = DP[leftd] = IP = i;
            if (x->cells) x->cells->glob++;
        }
        for (int q = leftd + 1; q <= rightd; q++) {
            if (x->cells) x->cells->glob++;
            const int sub = subst(x->p, ai, b1[q + low - 1 + i]);
            if (q != midd) {                                  /* :166-188 */
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
            } else {                                          /* on the mid diagonal :189-232 */
                int open = c - m; e -= h;
                if (open > e) { e = open; x->mp[1][i] = CP[q - 1]; }
                else          { x->mp[1][i] = IP; }
                x->mt[1][i] = 2;
                open = CC[q + 1] - m; d = DD[q + 1] - h;
                if (open > d) { d = open; x->mp[2][i] = CP[q + 1]; }
                else          { x->mp[2][i] = DP[q + 1]; }
                x->mt[2][i] = 1;
                c = CC[q] + sub;
                if (c < d || c < e) {
                    if (e > d) { c = e; x->mp[0][i] = x->mp[1][i]; x->mt[0][i] = 2; }
                    else       { c = d; x->mp[0][i] = x->mp[2][i]; x->mt[0][i] = 1; }
                } else { x->mp[0][i] = i - 1; x->mt[0][i] = 0; }
                if (c - g > e) { x->mp[1][i] = x->mp[0][i]; x->mt[1][i] = x->mt[0][i]; }
                if (c - g > d) { x->mp[2][i] = x->mp[0][i]; x->mt[2][i] = x->mt[0][i]; }
                CP[q] = DP[q] = IP = i;
                CC[q] = c; DD[q] = d;
            }
        }
    }

    /* which state ends the path (:237-248) */
    int k, l;
    if (te == 1 && d + g > c)      { k = DP[rightd]; l = 2; }
    else if (te == 2 && e + g > c) { k = IP;         l = 1; }
    else                           { k = CP[rightd]; l = 0; }
    if (rmid > N - M) l = 2; else if (rmid < N - M) l = 1;
    const int v = c;

    /* walk the crossing rows backwards, leaving forward links (:254-258) */
    int r = -1;
    while (k > -1) {
        x->fp[k] = r; x->ft[k] = (signed char)l;
        r = k;
        int nk = x->mp[l][r]; int nl = x->mt[l][r];
        k = nk; l = nl;
    }

    if (r == -1) {                       /* never touched the mid diagonal (:260-262) */
        if (rmid < 0) dc_align(x, a1, b1, M, N, rmid + 1, up, tb, te);
        else          dc_align(x, a1, b1, M, N, low, rmid - 1, tb, te);
        return v;
    }
    k = r; l = x->fp[k]; int kt = x->ft[k];
    /* first block: origin -> first crossing (:269-275) */
    if (rmid < 0) {
        dc_align(x, a1, b1, r - 1, r + rmid, rmid + 1, imin(up, r + rmid), tb, 1);
        put_del(x, 1);
    } else if (rmid > 0) {
        dc_align(x, a1, b1, r, r + rmid - 1, imax(-r, low), rmid - 1, tb, 2);
        put_ins(x, 1);
    }
    /* one block per consecutive pair of crossings (:278-293) */
    const int t2 = up - rmid - 1, t3 = low - rmid + 1;
    while (l > -1) {
        if (kt == 0) put_rep(x);
        else if (kt == 1) {              /* round the right-hand triangle */
            put_ins(x, 1);
            int t1 = l - k - 1;
            dc_align(x, a1 + k, b1 + k + rmid + 1, t1, t1, 0, imin(t1, t2), 2, 1);
            put_del(x, 1);
        } else {                         /* round the left-hand triangle */
            put_del(x, 1);
            int t1 = l - k - 1;
            dc_align(x, a1 + k + 1, b1 + k + rmid, t1, t1, imax(-t1, t3), 0, 1, 2);
            put_ins(x, 1);
        }
        k = l; l = x->fp[k]; kt = x->ft[k];
    }
    /* last block: last crossing -> (M,N) (:296-304) */
    if (N - M > rmid) {
        put_ins(x, 1);
        int t1 = k + rmid + 1;
        dc_align(x, a1 + k, b1 + t1, M - k, N - t1, 0, imin(N - t1, t2), 2, te);
    } else if (N - M < rmid) {
        put_del(x, 1);
        int t1 = M - (k + 1);
        dc_align(x, a1 + k + 1, b1 + k + rmid, t1, N - (k + rmid), imax(-t1, t3), 0, 1, te);
    }
    return v;
}

int orc_global_align(const orc_params* p, const char* A, const char* B, int M, int N,
                     int low, int up, int* S, int* nS, orc_cells* cells)
{
    gctx x; memset(&x, 0, sizeof(x));
    x.p = p; x.g = p->gapopen; x.h = p->gapextend; x.m = x.g + x.h;
    x.S = S; x.ns = 0; x.last = 0; x.cells = cells;
    /* band widened to hold diagonals 0 and N-M (globalalign.c:347-348) */
    low = imin(imax(-M, low), imin(N - M, 0));
    up  = imax(imin(N, up), imax(N - M, 0));
    int score;
    if (N <= 0) {                                    /* :350-353 */
        if (M > 0) put_del(&x, M);
        score = M <= 0 ? 0 : -(x.g + x.h * M);
    } else if (M <= 0) {                             /* :354-357 */
        put_ins(&x, N);
        score = -(x.g + x.h * N);
    } else if (up - low + 1 <= 1) {                  /* :358-365 */
        score = 0;
        for (int i = 0; i < M; i++) { put_rep(&x); score += subst(p, A[i], B[i]); }
    } else {
        const int band = up - low + 1;
        size_t nb = (size_t)(band + 2) * sizeof(int), nr = (size_t)(M + 1);
        x.cc = xalloc(nb); x.dd = xalloc(nb); x.cp = xalloc(nb); x.dp = xalloc(nb);
        for (int t = 0; t < 3; t++) { x.mp[t] = xalloc(nr * sizeof(int)); x.mt[t] = xalloc(nr); }
        x.fp = xalloc(nr * sizeof(int)); x.ft = xalloc(nr);
        score = dc_align(&x, A - 1, B - 1, M, N, low, up, 0, 0);
        free(x.cc); free(x.dd); free(x.cp); free(x.dp);
        for (int t = 0; t < 3; t++) { free(x.mp[t]); free(x.mt[t]); }
        free(x.fp); free(x.ft);
    }
    *nS = x.ns;
    return score;
}

/* ------------------------------------------------------------------------ */
/* local alignment in a band: end point (forward), start point (reverse)      */
/* ------------------------------------------------------------------------ */

int orc_local_align(const orc_params* p, const char* seq1, int M, const char* seq2, int N,
                    int low, int up, int* psi, int* psj, int* pei, int* pej,
                    int* S, int* nS, orc_cells* cells)
{
    const int G = p->gapopen, H = p->gapextend, m = G + H;
    *nS = 0;
    low = imax(-M, low);                              /* localalign.c:70-71 */
    up  = imin(N, up);
    const int band = up - low + 1;
    if (band < 1) { fprintf(stderr, "oracle: low > up (low:%d up:%d)\n", low, up); exit(1); }

    /* rolling rows indexed by diagonal offset t = (j - i) - low, one slack slot each side */
    int* Hp = xalloc(sizeof(int) * (size_t)(band + 2));
    int* Dp = xalloc(sizeof(int) * (size_t)(band + 2));
    int* Hn = xalloc(sizeof(int) * (size_t)(band + 2));
    int* Dn = xalloc(sizeof(int) * (size_t)(band + 2));
#define AT(arr, t) ((arr)[(t) + 1])

    const int si = imax(0, -up), ei = imin(M, N - low);
    for (int t = -1; t <= band; t++) { AT(Hp, t) = NEG_SENTINEL; AT(Dp, t) = NEG_SENTINEL; }
    for (int t = 0; t < band; t++) {                  /* row si (:88-99) */
        int j = si + low + t;
        if (j >= 0 && j <= N) { AT(Hp, t) = 0; AT(Dp, t) = -G; }
    }
    int best = 0, endi = si, endj = si + low;
    for (int i = si + 1; i <= ei; i++) {              /* forward (:100-131) */
        for (int t = -1; t <= band; t++) { AT(Hn, t) = NEG_SENTINEL; AT(Dn, t) = NEG_SENTINEL; }
        int tlo = imax(0, -i - low), thi = imin(band - 1, N - i - low);
        int e = NEG_SENTINEL, left = NEG_SENTINEL;
        for (int t = tlo; t <= thi; t++) {
            int j = i + low + t;
            int d = imax(AT(Hp, t + 1) - m, AT(Dp, t + 1) - H);   /* vertical: (i-1, j) */
            int c;
            if (j == 0) c = d;
            else {
                c = AT(Hp, t) + subst(p, seq1[i - 1], seq2[j - 1]);
                if (t > tlo) { e = imax(left - m, e - H); if (e > c) c = e; }
                if (d > c) c = d;
            }
            if (c < 0) c = 0;
            if (t == tlo) e = c - G;
            left = c;
            AT(Hn, t) = c; AT(Dn, t) = d;
            if (cells) cells->fwd++;
            if (c > best) { best = c; endi = i; endj = j; }   /* strict: first max, row-major */
        }
        int* tmp = Hp; Hp = Hn; Hn = tmp; tmp = Dp; Dp = Dn; Dn = tmp;
    }

    int starti = 0, startj = 0, found = 0;
    if (best > 0) {
        /* reverse, global-style from the virtual cell (endi+1, endj+1) (:132-176) */
        const int tend = (endj - endi) - low;             /* diagonal of the end cell */
        for (int t = -1; t <= band; t++) { AT(Hp, t) = NEG_SENTINEL; AT(Dp, t) = NEG_SENTINEL; }
        {
            int tl = imax(0, -endi - low);                /* leading horizontal gap row */
            AT(Hp, tend) = 0; AT(Dp, tend) = -G;
            int acc = -G;
            for (int t = tend - 1; t >= tl; t--) { acc -= H; AT(Hp, t) = acc; AT(Dp, t) = acc - G; }
        }
        for (int i = endi; i >= 1 && !found; i--) {
            for (int t = -1; t <= band; t++) { AT(Hn, t) = NEG_SENTINEL; AT(Dn, t) = NEG_SENTINEL; }
            int thi = imin(band - 1, tend + (endi - i) + 1);   /* column endj+1 or the band edge */
            int tlo = imax(0, 1 - i - low);                     /* column 1 */
            int e = NEG_SENTINEL, right = NEG_SENTINEL;
            for (int t = thi; t >= tlo; t--) {
                int j = i + low + t;
                int d = imax(AT(Hp, t - 1) - m, AT(Dp, t - 1) - H);   /* vertical: (i+1, j) */
                int c;
                if (t == thi) {
                    c = (j <= N) ? AT(Hp, t) + subst(p, seq1[i - 1], seq2[j - 1]) : NEG_SENTINEL;
                    if (j > N) c = AT(Hp, t - 1) - m;
                    if (d > c) c = d;
                    e = c - G;
                } else {
                    e = imax(right - m, e - H);
                    c = AT(Hp, t) + subst(p, seq1[i - 1], seq2[j - 1]);
                    if (e > c) c = e;
                    if (d > c) c = d;
                }
                right = c;
                AT(Hn, t) = c; AT(Dn, t) = d;
                if (cells) cells->rev++;
                if (c == best) { starti = i; startj = j; found = 1; break; }
            }
            int* tmp = Hp; Hp = Hn; Hn = tmp; tmp = Dp; Dp = Dn; Dn = tmp;
        }
    }
#undef AT
    free(Hp); free(Dp); free(Hn); free(Dn);
    if (best <= 0 || !found) return 0;      /* the reference returns garbage <= 0 here; caller discards */
    if (starti < 0 || starti > M || startj < 0 || startj > N) return 0;   /* :180-185 */
    *psi = starti; *psj = startj; *pei = endi; *pej = endj;
    if (endi - starti == 0 || endj - startj == 0) return 0;                /* :191-193 */
    return orc_global_align(p, seq1 + starti - 1, seq2 + startj - 1,
                            endi - starti + 1, endj - startj + 1,
                            low - (startj - starti), up - (startj - starti), S, nS, cells);
}

/* ------------------------------------------------------------------------ */
/* script -> CIGAR                                                            */
/* ------------------------------------------------------------------------ */

static void push_op(uint32_t* cigar, int* n, int op, int len)
{
    cigar[(*n)++] = ((uint32_t)len << 4) | (uint32_t)op;
}

int orc_fetch_cigar(const char* A, const char* B, int M, int N, const int* S,
                    int AP, int readlength, uint32_t* cigar, int* pnumops)
{
    int n = 0, i = 0, j = 0, mm = 0;
    int clip = AP - 1;                               /* globalalign.c:521-526 */
    if (clip > 0) push_op(cigar, &n, OP_SOFT, clip);
    int run_op = -1, run_len = 0, total = clip, pending = 0;
    while (i < M || j < N) {                         /* :532-591, one column at a time */
        int op;
        if (pending == 0 && *S == 0) { S++; op = (A[i] == B[j]) ? OP_EQ : OP_X; if (op == OP_X) mm++; i++; j++; }
        else {
            if (pending == 0) pending = *S++;
            if (pending > 0) { pending--; j++; op = OP_DEL; }
            else             { pending++; i++; op = OP_INS; }
        }
        if (run_op != -1 && run_op != op) { push_op(cigar, &n, run_op, run_len); total += run_len; run_len = 0; }
        run_op = op; run_len++;
    }
    if (run_op != -1 && run_len > 0) { push_op(cigar, &n, run_op, run_len); total += run_len; }
    if (total < readlength) push_op(cigar, &n, OP_SOFT, readlength - total);   /* :598-601 */
    *pnumops = n;
    return mm;
}

int orc_attempt_band_alignment(const orc_params* p,
                               const char* refseq, uint32_t zstart1, uint32_t end1,
                               const char* readseq, uint32_t zstart2, uint32_t end2,
                               int low, int up, int* pr1, int* pr2, int* pq1, int* pq2,
                               uint32_t* cigar, int* pscore, orc_cells* cells)
{
    const int N = (int)(end1 - zstart1), M = (int)(end2 - zstart2);
    int* S = xalloc(sizeof(int) * (size_t)(N + M + 2));
    int nS = 0, q1 = 0, r1 = 0, q2 = 0, r2 = 0, n = 0;
    int score = orc_local_align(p, readseq + zstart2, M, refseq + zstart1, N, low, up,
                                &q1, &r1, &q2, &r2, S, &nS, cells);
    if (pscore) *pscore = score > 0 ? score : 0;
    if (score <= 0) {                                /* alignment.c:365-372 */
        *pr1 = *pr2 = *pq1 = *pq2 = 0; free(S); return 0;
    }
    orc_fetch_cigar(readseq + zstart2 + q1 - 1, refseq + zstart1 + r1 - 1,
                    q2 - q1 + 1, r2 - r1 + 1, S, q1, M, cigar, &n);
    *pr1 = r1 + (int)zstart1 - 1; *pr2 = r2 + (int)zstart1;      /* :385-388 */
    *pq1 = q1 + (int)zstart2 - 1; *pq2 = q2 + (int)zstart2;
    free(S);
    return n;
}

/* ------------------------------------------------------------------------ */
/* the two-round driver                                                       */
/* ------------------------------------------------------------------------ */

static inline int cig_op(uint32_t c) { return (int)(c & 15u); }
static inline int cig_len(uint32_t c) { return (int)(c >> 4); }

/* alignment.c:219-303 */
static int count_matches(const uint32_t* c1, int n1, int q1, int q2,
                         const uint32_t* c2, int n2, int q3, int q4, int* pmm)
{
    int i, j, matches = 0, mm = 0;
    for (i = 0, j = q1; i < n1; i++) {
        int len = cig_len(c1[i]), op = cig_op(c1[i]);
        if (op != OP_DEL) j += len;
        if (j < q2) { if (op == OP_EQ) matches += len; else if (op == OP_X) mm += len; }
        if (j >= q2) {
            if (op == OP_EQ) matches += q2 - (j - len); else if (op == OP_X) mm += q2 - (j - len);
            break;
        }
    }
    for (i = 0, j = 0; i < n2; i++) {
        int len = cig_len(c2[i]), op = cig_op(c2[i]);
        if (op != OP_DEL) j += len;
        if (j >= q3) {
            if (op == OP_EQ) matches += j - q3; else if (op == OP_X) mm += j - q3;
            i++; break;
        }
    }
    for (; i < n2; i++) {
        int len = cig_len(c2[i]), op = cig_op(c2[i]);
        if (op != OP_DEL) j += len;
        if (j < q4) { if (op == OP_EQ) matches += len; else if (op == OP_X) mm += len; }
        if (j >= q4) {
            if (op == OP_EQ) matches += q4 - (j - len); else if (op == OP_X) mm += q4 - (j - len);
            break;
        }
    }
    *pmm = mm;
    return matches;
}

/* alignment.c:306-339 */
static int best_junction(int q1, int q2, const uint32_t* c1, int n1,
                         int q3, int q4, const uint32_t* c2, int n2, int readlength)
{
    int bestm = 0, bestmm = INT_MAX, index = -1;
    for (int i = q3; i <= q2; i++) {
        int mm, matches = count_matches(c1, n1, q1, i, c2, n2, i, q4, &mm);
        if (matches > bestm || (matches == bestm && mm < bestmm)) { bestm = matches; bestmm = mm; index = i; }
        if (matches == readlength && mm == 0) break;
    }
    if (index == -1) { fprintf(stderr, "oracle: no junction candidate\n"); exit(1); }
    return index;
}

/* alignment.c:478-532 */
static void prefix_clip(int len, uint32_t* cig, int* n)
{
    if (len == 0) return;
    if (cig_op(cig[0]) == OP_SOFT) cig[0] = ((uint32_t)(cig_len(cig[0]) + len) << 4) | OP_SOFT;
    else { memmove(cig + 1, cig, sizeof(uint32_t) * (size_t)(*n)); cig[0] = ((uint32_t)len << 4) | OP_SOFT; (*n)++; }
}
static void suffix_clip(int len, uint32_t* cig, int* n)
{
    if (len == 0) return;
    if (cig_op(cig[*n - 1]) == OP_SOFT) cig[*n - 1] = ((uint32_t)(cig_len(cig[*n - 1]) + len) << 4) | OP_SOFT;
    else { cig[*n] = ((uint32_t)len << 4) | OP_SOFT; (*n)++; }
}

/* new_readseg coordinate bookkeeping (readaln.c:24-99) */
static void emit_seg(orc_result* o, uint32_t cig, int* refindx, int* readindx)
{
    int op = cig_op(cig), len = cig_len(cig), n = o->nseg;
    if (n >= ORC_MAXSEG) { fprintf(stderr, "oracle: too many segments\n"); exit(1); }
    o->seg_op[n] = op; o->seg_len[n] = len; o->seg_start[n] = *refindx;
    switch (op) {
        case OP_EQ: case OP_X: *readindx += len; *refindx += len; break;
        case OP_INS: case OP_SOFT: *readindx += len; break;
        case OP_DEL: *refindx += len; break;
        default: fprintf(stderr, "oracle: unhandled cigar op %d\n", op); exit(1);
    }
    o->seg_end[n] = *refindx;
    o->nseg = n + 1;
}

/* readaln.c:348-458 */
static void stitch_segments(orc_result* o, int r1, const uint32_t* c1, int n1, int index,
                            int q2, int r2, const uint32_t* c2, int n2)
{
    int i, j, refindx = r1, readindx = 0;
    o->nseg = 0;
    for (i = 0, j = 0; i < n1; i++) {
        int op = cig_op(c1[i]), len = cig_len(c1[i]);
        if (op != OP_DEL) j += len;
        if (j <= index) emit_seg(o, c1[i], &refindx, &readindx);
        if (j > index) {
            int part = index - (j - len);
            if (part > 0) emit_seg(o, ((uint32_t)part << 4) | (uint32_t)op, &refindx, &readindx);
            break;
        }
    }
    int rindex = r2, nextindex = index;
    if (index >= q2) {
        int offset = 0;
        for (i = 0, j = 0; i < n2; i++) {
            int op = cig_op(c2[i]), len = cig_len(c2[i]);
            if (op != OP_DEL) j += len;
            if (j <= q2) { }
            else if (j <= index) {
                if (op != OP_INS) { offset += len; if ((j - len) <= q2) offset -= q2 - (j - len); }
            } else {
                if (op != OP_INS && (j - len) <= index) offset += index - (j - len);
            }
        }
        rindex = r2 + offset;
    } else {
        emit_seg(o, ((uint32_t)(q2 - index) << 4) | OP_INS, &refindx, &readindx);
        nextindex += q2 - index;
    }
    if (refindx < rindex) emit_seg(o, ((uint32_t)(rindex - refindx) << 4) | OP_DEL, &refindx, &readindx);
    for (i = 0, j = 0; i < n2; i++) {
        int op = cig_op(c2[i]), len = cig_len(c2[i]);
        if (op != OP_DEL) j += len;
        if (j > nextindex) {
            emit_seg(o, ((uint32_t)(j - nextindex) << 4) | (uint32_t)op, &refindx, &readindx);
            i++; break;
        }
    }
    for (; i < n2; i++) emit_seg(o, c2[i], &refindx, &readindx);
}

static void count_evidence(orc_result* o)
{
    o->nevidence = 0;
    for (int i = 0; i < o->nseg; i++)
        if (o->seg_op[i] == OP_DEL || o->seg_op[i] == OP_INS) o->nevidence++;
}

void orc_realign_read(const orc_params* p, const char* refseq, int reflength,
                      int position, int range1, const char* read, int readlen,
                      orc_result* o, orc_cells* cells)
{
    memset(o, 0, sizeof(*o));
    /* windows: alignment.c:774-783 (int32 arithmetic; maxdelsize is unsigned there but the
       sum is assigned to an int32_t) */
    int32_t distance = range1;
    const int32_t left1  = position >= distance ? position - distance : 0;
    const int32_t right1 = reflength < (position + distance) ? reflength : position + distance;
    distance = (int32_t)((unsigned)range1 + (unsigned)p->maxdelsize);
    const int32_t left2  = position >= distance ? position - distance : 0;
    const int32_t right2 = reflength < (position + distance) ? reflength : position + distance;
    const int32_t anchor = position;
    const unsigned readlength = (unsigned)readlen;
    const unsigned ethreshold = (unsigned)p->ethreshold;
    (void)right1;

    /* round 1 (alignment.c:555-566) */
    if (orc_would_assert(p, (uint32_t)right1 - (uint32_t)left1, readlength)) { o->status = ORC_ST_ASSERT; return; }
    orc_find_best_band(p, refseq, (uint32_t)left1, (uint32_t)right1, (uint32_t)anchor,
                       read, 0, readlength, &o->low1, &o->up1);
    int r1, r2, q1, q2;
    o->n1 = orc_attempt_band_alignment(p, refseq, (uint32_t)left1, (uint32_t)right1, read, 0, readlength,
                                       o->low1, o->up1, &r1, &r2, &q1, &q2, o->cigar1, &o->score1, cells);
    o->r1 = r1; o->r2 = r2; o->q1 = q1; o->q2 = q2;
    if (q1 == q2) { o->status = ORC_ST_UNALIGNED; return; }
    if (q1 == 0 && q2 == (int)readlength) {          /* :575-582 */
        stitch_segments(o, r1, o->cigar1, o->n1, (int)readlength, 0, -1, NULL, 0);
        count_evidence(o);
        o->status = ORC_ST_WHOLE;
        return;
    }
    /* leading / trailing '=' run (:584-599) */
    unsigned f_nonmatch, l_nonmatch;
    {
        int i, j;
        for (i = 0, j = 0; i < o->n1; i++) {
            int op = cig_op(o->cigar1[i]);
            if (i == 0 && op == OP_SOFT) continue;
            if (op != OP_EQ) break;
            j += cig_len(o->cigar1[i]);
        }
        f_nonmatch = (unsigned)j;
        for (i = o->n1 - 1, j = 0; i >= 0; i--) {
            int op = cig_op(o->cigar1[i]);
            if (i == o->n1 - 1 && op == OP_SOFT) continue;
            if (op != OP_EQ) break;
            j += cig_len(o->cigar1[i]);
        }
        l_nonmatch = (unsigned)j;
    }

    /* round 2 (:601-717); mixed signed/unsigned guards kept with the reference's types */
    uint32_t zs1, e1, anc, zs2, e2; int tail;   /* tail: 1 = slice is the read's tail (prefix clip) */
    if (r1 > anchor) {
        if (q1 == 0) {
            if (((readlength - f_nonmatch) < ethreshold) || ((right2 - r1 - f_nonmatch) < ethreshold)) { o->status = ORC_ST_SHORT; return; }
            zs1 = (uint32_t)r1 + f_nonmatch; e1 = (uint32_t)right2; anc = (uint32_t)r1; zs2 = f_nonmatch; e2 = readlength; tail = 1;
        } else if (q2 == (int)readlength) {
            if (((readlength - l_nonmatch) < ethreshold) || ((r2 - l_nonmatch - anchor) < ethreshold)) { o->status = ORC_ST_SHORT; return; }
            zs1 = (uint32_t)anchor; e1 = (uint32_t)r2 - l_nonmatch; anc = (uint32_t)r2; zs2 = 0; e2 = readlength - l_nonmatch; tail = 0;
        } else { o->status = ORC_ST_NOBRANCH; return; }
    } else if (r1 < anchor) {
        if (r2 >= anchor) { o->status = ORC_ST_NOBRANCH; return; }
        if (q1 == 0) {
            if (((readlength - f_nonmatch) < ethreshold) || ((anchor - r1 - f_nonmatch) < ethreshold)) { o->status = ORC_ST_SHORT; return; }
            zs1 = (uint32_t)r1 + f_nonmatch; e1 = (uint32_t)anchor; anc = (uint32_t)r1; zs2 = f_nonmatch; e2 = readlength; tail = 1;
        } else if (q2 == (int)readlength) {
            if (((readlength - l_nonmatch) < ethreshold) || ((r2 - l_nonmatch - left2) < ethreshold)) { o->status = ORC_ST_SHORT; return; }
            zs1 = (uint32_t)left2; e1 = (uint32_t)r2 - l_nonmatch; anc = (uint32_t)r2; zs2 = 0; e2 = readlength - l_nonmatch; tail = 0;
        } else { o->status = ORC_ST_NOBRANCH; return; }
    } else { o->status = ORC_ST_NOBRANCH; return; }

    if (orc_would_assert(p, e1 - zs1, e2 - zs2)) { o->status = ORC_ST_ASSERT; o->nseg = 0; return; }
    orc_find_best_band(p, refseq, zs1, e1, anc, read, zs2, e2, &o->low2, &o->up2);
    int r3, r4, q3, q4;
    o->n2 = orc_attempt_band_alignment(p, refseq, zs1, e1, read, zs2, e2, o->low2, o->up2,
                                       &r3, &r4, &q3, &q4, o->cigar2, &o->score2, cells);
    o->r3 = r3; o->r4 = r4; o->q3 = q3; o->q4 = q4;
    if (tail) {
        if (q4 != (int)readlength || q3 == q4) { o->status = ORC_ST_R2FAIL; return; }
        prefix_clip((int)f_nonmatch, o->cigar2, &o->n2);
    } else {
        if (q3 != 0 || q3 == q4) { o->status = ORC_ST_R2FAIL; return; }
        suffix_clip((int)l_nonmatch, o->cigar2, &o->n2);
    }

    /* combine (:719-758) */
    o->index = -1;
    if (q1 > q3 && q1 <= q4) {
        o->index = best_junction(q3, q4, o->cigar2, o->n2, q1, q2, o->cigar1, o->n1, (int)readlength);
        stitch_segments(o, r3, o->cigar2, o->n2, o->index, q1, r1, o->cigar1, o->n1);
    } else if (q3 > q1 && q3 <= q2) {
        o->index = best_junction(q1, q2, o->cigar1, o->n1, q3, q4, o->cigar2, o->n2, (int)readlength);
        stitch_segments(o, r1, o->cigar1, o->n1, o->index, q3, r3, o->cigar2, o->n2);
    } else if (q1 > q4 && r1 == r4) {
        o->index = q4;
        stitch_segments(o, r3, o->cigar2, o->n2, q4, q1, r1, o->cigar1, o->n1);
    } else if (q3 > q2 && r2 == r3) {
        o->index = q2;
        stitch_segments(o, r1, o->cigar1, o->n1, q2, q3, r3, o->cigar2, o->n2);
    } else { o->status = ORC_ST_NOCOMBINE; return; }
    count_evidence(o);
    o->status = ORC_ST_SPLIT;
}

/* batch loop for bench.py's cpu_baseline "port" leg (used only when oracle/_ref is absent) */
long orc_realign_batch(const orc_params* p, const char* refseq, int reflength, int n,
                       const char* reads, const long long* off, const int* position,
                       const int* range1, int* nseg_out)
{
    long total = 0;
    orc_result* o = xalloc(sizeof(orc_result));
    for (int i = 0; i < n; i++) {
        orc_realign_read(p, refseq, reflength, position[i], range1[i], reads + off[i],
                         (int)(off[i + 1] - off[i]), o, NULL);
        if (nseg_out) nseg_out[i] = o->nseg;
        total += o->nseg;
    }
    free(o);
    return total;
}


/* ======================================================================================
 * row f1: realign_with_indel (variant.c:1246-1424)
 * ====================================================================================== */
#include <ctype.h>

char* orc_indel_target(const char* reference, int rstart, int rstop, int vtype, int vstart, int vstop,
                       const char* alternate)
{
    const size_t len0 = (size_t)(rstop - rstart);
    size_t cap = len0 + 1;
    char* target = calloc(cap, 1);                               /* copy_partial_string, strings.c:3-8 */
    memcpy(target, reference + rstart, len0);
    if (vtype == 1) {                                            /* DELETION, variant.c:1261-1264 */
        memmove(target + vstart - rstart, target + vstop - rstart - 1, (size_t)(rstop - vstop + 2));
    } else {                                                     /* INSERTION, :1265-1271 */
        const size_t alen = strlen(alternate) - 1;
        target = realloc(target, cap + alen);
        memset(target + cap, 0, alen);                           /* ckreallocz, memalloc.c:45-55 */
        char* at = target + vstart - rstart;
        memmove(at + alen, at, strlen(at));
        memcpy(at, alternate + 1, alen);
    }
    return target;
}

void orc_indel_support_dp(const char* t1, int len1, const char* t2, int len2,
                          int* psubs, int* pindels, int* paligned, long long* cells)
{
    enum { SUB = 0, INS = 1, DEL = 2 };
    const int match = 2, mismatch = 1, gopen = 4, gextend = 1;   /* variant.c:1289-1292 */
    const size_t W = (size_t)len1 + 1;
    int* V = calloc(((size_t)len2 + 1) * W, sizeof(int));
    char* I = calloc(((size_t)len2 + 1) * W, 1);
    int* F = calloc(W, sizeof(int));                             /* starts at 0 (:1305) */
    for (int i = 0; i <= len2; i++) V[(size_t)i * W] = -gopen - i * gextend;
    for (int j = 0; j <= len1; j++) V[j] = -gopen - j * gextend;
    int max_score = 0, max_i = -1, max_j = -1;
    for (int i = 1; i <= len2; i++) {
        int E = 0;                                               /* restarts on every row (:1322) */
        for (int j = 1; j <= len1; j++) {
            int ifsub = V[(size_t)(i - 1) * W + j - 1];
            ifsub = toupper((unsigned char)t1[j - 1]) == toupper((unsigned char)t2[i - 1]) ? ifsub + match : ifsub - mismatch;
            const int vup = V[(size_t)(i - 1) * W + j] - gopen;
            const int ifins = (F[j] > vup ? F[j] : vup) - gextend;
            F[j] = ifins;
            const int vleft = V[(size_t)i * W + j - 1] - gopen;
            const int ifdel = (E > vleft ? E : vleft) - gextend;
            E = ifdel;
            const int ifindel = ifins > ifdel ? ifins : ifdel;
            int v = ifsub; char d = SUB;
            if (v < ifindel) { d = (ifins >= ifdel) ? INS : DEL; v = ifindel; }   /* :1336-1342 */
            V[(size_t)i * W + j] = v; I[(size_t)i * W + j] = d;
            if (v > max_score) { max_score = v; max_i = i; max_j = j; }          /* strict: first maximum */
        }
    }
    if (cells) *cells += (long long)len1 * len2;
    /* traceback while the score stays positive (:1377-1398), counting as :1403-1417 does: every column
       plus the terminating NUL, which is neither a gap nor a mismatch */
    int subs = 0, ins = 0, dels = 0, aligned = 1;
    int score = max_score, i = max_i, j = max_j;
    while (score > 0) {
        const char d = I[(size_t)i * W + j];
        if (d == SUB) { if (t1[j - 1] != t2[i - 1]) subs++; aligned++; i--; j--; }
        else if (d == INS) { ins++; aligned++; i--; }
        else { dels++; j--; }
        score = V[(size_t)i * W + j];
    }
    free(V); free(I); free(F);
    *psubs = subs; *pindels = ins + dels; *paligned = aligned;
}

void orc_realign_with_indel(const char* reference, int rstart, int rstop, const char* query, int qstart,
                            int qstop, int vtype, int vstart, int vstop, const char* alternate,
                            int* subs, int* indels, int* aligned)
{
    char* target = orc_indel_target(reference, rstart, rstop, vtype, vstart, vstop, alternate);
    const char* t2 = query + qstart;
    int len2 = qstop - qstart;
    if ((int)strlen(t2) < len2) len2 = (int)strlen(t2);          /* :1281-1283 */
    orc_indel_support_dp(target, (int)strlen(target), t2, len2, subs, indels, aligned, NULL);
    free(target);
}
