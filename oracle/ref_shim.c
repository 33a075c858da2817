/* TEST INFRASTRUCTURE ONLY -- never linked into the product (libindelgpu.so).
 *
 * Thin shim around the UNMODIFIED reference translation unit src/alignment.c
 * (found through -I$(REF)/src at build time; nothing is copied into this repo).
 * alignment.c keeps find_best_band / attempt_band_alignment /
 * attempt_diagonal_alignments `static`, so the only way to reach them without
 * editing the reference is to #include the .c file and export wrappers.
 *
 * Built by oracle/Makefile into oracle/_ref/libref_align.so (and, with
 * -DREFSHIM_TRACE, into the object that replaces alignment.o inside
 * oracle/_ref/indelminer_trace to record golden vectors from test_data).
 *
 * Interception: add_evidence_from_segment() (alignment.c:449-476) frees the
 * final segment list through free_readsegs(); we rename that one call inside
 * this TU so the list can be captured before it is released.
 */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

#define free_readsegs refshim_free_readsegs
#ifdef REFSHIM_TRACE
#define local_align refshim_local_align
#define attempt_pe_alignment refshim_inner_attempt_pe_alignment
#endif

#include "alignment.c"

#undef free_readsegs
#ifdef REFSHIM_TRACE
#undef local_align
#undef attempt_pe_alignment
#endif

void free_readsegs(readseg** prs);

#ifndef REFSHIM_TRACE
/* the seven globals alignment.c reads (alignment.c:3-9, defined in indelminer.c:31-44) */
uint klength = 6;
uint numgaps = 0;
uint maxdelsize = 1000;
bool debug_flag = FALSE;
FILE* debug_file = NULL;
uint32_t seed_mask = 1023;
uint ethreshold = 10;
#endif

/* ---- capture of the final segment list ---------------------------------- */
#define REFSHIM_MAXSEG 4096
static int cap_n = -1;
static int cap_op[REFSHIM_MAXSEG], cap_len[REFSHIM_MAXSEG];
static int cap_start[REFSHIM_MAXSEG], cap_end[REFSHIM_MAXSEG];
static int cap_armed = 0;

void refshim_free_readsegs(readseg** prs)
{
    if (cap_armed) {
        int n = 0;
        for (readseg* it = *prs; it && n < REFSHIM_MAXSEG; it = it->next, n++) {
            cap_op[n] = it->op; cap_len[n] = it->oplen;
            cap_start[n] = it->start; cap_end[n] = it->end;
        }
        cap_n = n;
    }
    free_readsegs(prs);
}

#ifndef REFSHIM_TRACE
void refshim_set_params(unsigned k, unsigned gaps, unsigned maxdel, unsigned ethr)
{
    klength = k; numgaps = gaps; maxdelsize = maxdel; ethreshold = ethr;
    seed_mask = (uint32_t)((1ULL << (2 * (k - 1))) - 1);   /* indelminer.c:1071 */
}

void refshim_find_best_band(char* ref, unsigned zs1, unsigned e1, unsigned anchor,
                            char* read, unsigned zs2, unsigned e2, int* low, int* up)
{
    find_best_band(ref, zs1, e1, anchor, read, zs2, e2, low, up);
}

/* returns numcigarops (0 when no alignment); cigar copied into out (cap maxc) */
int refshim_attempt_band_alignment(char* ref, unsigned zs1, unsigned e1,
                                   char* read, unsigned zs2, unsigned e2,
                                   int low, int up, int* r1, int* r2, int* q1, int* q2,
                                   uint32_t* out, int maxc)
{
    int n = 0;
    uint32_t* cig = ckallocz(sizeof(uint32_t));
    attempt_band_alignment(ref, zs1, e1, read, zs2, e2, low, up, r1, r2, q1, q2, &n, &cig);
    if (*q1 == *q2) n = 0;
    for (int i = 0; i < n && i < maxc; i++) out[i] = cig[i];
    ckfree(cig);
    return n;
}

/* Full two-round realignment of one read.  Mirrors attempt_pe_alignment's
 * window arithmetic (alignment.c:775-783) but takes reflength as an argument so
 * large contigs need no NUL terminator / strlen.  Returns the number of final
 * segments (0 = NULL result) and the number of evidence records in *nev. */
int refshim_realign(char* refseq, int reflength, int position, int range1, char* read,
                    int* seg_op, int* seg_len, int* seg_start, int* seg_end, int maxseg,
                    int* nev)
{
    int32_t left1, right1, left2, right2, distance;
    distance = range1;
    left1  = position >= distance ? position - distance : 0;
    right1 = reflength < (position + distance) ? reflength : position + distance;
    distance = range1 + maxdelsize;
    left2  = position >= distance ? position - distance : 0;
    right2 = reflength < (position + distance) ? reflength : position + distance;

    readaln rln; memset(&rln, 0, sizeof(rln));
    rln.qname = "r"; rln.tid = -1; rln.strand = '+'; rln.index = '1'; rln.qual = 60;
    uint32_t cigar = ((uint32_t)strlen(read) << BAM_CIGAR_SHIFT) + BAM_CSOFT_CLIP;
    int refindx = 0, readindx = 0;
    rln.segments = new_readseg(read, cigar, &refindx, &readindx);
    rln.segments->start = -1; rln.segments->end = -1;

    cap_n = -1; cap_armed = 1;
    evidence* ev = attempt_diagonal_alignments(&rln, refseq, left1, right1, left2, right2,
                                               position, rln.segments->sequence);
    cap_armed = 0;
    int ne = 0;
    for (evidence* e = ev; e; e = e->next) ne++;
    *nev = ne;
    int n = cap_n < 0 ? 0 : cap_n;
    if (cap_n < 0 && rln.segments) free_readsegs(&rln.segments);
    for (int i = 0; i < n && i < maxseg; i++) {
        seg_op[i] = cap_op[i]; seg_len[i] = cap_len[i];
        seg_start[i] = cap_start[i]; seg_end[i] = cap_end[i];
    }
    /* evidence records are leaked on purpose (test process; free_used_evidence
       lives in evidence.c but needs isused bookkeeping) */
    return n;
}

/* CPU-baseline loops for bench.py: the reference's own code over a batch, one call.
 * as_shipped = 1 goes through attempt_pe_alignment itself, i.e. pays the per-read
 * strlen(contig) of alignment.c:771 (refseq must be NUL-terminated);
 * as_shipped = 0 calls attempt_diagonal_alignments with the windows precomputed (function level). */
long refshim_realign_batch(char* refseq, int reflength, int n, const char* reads, const long long* off,
                           const int* position, const int* range1, int as_shipped, int* nseg_out)
{
    long total = 0;
    char* sequences[1] = { refseq };
    for (int i = 0; i < n; i++) {
        int len = (int)(off[i + 1] - off[i]);
        char* read = ckallocz(len + 1);
        memcpy(read, reads + off[i], len);
        int ns;
        if (as_shipped) {
            readaln rln; memset(&rln, 0, sizeof(rln));
            rln.qname = "r"; rln.tid = -1; rln.strand = '+'; rln.index = '1'; rln.qual = 60;
            uint32_t cigar = ((uint32_t)len << BAM_CIGAR_SHIFT) + BAM_CSOFT_CLIP;
            int refindx = 0, readindx = 0;
            rln.segments = new_readseg(read, cigar, &refindx, &readindx);
            int range[2] = { 0, range1[i] };
            cap_n = -1; cap_armed = 1;
            evidence* ev = attempt_pe_alignment(sequences, 0, position[i], range, &rln);
            cap_armed = 0;
            ns = cap_n < 0 ? 0 : cap_n;
            if (cap_n < 0 && rln.segments) free_readsegs(&rln.segments);
            while (ev) { evidence* nx = ev->next; ev->isused = TRUE; ev->next = NULL; free_used_evidence(ev); ev = nx; }
        } else {
            int sop[REFSHIM_MAXSEG], sl[REFSHIM_MAXSEG], ss[REFSHIM_MAXSEG], se[REFSHIM_MAXSEG], nev;
            ns = refshim_realign(refseq, reflength, position[i], range1[i], read, sop, sl, ss, se, REFSHIM_MAXSEG, &nev);
        }
        if (nseg_out) nseg_out[i] = ns;
        total += ns;
        ckfree(read);
    }
    return total;
}
#endif /* !REFSHIM_TRACE */

#ifdef REFSHIM_TRACE
/* ---- golden-vector recorder (linked into indelminer_trace) --------------- */
int local_align(char* seq1, const int seq1len, char* seq2, const int seq2len,
                const int indx1, const int indx2, int* const psi, int* const psj,
                int* const pei, int* const pej, int* const S);

static FILE* trace_fp(void)
{
    static FILE* fp = NULL;
    if (!fp) {
        const char* p = getenv("REFSHIM_TRACE_FILE");
        fp = fopen(p ? p : "refshim_trace.tsv", "w");
        if (!fp) { perror("trace"); exit(1); }
    }
    return fp;
}

/* one line per local_align call made by attempt_band_alignment (alignment.c:361) */
int refshim_local_align(char* seq1, const int seq1len, char* seq2, const int seq2len,
                        const int indx1, const int indx2, int* const psi, int* const psj,
                        int* const pei, int* const pej, int* const S)
{
    int score = local_align(seq1, seq1len, seq2, seq2len, indx1, indx2, psi, psj, pei, pej, S);
    FILE* fp = trace_fp();
    int ok = score > 0;
    fprintf(fp, "LA\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.*s\t%.*s\n",
            seq1len, seq2len, indx1, indx2, ok ? score : 0,
            ok ? *psi : 0, ok ? *psj : 0, ok ? *pei : 0, ok ? *pej : 0,
            seq1len, seq1, seq2len, seq2);
    return score;
}

/* one line per attempt_pe_alignment call (indelminer.c:411,486) + final segments;
 * written after the LA lines of the same read */
evidence* attempt_pe_alignment(char** const sequences, const int32_t tid,
                               const int32_t position, const int* const range,
                               readaln* const rln)
{
    FILE* fp = trace_fp();
    char* readcopy = strdup(rln->segments->sequence);
    cap_n = -1; cap_armed = 1;
    evidence* ev = refshim_inner_attempt_pe_alignment(sequences, tid, position, range, rln);
    cap_armed = 0;
    int ne = 0;
    for (evidence* e = ev; e; e = e->next) ne++;
    fprintf(fp, "PE\t%d\t%d\t%d\t%d\t%s\t%d\t%d\t", tid, position, range[0], range[1],
            readcopy, ne, cap_n < 0 ? 0 : cap_n);
    for (int i = 0; i < cap_n; i++)
        fprintf(fp, "%s%d:%d:%d:%d", i ? "," : "", cap_op[i], cap_len[i], cap_start[i], cap_end[i]);
    if (cap_n <= 0) fprintf(fp, ".");
    fprintf(fp, "\n");
    fflush(fp);
    free(readcopy);
    return ev;
}
#endif
