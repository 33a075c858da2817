/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for indelMINER's split-read realignment hot path: a plain-C
 * restatement of the reference algorithm, used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg as the CHECKER for the CUDA path.  Nothing in
 * the product (indelminer_b200/, libindelgpu.so) may include, link or call it.
 *
 * Parity pinning: every function here is differentially tested against the
 * reference's own object code (oracle/_ref/libref_dp.so, libref_align.so, built
 * by oracle/Makefile from the sources under /root/reference) and against the
 * golden vectors traced from the reference run on its test_data
 * (tests/golden/testdata_trace.tsv.gz, 1255 local_align + 697 attempt_pe_alignment
 * calls).  See tests/test_oracle_vs_reference.py and tests/test_golden_trace.py.
 */
#ifndef INDEL_ORACLE_H
#define INDEL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the globals alignment.c reads (alignment.c:3-9) + localalign.c:10-13 scoring */
typedef struct {
    int klength;      /* -k, default 6   (indelminer.c:933)  */
    int numgaps;      /* -g, default 0   (indelminer.c:934)  */
    int maxdelsize;   /* -s, default 1000 (indelminer.c:930) */
    int ethreshold;   /* -n, default 10  (indelminer.c:940)  */
    int match;        /* +1   localalign.c:12 */
    int mismatch;     /* -10  localalign.c:13 */
    int gapopen;      /* 10   localalign.c:10 */
    int gapextend;    /* 10   localalign.c:11 */
} orc_params;

void orc_default_params(orc_params* p);

/* DP cell counters: the denominators of GCUPS */
typedef struct {
    long long fwd;    /* cells of the forward sweep   (localalign.c:100-131) */
    long long rev;    /* cells of the reverse sweep until the first hit (localalign.c:144-176) */
    long long glob;   /* cells swept by every align() invocation incl. recursion (globalalign.c:147-234) */
} orc_cells;

/* alignment.c:393-447 (with read_seeds :29-68, bin_diagonals :70-128,
 * bin_bands :130-140, select_band :142-181) */
void orc_find_best_band(const orc_params* p,
                        const char* refseq, uint32_t zstart1, uint32_t end1, uint32_t anchor,
                        const char* readseq, uint32_t zstart2, uint32_t end2,
                        int* plow, int* pup);

/* localalign.c:15-196.  seq1 = read (length M), seq2 = reference window (length N),
 * 0-based pointers to the first symbol.  Returns the score (0 = no alignment),
 * 1-based inclusive end points and the edit script S (0 = REP, +k = k ref-only
 * symbols, -k = k read-only symbols); *nS = number of script entries. */
int orc_local_align(const orc_params* p, const char* seq1, int M, const char* seq2, int N,
                    int low, int up, int* psi, int* psj, int* pei, int* pej,
                    int* S, int* nS, orc_cells* cells);

/* globalalign.c:333-401 (ALIGN) incl. the divide-and-conquer align() :66-307.
 * A, B are 0-based pointers to the first symbol. */
int orc_global_align(const orc_params* p, const char* A, const char* B, int M, int N,
                     int low, int up, int* S, int* nS, orc_cells* cells);

/* globalalign.c:507-604.  A, B 0-based pointers to the first ALIGNED symbol;
 * AP = 1-based read position of A[0]; returns the mismatch count. */
int orc_fetch_cigar(const char* A, const char* B, int M, int N, const int* S,
                    int AP, int readlength, uint32_t* cigar, int* pnumops);

/* alignment.c:343-391.  Returns numcigarops (0 when score <= 0; then r1=r2=q1=q2=0). */
int orc_attempt_band_alignment(const orc_params* p,
                               const char* refseq, uint32_t zstart1, uint32_t end1,
                               const char* readseq, uint32_t zstart2, uint32_t end2,
                               int low, int up, int* pr1, int* pr2, int* pq1, int* pq2,
                               uint32_t* cigar, int* pscore, orc_cells* cells);

#define ORC_MAXSEG 1024

/* result of the whole two-round realignment of one read:
 * attempt_pe_alignment (alignment.c:764-799) -> attempt_diagonal_alignments (:539-759)
 * -> update_readsegs (readaln.c:348-458) */
typedef struct {
    int status;               /* ORC_ST_*: which exit of attempt_diagonal_alignments was taken */
    int nseg;                 /* 0 = the reference returns NULL */
    int nevidence;            /* number of D / I segments (alignment.c:449-476) */
    int seg_op[ORC_MAXSEG];
    int seg_len[ORC_MAXSEG];
    int seg_start[ORC_MAXSEG];
    int seg_end[ORC_MAXSEG];
    /* intermediates, for debugging and kernel-level parity */
    int low1, up1, r1, r2, q1, q2, n1, score1;
    int low2, up2, r3, r4, q3, q4, n2, score2;
    int index;
    uint32_t cigar1[ORC_MAXSEG];
    uint32_t cigar2[ORC_MAXSEG];
} orc_result;

enum {
    ORC_ST_UNALIGNED = 0,     /* q1 == q2                         alignment.c:568 */
    ORC_ST_WHOLE = 1,         /* whole read aligned in round 1    alignment.c:575 */
    ORC_ST_SHORT = 2,         /* ethreshold guard                 alignment.c:608,631,665,687 */
    ORC_ST_R2FAIL = 3,        /* round 2 did not reach the read end  :623,645,679,701 */
    ORC_ST_NOBRANCH = 4,      /* neither q1==0 nor q2==L / r2>=anchor / r1==anchor  :651,657,707,712 */
    ORC_ST_NOCOMBINE = 5,     /* segments neither overlap nor abut :750 */
    ORC_ST_SPLIT = 6,         /* two segments combined            :724-749 */
    ORC_ST_ASSERT = 7         /* the reference stops the program: forceassert(numdiagonals > numgaps), alignment.c:405 */
};

int orc_would_assert(const orc_params* p, uint32_t N, uint32_t M);

void orc_realign_read(const orc_params* p, const char* refseq, int reflength,
                      int position, int range1, const char* read, int readlength,
                      orc_result* out, orc_cells* cells);

long orc_realign_batch(const orc_params* p, const char* refseq, int reflength, int n,
                       const char* reads, const long long* off, const int* position,
                       const int* range1, int* nseg_out);

/* ---- row f1 of SURVEY.md section 8: realign_with_indel (variant.c:1246-1424) ------------------
 * "Would this read support the known indel?"  The reference splices the variant into a copy of the
 * reference interval [rstart, rstop) (the target), aligns query[qstart, qstop) against it with a
 * full-matrix affine DP (match +2, mismatch -1, gap open 4, gap extend 1; E restarts at 0 on every
 * row and F starts at 0: both kept), traces back from the first maximum while the score stays
 * positive, and counts substitutions, gap columns and aligned columns (one more than there are:
 * the counting loop starts ON the terminating NUL, variant.c:1405-1417; kept).
 * vtype: 0 = INSERTION, 1 = DELETION (varianttype, evidence.h:14-18). */

/* the target the reference builds (variant.c:1259-1275); returns a malloc'ed NUL-terminated string */
char* orc_indel_target(const char* reference, int rstart, int rstop, int vtype, int vstart, int vstop,
                       const char* alternate);

/* the DP + traceback + counting on an already built target (variant.c:1277-1423);
 * *cells (optional) += len1 * len2 */
void orc_indel_support_dp(const char* target, int len1, const char* query, int len2,
                          int* subs, int* indels, int* aligned, long long* cells);

/* both, with the reference's own argument list */
void orc_realign_with_indel(const char* reference, int rstart, int rstop, const char* query, int qstart,
                            int qstop, int vtype, int vstart, int vstop, const char* alternate,
                            int* subs, int* indels, int* aligned);

#ifdef __cplusplus
}
#endif
#endif
