/* libindelgpu.so -- B200 (sm_100a) split-read realignment for indelMINER.
 *
 * Drop-in boundary for the hot path of ratan-lab/indelMINER:
 *   attempt_pe_alignment        src/alignment.h:21-25   (src/alignment.c:764-799)
 *     find_best_band            src/alignment.c:393-447 (static)
 *     attempt_band_alignment    src/alignment.c:343-391 (static)
 *       local_align             src/localalign.h:15-25
 *         ALIGN                 src/globalalign.h:19-28
 *       fetch_cigar             src/globalalign.h:39-48
 *     update_readsegs           src/readaln.h:56-64     (segment list; materialised by the host)
 *
 * Plain C ABI: pointers and sizes only, no CUDA or torch types.  `stream` arguments are a
 * cudaStream_t passed as void* (NULL = the context's own stream).  Buffers named h_* live in
 * host memory (pinned memory from indelgpu_host_alloc makes the copies asynchronous), d_* in
 * device memory of the context's GPU.
 *
 * Error convention.  The batched API returns 0 on success and a negative INDELGPU_E* code on
 * failure, with text from indelgpu_last_error().  The reference-prototype entry points at the
 * bottom keep the reference's own convention (message on stderr + exit(EXIT_FAILURE),
 * src/asserts.h:12-19, src/errors.c:15-27): they have no way to return an error.
 * There is NO CPU fallback anywhere in this library: without a usable GPU every call fails.
 */
#ifndef INDELGPU_H
#define INDELGPU_H

#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define INDELGPU_VERSION 100

enum {
    INDELGPU_OK = 0,
    INDELGPU_ECUDA = -1,      /* a CUDA runtime call failed                         */
    INDELGPU_EINVAL = -2,     /* bad argument (NULL, negative size, unknown contig) */
    INDELGPU_ELIMIT = -3,     /* a size exceeds what the kernels are built for      */
    INDELGPU_ENOMEM = -4      /* host or device allocation failed                   */
};

/* The globals src/alignment.c reads (alignment.c:3-9; set by getopt in indelminer.c:948-1025)
 * plus the compile-time scoring constants of src/localalign.c:10-13. */
typedef struct indelgpu_params {
    int32_t klength;      /* -k  default 6, 2..15          (indelminer.c:933,1028) */
    int32_t numgaps;      /* -g  default 0                 (indelminer.c:934)      */
    int32_t maxdelsize;   /* -s  default 1000              (indelminer.c:930)      */
    int32_t ethreshold;   /* -n  default 10                (indelminer.c:940)      */
    int32_t match;        /* +1                            (localalign.c:12)       */
    int32_t mismatch;     /* -10                           (localalign.c:13)       */
    int32_t gapopen;      /* 10                            (localalign.c:10)       */
    int32_t gapextend;    /* 10                            (localalign.c:11)       */
} indelgpu_params;

typedef struct indelgpu_ctx indelgpu_ctx;

int  indelgpu_version(void);
void indelgpu_default_params(indelgpu_params* p);
const char* indelgpu_last_error(void);

/* One context per host worker thread / GPU (the reference's ALIGN keeps file-scope state,
 * globalalign.c:19-37; this library keeps none). */
/* number of CUDA devices this process can see (0 when there is none: every other entry point then fails) */
int           indelgpu_device_count(void);
indelgpu_ctx* indelgpu_create(int device, const indelgpu_params* p);
void          indelgpu_destroy(indelgpu_ctx* ctx);
int           indelgpu_device(const indelgpu_ctx* ctx);
int           indelgpu_sm_count(const indelgpu_ctx* ctx);

/* Pinned host memory for batch staging. */
void* indelgpu_host_alloc(size_t bytes);
void  indelgpu_host_free(void* p);

/* Upload the reference contigs: replaces `char** sequences` of attempt_pe_alignment
 * (built by read_reference, src/shared.c:46-82: one upper-cased string per contig).
 * Kept resident in HBM as raw bytes (DP compares raw bytes, localalign.c:61-67) plus a
 * 2-bit packed copy with every non-ACGT byte coded as A (base2bits, alignment.c:11-24). */
int indelgpu_set_reference(indelgpu_ctx* ctx, int32_t ncontigs,
                           const char* const* sequences, const int64_t* lengths);

/* ---- batched attempt_pe_alignment ------------------------------------------------------
 * Read i is h_read_bases[h_read_off[i] .. h_read_off[i+1]) (ASCII, already reverse-
 * complemented by the caller as indelminer.c:404-409,479-484 does), realigned near
 * (tid[i], position[i]) with range1[i] = range[1] of its read group (alignment.c:775,780).
 *
 * Per read the result is what update_readsegs (readaln.c:348-458) would have built:
 *   status[i]   INDELGPU_ST_* (which exit of attempt_diagonal_alignments was taken)
 *   nseg[i]     number of segments; 0 <=> the reference returns NULL
 *   rstart[i]   refindx the first segment starts at (r1 or r3)
 *   seg_off[i]  index of the read's first segment word in segs[]
 *   segs[]      BAM-style words (len << 4 | op), ops = 7, X 8, I 1, D 2, S 4; walking them with
 *               new_readseg's bookkeeping (readaln.c:24-99) from rstart gives every start/end.
 * Evidence (alignment.c:449-476) = one record per I / D segment, built by the host.
 */
enum {
    INDELGPU_ST_UNALIGNED = 0,   /* q1 == q2                               alignment.c:568        */
    INDELGPU_ST_WHOLE     = 1,   /* whole read aligned in round 1          alignment.c:575        */
    INDELGPU_ST_SHORT     = 2,   /* ethreshold guard                       :608,631,665,687       */
    INDELGPU_ST_R2FAIL    = 3,   /* round 2 did not reach the read end     :623,645,679,701       */
    INDELGPU_ST_NOBRANCH  = 4,   /* no round-2 branch applies              :651,657,707,712       */
    INDELGPU_ST_NOCOMBINE = 5,   /* segments neither overlap nor abut      :750                   */
    INDELGPU_ST_SPLIT     = 6,   /* two segments stitched                  :724-749               */
    INDELGPU_ST_ASSERT    = 7    /* input on which the reference itself aborts (forceassert /
                                    exit): bad contig, window asserts :548-553, numdiagonals <=
                                    numgaps :405; nseg = 0 and the batch call returns ELIMIT     */
};

/* Optional per-read intermediates (kernel-level parity tests, debugging). */
typedef struct indelgpu_detail {
    int32_t low1, up1, r1, r2, q1, q2, n1, score1;   /* round 1: find_best_band + attempt_band_alignment */
    int32_t low2, up2, r3, r4, q3, q4, n2, score2;   /* round 2 (n2 counts the added soft clip)          */
    int32_t index;                                    /* junction passed to update_readsegs               */
    int32_t cells_fwd, cells_rev, cells_glob;         /* DP cells swept (GCUPS numerator)      */
} indelgpu_detail;

typedef struct indelgpu_batch {
    int32_t        n;             /* reads in the batch                      */
    const uint8_t* read_bases;    /* concatenated read bytes                 */
    const int64_t* read_off;      /* n + 1 offsets into read_bases           */
    const int32_t* tid;           /* contig index of the mate                */
    const int32_t* position;      /* 0-based mate position (core.mpos)       */
    const int32_t* range1;        /* max proper insert of the read group     */
} indelgpu_batch;

typedef struct indelgpu_result {
    int32_t*  status;             /* n                                       */
    int32_t*  nseg;               /* n                                       */
    int32_t*  rstart;             /* n                                       */
    int64_t*  seg_off;            /* n                                       */
    uint32_t* segs;               /* seg_capacity words                      */
    int64_t   seg_capacity;       /* in: words available in segs             */
    int64_t   seg_count;          /* out: words used                         */
    indelgpu_detail* detail;      /* n, or NULL                              */
    uint32_t* cigar1;             /* n * cigar_stride words, or NULL (round-1 CIGAR, debug) */
    uint32_t* cigar2;             /* n * cigar_stride words, or NULL (round-2 CIGAR, debug) */
    int32_t   cigar_stride;
} indelgpu_result;

/* worst-case number of segment words a batch can produce */
int64_t indelgpu_seg_bound(int32_t n, int64_t total_read_bases);

/* Host buffers in, host buffers out: H2D copies, kernels, D2H copies, synchronised on return.
 * Batches of >= 2^18 reads are processed in chunks of 2^18 reads (the first three are 1/8, 1/4 and 1/2 of
 * that, so that the first kernel starts after a short copy, and the last three shrink the same way so that
 * little is left to copy back after the last kernel) whose copies overlap the kernels of
 * the neighbouring chunks (pinned buffers from indelgpu_host_alloc make the copies asynchronous);
 * results do not depend on the chunking.  Contexts are independent and may be driven from different
 * host threads, one thread per context at a time. */
int indelgpu_realign_batch(indelgpu_ctx* ctx, const indelgpu_batch* h_in, indelgpu_result* h_out);

/* The same with the reads as the BAM holds them (bam1_seq: 4 bits per base, high nibble first, every read starting
 * on a byte): the host copies the record's bytes instead of expanding them with bit2char (readaln.c:4-17) and
 * reversing them (indelminer.c:404-409, 479-484) -- both happen on the device -- and the batch is 0.6 bytes per
 * base on the wire instead of 1.  Codes other than 1, 2, 4, 8, 15 (A C G T N) are inputs the reference stops on:
 * the call returns INDELGPU_ELIMIT.  Debug outputs (detail, cigar1, cigar2) are not available here. */
typedef struct indelgpu_batch4 {
    int32_t        n;
    const uint8_t* seq4;          /* concatenated 4-bit reads                               */
    const int64_t* byte_off;      /* n + 1 byte offsets into seq4; read i fills (len[i] + 1) / 2 bytes */
    const int32_t* len;           /* n   read lengths in bases (core.l_qseq)                */
    const uint8_t* flags;         /* n   bit 0: reverse-complement the read before aligning */
    const int32_t* tid;
    const int32_t* position;
    const int32_t* range1;
} indelgpu_batch4;
int indelgpu_realign_batch4(indelgpu_ctx* ctx, const indelgpu_batch4* h_in, indelgpu_result* h_out);

/* Device buffers in and out (all pointers in *d_in / *d_out are device pointers, the structs
 * themselves live on the host).  d_in->read_bases must be 16-byte aligned and readable up to the next
 * 16-byte boundary past its last base (the kernels stage reads with TMA bulk copies; any cudaMalloc'ed
 * buffer qualifies).  d_out->seg_count is not filled; the count is left in the
 * int64 device word *d_seg_count.  Asynchronous on `stream`. */
int indelgpu_realign_batch_device(indelgpu_ctx* ctx, const indelgpu_batch* d_in,
                                  int32_t max_read_len, int32_t max_range1,
                                  indelgpu_result* d_out, int64_t* d_seg_count, void* stream);

/* Counters of the last batch call; synchronises the context's stream.
 *   out[0..2]  DP cells swept: forward, reverse, ALIGN (the GCUPS numerator)
 *   out[3]     algorithmic bytes: sum over alignments of N + M + 4 * (6 + ncigar) */
int indelgpu_last_counters(indelgpu_ctx* ctx, int64_t out[4]);

/* ALIGN cells of the last indelgpu_band_align_batch that the cell counters include (they count what the
 * reference sweeps) but that were not swept on the GPU: the unique-diagonal shortcut.  Executed cells =
 * forward + reverse + align - *out. */
int indelgpu_last_shortcut_cells(indelgpu_ctx* ctx, int64_t* out);

/* Error flag the kernels of the last batch left on the device (waits for them): 0 none; 1 at least one
 * read was rejected (status INDELGPU_ST_ASSERT); 2 the segment buffer overflowed -- nseg / seg_off are
 * set but the words were not written, compare *d_seg_count with seg_capacity; 3 a TMA bulk copy never
 * completed and the outputs of the affected reads are undefined.  indelgpu_realign_batch checks this
 * itself and returns an error; callers of indelgpu_realign_batch_device must ask.  < 0: INDELGPU_E*. */
int indelgpu_last_error_flag(indelgpu_ctx* ctx);

/* device time (CUDA events on the context's stream) of the kernel the last
 * indelgpu_band_align_batch call launched, in ms; < 0 when nothing was timed */
double indelgpu_last_kernel_ms(indelgpu_ctx* ctx);

/* measured INT32 issue rate of this GPU in Gop/s (independent add + max chains): the denominator
 * of the banded-DP roofline (it is not in MEASURED_PEAKS.json) */
int indelgpu_int32_peak(indelgpu_ctx* ctx, double* gops);

/* number of kernels the last indelgpu_realign_batch[_device] call launched */
int indelgpu_last_launch_count(const indelgpu_ctx* ctx);

/* ---- batched kernel-level entry points (band sweeps, parity tests) ---------------------
 * n independent (read, window) tasks: read j = h_reads[read_off[j]..read_off[j+1]),
 * window j = h_refs[ref_off[j]..ref_off[j+1]).  Semantics of one task = find_best_band /
 * local_align (+ALIGN) + fetch_cigar of the reference with absolute offsets zstart = 0. */
int indelgpu_find_best_band_batch(indelgpu_ctx* ctx, int32_t n,
                                  const uint8_t* h_reads, const int64_t* h_read_off,
                                  const uint8_t* h_refs, const int64_t* h_ref_off,
                                  const int32_t* h_anchor_rel,      /* anchor - zstart1 */
                                  int32_t* h_low, int32_t* h_up);

int indelgpu_band_align_batch(indelgpu_ctx* ctx, int32_t n,
                              const uint8_t* h_reads, const int64_t* h_read_off,
                              const uint8_t* h_refs, const int64_t* h_ref_off,
                              const int32_t* h_low, const int32_t* h_up,
                              int32_t* h_score,            /* n; 0 = no alignment                 */
                              int32_t* h_ends,             /* 4n: si, sj, ei, ej (1-based incl.)  */
                              int32_t* h_ncigar,           /* n                                   */
                              uint32_t* h_cigar,           /* n * cigar_stride                    */
                              int32_t cigar_stride,
                              int32_t* h_script,           /* n * script_stride ints, or NULL     */
                              int32_t script_stride,
                              int64_t* h_cells);           /* 3: fwd, rev, glob totals, or NULL   */

/* ---- known-indel support check ------------------------------
 * Batched realign_with_indel (src/variant.c:1246-1424), the DP behind is_indel_supported
 * (variant.h:91, annotate mode): task j aligns query j = h_queries[query_off[j]..query_off[j+1])
 * -- the read slice [qstart, qstop) of variant.c:1278-1283 -- against target j, the reference
 * interval with the variant spliced in that variant.c:1259-1275 builds (the host keeps building it
 * with the reference's own string code).  Outputs per task are the three counters check_for_indel
 * compares (variant.c:1549-1553): substitutions, gap columns, aligned columns (+1, as the reference
 * counts).  Lengths up to 8000 (scores are kept in 16 bits). */
int indelgpu_indel_support_batch(indelgpu_ctx* ctx, int32_t n,
                                 const uint8_t* h_targets, const int64_t* h_target_off,
                                 const uint8_t* h_queries, const int64_t* h_query_off,
                                 int32_t* h_subs, int32_t* h_indels, int32_t* h_aligned,
                                 int64_t* h_cells /* 1: sum of len1 * len2, or NULL */);

/* ---- reference-prototype entry points (1-element batches on the default context) --------
 * Same names, arguments and results as the reference so the rest of the C caller links
 * unchanged.  The default context is created on first use on device $INDELGPU_DEVICE (0). */

/* src/localalign.h:15-25 */
int local_align(char* seq1, const int seq1len, char* seq2, const int seq2len,
                const int indx1, const int indx2,
                int* const psi, int* const psj, int* const pei, int* const pej, int* const S);

/* src/globalalign.h:19-28; A and B point one element BEFORE the first symbol */
int ALIGN(char* A, char* B, int M, int N, int low, int up, int W[][128], int G, int H, int* S);

/* src/globalalign.h:30-37 (debug pretty-printer; formats on the host) */
int DISPLAY(FILE* F, char* A, char* B, int M, int N, int* S, int AP, int BP);

/* src/globalalign.h:39-48; *pcigar is caller-allocated (>= 1 word) and may be realloc'ed */
int fetch_cigar(char* A, char* B, int M, int N, int* S, int AP, int BP,
                const int readlength, int* const pnumops, uint32_t** pcigar);

#ifdef __cplusplus
}
#endif
#endif
