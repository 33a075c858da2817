/* Row f2 of SURVEY.md section 8: the in-process batched driver -- ONE pass over the BAM, no fork, the
 * caller unchanged (INDELGPU_MODE=inline).
 *
 * indelminer() hands every contig to   bam_fetch(fp, fp_index, i, beg, end, alndata, fetch_func)
 * (src/indelminer.c:797,800), and bam_fetch (samtools-0.1.19/bam_index.c:715-726) calls fetch_func once
 * per record.  The build compiles the reference's indelminer.c with -Dbam_fetch=indelgpu_bam_fetch -- a
 * compiler flag, no source change -- so that loop is the one below instead:
 *
 *   producer thread   reads the records of the region in blocks (bam_iter_read on the caller's handle),
 *                     decides from flags / CIGAR / MQ alone which of them fetch_func will hand to
 *                     attempt_pe_alignment (the tests of indelminer.c:354-368, :402, :425-473), decodes
 *                     those reads from the BAM's 4-bit form straight into the pinned SoA batch of
 *                     host/indelgpu_batch.c, oriented as :404-409 / :479-484 orient them, and realigns the
 *                     whole block with ONE indelgpu_realign_batch;
 *   main thread       calls fetch_func on the records of the previous block, in BAM order, exactly as
 *                     bam_fetch would.  When fetch_func reaches attempt_pe_alignment
 *                     (host/indelgpu_attempt.c), the answer is waiting: indelgpu_inline_lookup.
 *
 * Exactness does not rest on the producer predicting fetch_func correctly.  A prefetched answer is
 * only used when its key -- contig, position, range[1], and every base of the read -- equals the
 * arguments of the actual call; anything else (a call the producer did not foresee, a read group whose
 * range is not known yet, a full batch) is computed on the spot by the per-read path, and a prefetched
 * answer nobody asks for is dropped.  attempt_pe_alignment is a pure function of those arguments
 * (SURVEY.md 0.9), so the evidence lists, the READCHUNK flushes (indelminer.c:617) and the VCF are the
 * reference's.  range[1] comes from a hashtable private to indelminer.c; it is learnt per read group from
 * the first call that passes it (indelgpu_inline_learn), which is why the first blocks are small.
 *
 * bam_fetch calls made from inside fetch_func (find_mate_rln, indelminer.c:262-269) pass through.
 * No alignment code and no CPU fallback in this file.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bam.h"            /* bundled samtools-0.1.19 (the plain bam_fetch: this file is compiled without the -D) */
#include "errors.h"
#include "memalloc.h"

#include "indelgpu.h"
#include "indelgpu_batch.h"
#include "indelgpu_glue.h"

extern int qthreshold;                     /* -q, indelminer.c:51 */

/* ---- read group -> range[1], learnt from the calls ---------------------------------------------- */
enum { MAXRG = 512, RGNAME = 120 };
typedef struct { char name[RGNAME]; int32_t range1; } rg_entry;
static rg_entry g_rg[MAXRG];
static int g_nrg = 0;
static pthread_mutex_t g_rg_mu = PTHREAD_MUTEX_INITIALIZER;

static const char* rg_of(const bam1_t* b)
{
    const uint8_t* rg = bam_aux_get(b, "RG");              /* indelminer.c:366-370 */
    return rg != NULL ? bam_aux2Z(rg) : "generic";
}

static int32_t rg_range1(const char* name)
{
    int32_t r = -1;
    pthread_mutex_lock(&g_rg_mu);
    for (int i = 0; i < g_nrg; i++)
        if (strncmp(g_rg[i].name, name, RGNAME - 1) == 0) { r = g_rg[i].range1; break; }
    pthread_mutex_unlock(&g_rg_mu);
    return r;
}

/* ---- one block of records with its prefetched batch ---------------------------------------------- */
typedef struct {
    bam1_t* recs; int nrec, caprec;
    int32_t* cand;                 /* per record: index in `batch`, or -1 */
    igb_batch* batch;              /* one of the two pinned batches (serial & 1): only the records are kept for long */
    int answered;                  /* the batch results are valid */
    int serial;                    /* which block of the current region this is */
    int last;                      /* the region ends with this block */
    int ret;                       /* bam_iter_read's last return value */
    int32_t maxspan;               /* the longest reference span (bam_calend - pos) of any record up to and including this block */
} pf_block;

/* A ring of blocks: the consumer is at block c, the producer may fill c + 1, and blocks c - 3 .. c stay readable for
 * indelgpu_bam_fetch_cov (the per-variant coverage fetch looks back about one READCHUNK of records). */
enum { NBLK = 6, KEEP = 3 };
static pf_block g_blk[NBLK];
static igb_batch* g_batch[2];
static int g_produced = 0;                 /* blocks of the current region the producer has finished */
static int g_fetch_tid = -1, g_fetch_beg = 0, g_fetch_end = 0, g_cur_serial = -1;
static int g_last_serial = -1;             /* after the region's last block: the ring stays valid until the next region starts */
static long long g_cov_served = 0, g_cov_fallback = 0;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;

static int g_consumed = 0;                 /* blocks of the current region fetch_func has finished */

static const pf_block* g_cur_blk = NULL;   /* the record fetch_func is looking at right now */
static int g_cur_idx = -1;
static int g_in_fetch = 0;

static long long g_hits = 0, g_direct = 0, g_prefetched = 0, g_batches = 0, g_records = 0;
static double g_t_read = 0, g_t_gpu = 0, g_t_prod_wait = 0, g_t_cons_wait = 0, g_t_func = 0;     /* seconds, INDELGPU_VERBOSE */

#include <time.h>
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}
static int g_stats_registered = 0;

static void print_stats(void)
{
    fprintf(stderr, "libindelgpu: inline mode: %lld BAM records, %lld calls answered from %lld prefetched batches "
                    "(%lld reads realigned in them), %lld computed per read\n",
            g_records, g_hits, g_batches, g_prefetched, g_direct);
    fprintf(stderr, "libindelgpu: inline mode: %lld per-variant region fetches served from the retained records, %lld from the BAM\n",
            g_cov_served, g_cov_fallback);
    if (getenv("INDELGPU_VERBOSE"))
        fprintf(stderr, "libindelgpu: inline mode: prefetch thread %.2f s reading + classifying, %.2f s in indelgpu_realign_batch, %.2f s waiting for "
                        "fetch_func; main thread %.2f s in fetch_func, %.2f s waiting for the prefetch thread\n",
                g_t_read, g_t_gpu, g_t_prod_wait, g_t_func, g_t_cons_wait);
}

static int block_records(int serial)
{
    /* small first blocks: nothing can be prefetched before the first calls have taught this file the
     * reference strings and the read-group ranges */
    static int maxrec = 0;
    if (maxrec == 0) {
        const char* e = getenv("INDELGPU_INLINE_BLOCK");
        maxrec = e != NULL && atoi(e) >= 16 ? atoi(e) : 65536;
    }
    const int ramp[] = {256, 1024, 8192};
    if (serial < 3 && ramp[serial] < maxrec) return ramp[serial];
    return maxrec;
}

static const char kNt16[] = "=ACMGRSVTWYHKDBN";        /* bam.h: bam_nt16_rev_table */

/* Will fetch_func call attempt_pe_alignment for this record?  If so write the read as it will pass it
 * (ASCII, reverse-complemented or not) and return its length; 0 otherwise.  A wrong guess costs time,
 * never correctness (see the head of this file). */
static int foresee_call(const bam1_t* b, char* out, int cap)
{
    const uint32_t flag = b->core.flag;
    if (flag & (0x100 | 0x200 | 0x400 | 0x800)) return 0;                 /* indelminer.c:348-351 */
    if ((flag & 0x1) == 0) return 0;                                        /* :361 */
    const int aligned = (flag & 0x4) == 0, mate_aligned = (flag & 0x8) == 0;
    const int is_rc = (flag & 0x10) != 0, is_mate_rc = (flag & 0x20) != 0;
    if (aligned && mate_aligned && b->core.tid != b->core.mtid) return 0;  /* :364-366 */
    int revcomp;
    if (!aligned && mate_aligned) {                                         /* :384-424 */
        revcomp = !is_mate_rc;                                              /* :404 */
    } else if (aligned && mate_aligned && (flag & 0x2)) {                   /* :425-492 */
        const uint32_t* cig = bam1_cigar(b);
        const int nc = b->core.n_cigar;
        int ndel = 0, nins = 0, nclip = 0, clip3 = 0;
        for (int i = 0; i < nc; i++) {
            const int op = cig[i] & BAM_CIGAR_MASK;
            if (op == BAM_CDEL) ndel++;
            else if (op == BAM_CINS) nins++;
            else if (op == BAM_CSOFT_CLIP) {
                nclip++;
                if ((!is_rc && i == nc - 1) || (is_rc && i == 0)) clip3 = 1;   /* :439-443 */
            } else if (op != BAM_CMATCH && op != BAM_CEQUAL && op != BAM_CDIFF) return 0;   /* new_readaln stops the program */
        }
        if (ndel + nins + nclip == 0) return 0;
        if ((nclip == 0 || (nclip == 1 && clip3)) && ndel == 0 && nins == 0) return 0;     /* :455-458 */
        revcomp = (is_rc && is_mate_rc) || (!is_rc && !is_mate_rc);         /* :479 */
    } else return 0;
    const uint8_t* pmmq = bam_aux_get(b, "MQ");                             /* :389-400, :461-471 */
    const int mmq = pmmq != NULL ? (int)bam_aux2i(pmmq) : (int)b->core.qual;
    if (mmq < qthreshold) return 0;
    const int len = b->core.l_qseq;
    if (len <= 0 || len > cap) return 0;
    const uint8_t* seq = bam1_seq(b);
    for (int i = 0; i < len; i++) {
        const int code = bam1_seqi(seq, i);
        if (code != 1 && code != 2 && code != 4 && code != 8 && code != 15) return 0;   /* bit2char stops the program (readaln.c:4-17) */
        out[i] = kNt16[code];
    }
    if (revcomp) {                                                          /* sequences.c:204-220 on A C G T N */
        for (int i = 0, j = len - 1; i <= j; i++, j--) {
            const char a = out[i], c = out[j];
            const char ca = a == 'A' ? 'T' : a == 'C' ? 'G' : a == 'G' ? 'C' : a == 'T' ? 'A' : 'N';
            const char cc = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
            out[i] = cc; out[j] = ca;
        }
    }
    return len;
}

typedef struct {
    bamFile fp; bam_iter_t iter; int tid;
} producer_arg;

/* fill one block: records, candidates, one GPU batch */
static void fill_block(pf_block* k, const producer_arg* pa, int serial)
{
    const int want = block_records(serial);
    if (k->caprec < want) {
        k->recs = ckrealloc(k->recs, sizeof(bam1_t) * (size_t)want);
        memset(k->recs + k->caprec, 0, sizeof(bam1_t) * (size_t)(want - k->caprec));
        k->cand = ckrealloc(k->cand, sizeof(int32_t) * (size_t)want);
        k->caprec = want;
    }
    if (g_batch[serial & 1] == NULL) {
        /* candidates are a small share of the records; a batch that fills up just leaves the rest of the
         * block to the per-read path */
        const int cap = block_records(1 << 20) / 4 + 256;
        g_batch[serial & 1] = igb_create(cap, (int64_t)cap * 256);
        if (g_batch[serial & 1] == NULL) fatalf("libindelgpu: inline mode: cannot allocate the pinned batch (%s)", indelgpu_last_error());
    }
    k->batch = g_batch[serial & 1];
    igb_clear(k->batch);
    k->nrec = 0; k->answered = 0; k->last = 0; k->ret = 0; k->serial = serial;
    int32_t maxspan = serial > 0 ? g_blk[(serial - 1) % NBLK].maxspan : 1;
    int32_t dtid = 0;
    pthread_mutex_lock(&indelgpu_glue_gpu_mu);
    indelgpu_ctx* ctx = indelgpu_glue_ctx_peek(pa->tid, &dtid);
    pthread_mutex_unlock(&indelgpu_glue_gpu_mu);
    char read[1024];
    const double t_r0 = now_s();
    while (k->nrec < want) {
        bam1_t* b = &k->recs[k->nrec];
        const int ret = bam_iter_read(pa->fp, pa->iter, b);
        if (ret < 0) { k->last = 1; k->ret = ret; break; }
        if (b->core.n_cigar) { const int32_t span = (int32_t)bam_calend(&b->core, bam1_cigar(b)) - b->core.pos; if (span > maxspan) maxspan = span; }
        int32_t slot = -1;
        if (ctx != NULL && b->core.mtid == pa->tid) {
            const int len = foresee_call(b, read, (int)sizeof(read));
            if (len > 0) {
                const int32_t range1 = rg_range1(rg_of(b));
                if (range1 >= 0 && igb_push(k->batch, read, len, dtid, b->core.mpos, range1) == 0) slot = k->batch->n - 1;
            }
        }
        k->cand[k->nrec++] = slot;
    }
    k->maxspan = maxspan;
    const double t_r1 = now_s();
    g_t_read += t_r1 - t_r0;
    if (k->batch->n > 0) {
        pthread_mutex_lock(&indelgpu_glue_gpu_mu);
        const int rc = igb_run(k->batch, ctx);
        pthread_mutex_unlock(&indelgpu_glue_gpu_mu);
        g_t_gpu += now_s() - t_r1;
        /* a batch that fails (e.g. it holds a read on which the reference itself would abort, status 7) is
         * dropped: its reads go through the per-read path, which reports the error for the call that is
         * actually made -- a foreseen call that never happens must not stop the run */
        k->answered = rc == 0;
        if (rc == 0) { g_batches++; g_prefetched += k->batch->n; }
    }
}

static void* producer_main(void* arg)
{
    const producer_arg* pa = arg;
    for (int serial = 0;; serial++) {
        pf_block* k = &g_blk[serial % NBLK];
        const double t_w0 = now_s();
        pthread_mutex_lock(&g_mu);
        while (serial > g_consumed + 1) pthread_cond_wait(&g_cv, &g_mu);      /* at most one block ahead of fetch_func */
        /* Until a call has taught this file the contig's context and a read group's range nothing can be
         * prefetched: do not run ahead of fetch_func then, or the blocks filled in the meantime are all misses */
        while (g_consumed < serial) {
            int32_t dtid;
            pthread_mutex_unlock(&g_mu);
            pthread_mutex_lock(&indelgpu_glue_gpu_mu);
            const int ready = indelgpu_glue_ctx_peek(pa->tid, &dtid) != NULL;
            pthread_mutex_unlock(&indelgpu_glue_gpu_mu);
            pthread_mutex_lock(&g_rg_mu);
            const int known = g_nrg > 0;
            pthread_mutex_unlock(&g_rg_mu);
            pthread_mutex_lock(&g_mu);
            if (ready && known) break;
            if (g_consumed < serial) pthread_cond_wait(&g_cv, &g_mu);
        }
        pthread_mutex_unlock(&g_mu);
        g_t_prod_wait += now_s() - t_w0;
        fill_block(k, pa, serial);
        const int last = k->last;
        pthread_mutex_lock(&g_mu);
        g_produced = serial + 1;
        pthread_cond_broadcast(&g_cv);
        pthread_mutex_unlock(&g_mu);
        if (last) break;
    }
    return NULL;
}

int indelgpu_bam_fetch(bamFile fp, const bam_index_t* idx, int tid, int beg, int end, void* data, bam_fetch_f func)
{
    /* nested call, or another mode: samtools' own loop.  The mode is read from the environment here, NOT through
     * indelgpu_glue_mode(): that call decides the mode for good and, for INDELGPU_MODE=auto, forks -- which must
     * happen inside the first attempt_pe_alignment call, where the replaying parent then blocks on the pipe
     * until the recording child has given itself private descriptions of the open BAM */
    if (g_in_fetch || !indelgpu_glue_env_is_inline())
        return bam_fetch(fp, idx, tid, beg, end, data, func);
    if (indelgpu_glue_mode() != MODE_INLINE) fatalf("libindelgpu: INDELGPU_MODE changed while running");
    if (!g_stats_registered) { atexit(print_stats); g_stats_registered = 1; }
    g_in_fetch = 1;
    producer_arg pa;
    pa.fp = fp; pa.tid = tid;
    pa.iter = bam_iter_query(idx, tid, beg, end);
    g_consumed = 0; g_produced = 0; g_last_serial = -1;
    g_fetch_tid = tid; g_fetch_beg = beg; g_fetch_end = end;
    pthread_t thr;
    if (pthread_create(&thr, NULL, producer_main, &pa) != 0) fatalf("libindelgpu: inline mode: cannot start the prefetching thread");
    int ret = 0;
    for (int serial = 0;; serial++) {
        pf_block* k = &g_blk[serial % NBLK];
        const double t_c0 = now_s();
        pthread_mutex_lock(&g_mu);
        while (g_produced <= serial) pthread_cond_wait(&g_cv, &g_mu);
        pthread_mutex_unlock(&g_mu);
        const double t_c1 = now_s();
        g_t_cons_wait += t_c1 - t_c0;
        g_cur_blk = k; g_cur_serial = serial;
        for (int i = 0; i < k->nrec; i++) {
            g_cur_idx = i;
            func(&k->recs[i], data);                                        /* bam_index.c:722 */
        }
        g_records += k->nrec;
        g_t_func += now_s() - t_c1;
        g_cur_blk = NULL; g_cur_idx = -1; g_cur_serial = -1;
        const int last = k->last;
        if (last) g_last_serial = serial;
        ret = k->ret;
        pthread_mutex_lock(&g_mu);
        g_consumed = serial + 1;
        pthread_cond_broadcast(&g_cv);
        pthread_mutex_unlock(&g_mu);
        if (last) break;
    }
    pthread_join(thr, NULL);
    bam_iter_destroy(pa.iter);
    g_in_fetch = 0;
    return ret == -1 ? 0 : ret;                                            /* bam_index.c:725 */
}

/* Row f3, second half.  calculate_cov_params (shared.c:178-212, compiled with -Dbam_fetch=indelgpu_bam_fetch_cov) fetches
 * the region of every printed variant: samtools starts at the linear-index offset of the region's 16 kb window and parses
 * ~1 500 records to find the ~50 that overlap -- 3.7 times the whole file over a 30x run, all on the main thread.  Those
 * records went through this file a moment ago: the consumer's block and the three before it are still in memory, in file
 * order.  A fetch whose region they cover completely is answered from them with bam_iter_read's own test
 * (bam_index.c:642-655, :695-701: same contig, pos < end, end of the alignment > beg) in the same order; anything else
 * goes to samtools. */
int indelgpu_bam_fetch_cov(bamFile fp, const bam_index_t* idx, int tid, int beg, int end, void* data, bam_fetch_f func)
{
    if (beg < 0) beg = 0;                                                   /* bam_iter_query, bam_index.c:604 */
    const int c = g_cur_serial >= 0 ? g_cur_serial : g_last_serial;        /* the variants left at the end of a contig are printed after its fetch */
    if (c < 0 || tid != g_fetch_tid || end <= beg || beg < g_fetch_beg || end > g_fetch_end ||
        getenv("INDELGPU_NO_BAM_CACHE") != NULL) { g_cov_fallback++; return bam_fetch(fp, idx, tid, beg, end, data, func); }
    const int lo = c - KEEP > 0 ? c - KEEP : 0;
    const pf_block* kc = &g_blk[c % NBLK];
    const pf_block* k0 = &g_blk[lo % NBLK];
    /* nothing that overlaps the region may lie before the first kept record or after the last one */
    const int before_ok = lo == 0 || (k0->nrec > 0 && (int64_t)k0->recs[0].core.pos + kc->maxspan <= beg);
    const int after_ok = kc->last || (kc->nrec > 0 && kc->recs[kc->nrec - 1].core.pos >= end);
    if (!before_ok || !after_ok) { g_cov_fallback++; return bam_fetch(fp, idx, tid, beg, end, data, func); }
    const int64_t from = (int64_t)beg - kc->maxspan;                        /* records that start before this cannot reach beg */
    for (int sidx = lo; sidx <= c; sidx++) {
        const pf_block* k = &g_blk[sidx % NBLK];
        if (k->nrec == 0 || k->recs[k->nrec - 1].core.pos < from) continue;
        int a = 0, z = k->nrec;                                             /* first record with pos >= from */
        while (a < z) { const int m = (a + z) >> 1; if (k->recs[m].core.pos < from) a = m + 1; else z = m; }
        for (int i = a; i < k->nrec; i++) {
            const bam1_t* b = &k->recs[i];
            if (b->core.pos >= end) { g_cov_served++; return 0; }
            const uint32_t rend = b->core.n_cigar ? bam_calend(&b->core, bam1_cigar(b)) : (uint32_t)b->core.pos + 1;
            if (rend > (uint32_t)beg) func(b, data);
        }
    }
    g_cov_served++;
    return 0;
}

int indelgpu_inline_lookup(int32_t tid, int32_t position, int32_t range1, const char* read, int32_t readlen,
                           int32_t* nseg, int32_t* rstart, const uint32_t** words)
{
    const pf_block* k = g_cur_blk;
    if (k == NULL || g_cur_idx < 0 || !k->answered) return 0;
    const int32_t s = k->cand[g_cur_idx];
    if (s < 0) return 0;
    const igb_batch* b = k->batch;
    const bam1_t* rec = &k->recs[g_cur_idx];
    /* the key: every argument the result depends on */
    if (rec->core.mtid != tid || b->position[s] != position || b->range1[s] != range1 ||
        b->off[s + 1] - b->off[s] != (int64_t)readlen || memcmp(b->bases + b->off[s], read, (size_t)readlen) != 0) return 0;
    *words = igb_segments(b, s, nseg, rstart);
    g_hits++;
    return 1;
}

void indelgpu_inline_learn(int32_t range1)
{
    g_direct++;
    if (g_cur_blk == NULL || g_cur_idx < 0) return;
    const char* name = rg_of(&g_cur_blk->recs[g_cur_idx]);
    pthread_mutex_lock(&g_rg_mu);
    int i;
    for (i = 0; i < g_nrg; i++)
        if (strncmp(g_rg[i].name, name, RGNAME - 1) == 0) break;
    if (i == g_nrg && g_nrg < MAXRG) { strncpy(g_rg[i].name, name, RGNAME - 1); g_rg[i].name[RGNAME - 1] = '\0'; g_nrg++; }
    if (i < g_nrg) g_rg[i].range1 = range1;
    pthread_mutex_unlock(&g_rg_mu);
}
