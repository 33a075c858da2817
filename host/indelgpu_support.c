/* Host-side glue for annotate mode, in C like the reference's host code: a replacement for the
 * file-local function
 *
 *     static void realign_with_indel(const char* const reference, const int rstart, const int rstop,
 *                                    const readaln* const rln, const int qstart, const int qstop,
 *                                    const knownvariant* const variant,
 *                                    int* const alnsubs, int* const alnindels, int* const alnaligned);
 *                                                                            (src/variant.c:1246-1424)
 *
 * with the same arguments and results, computed on the GPU through indelgpu_indel_support_batch
 * (include/indelgpu.h).  The function is `static` in the reference, so binding it is a two-line source
 * change in variant.c: the definition at :1246 goes (or is renamed) and this
 * file's prototype is declared instead; check_for_indel (:1427-1558), is_indel_supported and the rest
 * of variant.c stay as they are.  oracle/Makefile target `gpuprog_annotate` builds exactly that from the
 * reference sources where they lie.
 *
 * What stays on the host is what the reference does before its DP: the target string -- the
 * reference interval with the variant spliced in (variant.c:1259-1275) -- and the query slice
 * (:1278-1283).  There is no alignment code and no CPU fallback in this file.
 *
 * Batched operation with the caller unchanged, as in host/indelgpu_attempt.c: realign_with_indel is a
 * pure function of its arguments, and its result only decides whether LATER reads of the same variant
 * are looked at (check_for_indel returns early once a read supports the variant, :1444-1448).  So
 *   INDELGPU_MODE=record  answers every call "does not support" -- the run then makes a superset of the
 *                         calls, in the same order -- queues the pairs and, at exit, scores the whole
 *                         queue with ONE indelgpu_indel_support_batch into $INDELGPU_REPLAY_FILE.support;
 *   INDELGPU_MODE=replay  answers every call from that file, skipping the records of the calls it no
 *                         longer makes (each record carries the call's coordinates);
 *   INDELGPU_MODE=auto    both in one command (the fork lives in indelgpu_attempt.c); the replaying parent
 *                         waits for the recording child before its first call here.
 *   INDELGPU_MODE=inline  no second run and no fork: is_indel_supported's bam_fetch (variant.c:1567; the hooked
 *                         variant.c is compiled with -Dbam_fetch=indelgpu_bam_fetch_support) reads the records
 *                         of the variant's region once and shows them to check_for_indel TWICE.  The first time
 *                         this file answers "does not support" and queues the pairs; if no read supported the
 *                         variant by its CIGAR alone, the queue is scored with one indelgpu_indel_support_batch
 *                         and the second pass answers from it, in call order.  The only result of the loop is
 *                         variant->diffsample_support, an OR over the reads (variant.c:1444-1448, :1549-1553),
 *                         so showing the records twice changes nothing else.
 * Without INDELGPU_MODE every call is a 1-pair batch.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "variant.h"        /* the reference's headers: knownvariant, readaln, varianttype */
#include "errors.h"
#include "memalloc.h"

#include "indelgpu.h"
#include "indelgpu_support.h"

#include "indelgpu_glue.h"      /* shared with host/indelgpu_attempt.c: the operating mode */

static indelgpu_ctx* g_ctx = NULL;

/* what identifies a call: enough to pair a replayed call with its record */
typedef struct { int32_t tid, vstart, vstop, vtype, rstart, rstop, qstart, qstop, len1, len2; } callkey;

static callkey* g_keys = NULL;  static int64_t g_n = 0, g_cap = 0;
static uint8_t* g_tgt = NULL;   static int64_t g_ntgt = 0, g_captgt = 0;
static uint8_t* g_qry = NULL;   static int64_t g_nqry = 0, g_capqry = 0;
static int64_t* g_toff = NULL;  static int64_t* g_qoff = NULL;

static char* support_path(void)
{
    static char path[4096];
    snprintf(path, sizeof(path), "%s.support", indelgpu_glue_replay_path());
    return path;
}

static void support_ctx(void)
{
    if (g_ctx != NULL) return;
    const char* dev = getenv("INDELGPU_DEVICE");       /* same device as the realignment glue (one process per GPU) */
    g_ctx = indelgpu_create(dev ? atoi(dev) : 0, NULL);
    if (g_ctx == NULL) fatalf("libindelgpu: %s", indelgpu_last_error());
}

/* inline mode: which pass of indelgpu_bam_fetch_support is running */
enum { PASS_NONE = 0, PASS_COLLECT = 1, PASS_ANSWER = 2 };
static int g_pass = PASS_NONE;
static int32_t* g_ans = NULL;   static int64_t g_anscap = 0, g_anspos = 0;     /* subs, indels, aligned per queued pair */
static long long g_inl_variants = 0, g_inl_batches = 0, g_inl_pairs = 0;

static void flush_support(void)
{
    const int64_t n = g_n;
    int32_t* subs = ckallocz(sizeof(int32_t) * (size_t)(n + 1));
    int32_t* indels = ckallocz(sizeof(int32_t) * (size_t)(n + 1));
    int32_t* aligned = ckallocz(sizeof(int32_t) * (size_t)(n + 1));
    if (n > 0) {
        g_toff[n] = g_ntgt; g_qoff[n] = g_nqry;
        if (indelgpu_indel_support_batch(g_ctx, (int32_t)n, g_tgt, g_toff, g_qry, g_qoff, subs, indels, aligned, NULL) != 0)
            fatalf("libindelgpu: indelgpu_indel_support_batch (record mode): %s", indelgpu_last_error());
    }
    FILE* f = fopen(support_path(), "wb");
    if (f == NULL) fatalf("libindelgpu: cannot write %s", support_path());
    const int64_t magic = 0x3154525050534449LL;          /* "IDSPPRT1" */
    fwrite(&magic, 8, 1, f); fwrite(&n, 8, 1, f);
    for (int64_t i = 0; i < n; i++) {
        fwrite(&g_keys[i], sizeof(callkey), 1, f);
        fwrite(&subs[i], 4, 1, f); fwrite(&indels[i], 4, 1, f); fwrite(&aligned[i], 4, 1, f);
    }
    if (fclose(f) != 0) fatalf("libindelgpu: error writing %s", support_path());
    fprintf(stderr, "libindelgpu: %lld (variant, read) pairs scored in one batch, results in %s\n", (long long)n, support_path());
}

typedef struct { callkey key; int32_t subs, indels, aligned; } record;
static record* g_rec = NULL;  static int64_t g_nrec = 0, g_pos = 0;  static int g_loaded = 0;

static void load_support(void)
{
    /* auto mode: the recording run writes this file when it exits; take delivery of the rest of its stream
     * and wait for it (it is usually far ahead: it answers every call at once) */
    if (indelgpu_glue_recording_runs()) indelgpu_glue_wait_recording();
    FILE* f = fopen(support_path(), "rb");
    if (f == NULL) fatalf("libindelgpu: cannot read %s (run with INDELGPU_MODE=record first)", support_path());
    int64_t hdr[2];
    if (fread(hdr, 8, 2, f) != 2 || hdr[0] != 0x3154525050534449LL) fatalf("libindelgpu: %s is not a support replay file", support_path());
    g_nrec = hdr[1];
    g_rec = ckalloc(sizeof(record) * (size_t)(g_nrec + 1));
    if (g_nrec > 0 && fread(g_rec, sizeof(record), (size_t)g_nrec, f) != (size_t)g_nrec) fatalf("libindelgpu: short read on %s", support_path());
    fclose(f);
    if (indelgpu_glue_replay_is_temporary()) unlink(support_path());
    g_loaded = 1;
}

void realign_with_indel(const char* const reference, const int rstart, const int rstop,
                        const readaln* const rln, const int qstart, const int qstop,
                        const knownvariant* const variant,
                        int* const alnsubs, int* const alnindels, int* const alnaligned)
{
    const char* query = rln->segments->sequence;
    /* the query slice (variant.c:1278-1283) */
    const char* t2 = query + qstart;
    int len2 = qstop - qstart;
    {
        const int rest = (int)strlen(t2);
        if (rest < len2) len2 = rest;
    }
    const int mode = indelgpu_glue_mode();
    const int alen = (int)strlen(variant->alternate) - 1;
    const int vs = (int)variant->start, ve = (int)variant->stop;
    if (variant->type != DELETION && variant->type != INSERTION) fatalf("Unknown type of variant: %d\n", variant->type);
    if (rstart < 0 || rstop < rstart || vs < rstart) fatalf("libindelgpu: realign_with_indel: interval [%d, %d) does not hold the variant at %d", rstart, rstop, vs);

    callkey key;
    memset(&key, 0, sizeof(key));
    key.tid = variant->tid; key.vstart = vs; key.vstop = ve; key.vtype = (int32_t)variant->type;
    key.rstart = rstart; key.rstop = rstop; key.qstart = qstart; key.qstop = qstop; key.len2 = len2;

    if (mode == MODE_REPLAY) {
        if (!g_loaded) load_support();
        /* the replay run stops looking at a variant's reads once one supports it: skip those records */
        while (g_pos < g_nrec) {
            const callkey* k = &g_rec[g_pos].key;
            if (k->tid == key.tid && k->vstart == key.vstart && k->vstop == key.vstop && k->vtype == key.vtype &&
                k->rstart == key.rstart && k->rstop == key.rstop && k->qstart == key.qstart && k->qstop == key.qstop &&
                k->len2 == key.len2) break;
            g_pos++;
        }
        if (g_pos >= g_nrec) fatalf("libindelgpu: support replay file exhausted (different command line than the recording run?)");
        *alnsubs = g_rec[g_pos].subs; *alnindels = g_rec[g_pos].indels; *alnaligned = g_rec[g_pos].aligned;
        g_pos++;
        return;
    }

    /* the target: the reference interval with the variant in it (variant.c:1259-1275) */
    const int len0 = rstop - rstart;
    char* target = ckallocz((size_t)len0 + (size_t)(alen > 0 ? alen : 0) + 2);
    const int head = vs - rstart;                        /* bases before the variant's anchor base */
    int len1;
    if (variant->type == DELETION) {
        /* keep [rstart, vstart), then continue at vstop - 1: the deleted bases vstart .. vstop-2 drop out */
        memcpy(target, reference + rstart, (size_t)head);
        const int tail0 = ve - 1;
        const int ntail = rstop - tail0 > 0 ? rstop - tail0 : 0;
        memcpy(target + head, reference + tail0, (size_t)ntail);
        len1 = head + ntail;
        /* the reference moves rstop - vstop + 2 bytes, i.e. one byte past its copy's NUL: the string ends at the NUL */
    } else {
        memcpy(target, reference + rstart, (size_t)head);
        memcpy(target + head, variant->alternate + 1, (size_t)alen);
        memcpy(target + head + alen, reference + vs, (size_t)(rstop - vs));
        len1 = len0 + alen;
    }
    target[len1] = '\0';
    len1 = (int)strlen(target);                          /* a NUL inside the interval ends it, as strlen does there */
    key.len1 = len1;

    support_ctx();
    if (mode == MODE_INLINE && g_pass == PASS_ANSWER) {
        if (g_anspos >= g_n) fatalf("libindelgpu: inline support: the second pass makes more calls than the first");
        const callkey* k = &g_keys[g_anspos];
        if (k->rstart != key.rstart || k->rstop != key.rstop || k->qstart != key.qstart || k->qstop != key.qstop ||
            k->len1 != key.len1 || k->len2 != key.len2) fatalf("libindelgpu: inline support: the second pass is out of step");
        *alnsubs = g_ans[3 * g_anspos]; *alnindels = g_ans[3 * g_anspos + 1]; *alnaligned = g_ans[3 * g_anspos + 2];
        g_anspos++;
        ckfree(target);
        return;
    }
    if (mode == MODE_RECORD || (mode == MODE_INLINE && g_pass == PASS_COLLECT)) {
        static int registered = 0;
        if (mode == MODE_RECORD && !registered) { atexit(flush_support); registered = 1; }     /* after the CUDA runtime's own handler */
        if (g_n + 2 > g_cap) {
            g_cap = g_cap ? 2 * g_cap : 4096;
            g_keys = ckrealloc(g_keys, sizeof(callkey) * (size_t)g_cap);
            g_toff = ckrealloc(g_toff, sizeof(int64_t) * (size_t)(g_cap + 1));
            g_qoff = ckrealloc(g_qoff, sizeof(int64_t) * (size_t)(g_cap + 1));
        }
        if (g_ntgt + len1 + 16 > g_captgt) { g_captgt = 2 * (g_captgt + len1) + 4096; g_tgt = ckrealloc(g_tgt, (size_t)g_captgt); }
        if (g_nqry + len2 + 16 > g_capqry) { g_capqry = 2 * (g_capqry + len2) + 4096; g_qry = ckrealloc(g_qry, (size_t)g_capqry); }
        g_keys[g_n] = key; g_toff[g_n] = g_ntgt; g_qoff[g_n] = g_nqry;
        memcpy(g_tgt + g_ntgt, target, (size_t)len1); g_ntgt += len1;
        memcpy(g_qry + g_nqry, t2, (size_t)(len2 > 0 ? len2 : 0)); g_nqry += len2 > 0 ? len2 : 0;
        g_n++;
        ckfree(target);
        *alnsubs = INT_MAX; *alnindels = INT_MAX; *alnaligned = 0;      /* "does not support": every later call is made too */
        return;
    }

    const int64_t toff[2] = {0, len1}, qoff[2] = {0, len2 > 0 ? len2 : 0};
    int32_t s = 0, g = 0, a = 0;
    const uint8_t dummy = 0;
    pthread_mutex_lock(&indelgpu_glue_gpu_mu);
    const int rc = indelgpu_indel_support_batch(g_ctx, 1, len1 > 0 ? (const uint8_t*)target : &dummy, toff,
                                                len2 > 0 ? (const uint8_t*)t2 : &dummy, qoff, &s, &g, &a, NULL);
    pthread_mutex_unlock(&indelgpu_glue_gpu_mu);
    if (rc != 0)
        fatalf("libindelgpu: indelgpu_indel_support_batch: %s", indelgpu_last_error());
    ckfree(target);
    *alnsubs = s; *alnindels = g; *alnaligned = a;
}

/* ---- inline mode: is_indel_supported's record loop (variant.c:1567) ---------------------------------- */
static void print_inline_stats(void)
{
    fprintf(stderr, "libindelgpu: inline mode: %lld known variants checked, %lld (variant, read) pairs scored in %lld batches\n",
            g_inl_variants, g_inl_pairs, g_inl_batches);
}

int indelgpu_bam_fetch_support(bamFile fp, const bam_index_t* idx, int tid, int beg, int end, void* data, bam_fetch_f func)
{
    if (!indelgpu_glue_env_is_inline()) return bam_fetch(fp, idx, tid, beg, end, data, func);   /* no fork from here (see indelgpu_inline.c) */
    static bam1_t* recs = NULL; static int caprec = 0;
    static int registered = 0;
    if (!registered) { atexit(print_inline_stats); registered = 1; }
    int nrec = 0, ret;
    bam_iter_t iter = bam_iter_query(idx, tid, beg, end);
    for (;;) {
        if (nrec == caprec) {
            const int ncap = caprec ? 2 * caprec : 256;
            recs = ckrealloc(recs, sizeof(bam1_t) * (size_t)ncap);
            memset(recs + caprec, 0, sizeof(bam1_t) * (size_t)(ncap - caprec));
            caprec = ncap;
        }
        ret = bam_iter_read(fp, iter, &recs[nrec]);
        if (ret < 0) break;
        nrec++;
    }
    bam_iter_destroy(iter);
    knownvariant* variant = data;
    g_inl_variants++;
    g_n = 0; g_ntgt = 0; g_nqry = 0;
    g_pass = PASS_COLLECT;
    for (int i = 0; i < nrec; i++) func(&recs[i], data);
    g_pass = PASS_NONE;
    if (!variant->diffsample_support && g_n > 0) {
        if (3 * g_n > g_anscap) { g_anscap = 6 * g_n + 64; g_ans = ckrealloc(g_ans, sizeof(int32_t) * (size_t)g_anscap); }
        int32_t* subs = ckalloc(sizeof(int32_t) * (size_t)g_n * 3);
        g_toff[g_n] = g_ntgt; g_qoff[g_n] = g_nqry;
        const uint8_t dummy = 0;
        pthread_mutex_lock(&indelgpu_glue_gpu_mu);
        const int rc = indelgpu_indel_support_batch(g_ctx, (int32_t)g_n, g_ntgt > 0 ? g_tgt : &dummy, g_toff, g_nqry > 0 ? g_qry : &dummy, g_qoff,
                                                    subs, subs + g_n, subs + 2 * g_n, NULL);
        pthread_mutex_unlock(&indelgpu_glue_gpu_mu);
        if (rc != 0) fatalf("libindelgpu: indelgpu_indel_support_batch (inline mode): %s", indelgpu_last_error());
        for (int64_t i = 0; i < g_n; i++) { g_ans[3 * i] = subs[i]; g_ans[3 * i + 1] = subs[g_n + i]; g_ans[3 * i + 2] = subs[2 * g_n + i]; }
        ckfree(subs);
        g_inl_batches++; g_inl_pairs += g_n;
        g_anspos = 0;
        g_pass = PASS_ANSWER;
        for (int i = 0; i < nrec; i++) func(&recs[i], data);
        g_pass = PASS_NONE;
    }
    g_n = 0; g_ntgt = 0; g_nqry = 0;
    return ret == -1 ? 0 : ret;
}
