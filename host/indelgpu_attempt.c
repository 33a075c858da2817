/* Host-side glue, in C like the reference's host code: a replacement for the translation unit
 * src/alignment.c that keeps its one public entry point
 *
 *     evidence* attempt_pe_alignment(char** const sequences, const int32_t tid,
 *                                    const int32_t position, const int* const range,
 *                                    readaln* const rln);           (src/alignment.h:21-25)
 *
 * with the same arguments, result and ownership rules, and computes it on the GPU through the
 * C ABI of libindelgpu.so (include/indelgpu.h).  Link it INSTEAD of alignment.o, localalign.o and
 * globalalign.o (src/Makefile:75-80); fetch_func (src/indelminer.c:411,486) and everything
 * downstream (evidence.c, graph.c, variant.c, VCF output) stay byte-for-byte the reference's.
 *
 * It includes the reference's own headers (readaln.h, evidence.h, slinklist.h), found with
 * -I<reference>/src at build time; nothing of the reference is copied here.  What this file does on
 * the host is exactly what the reference does AFTER its alignments are known:
 *   - rebuild rln->segments from the segment words with new_readseg (readaln.c:24-99), the way
 *     update_readsegs (readaln.c:348-458) assembles its list,
 *   - one evidence per D / I segment, prepended (add_evidence_from_segment, alignment.c:449-476).
 * There is no alignment code and no CPU fallback in this file.
 *
 * Two ways to use it:
 *   1. unchanged caller: nothing to add.  The flags are read from the reference's own globals
 *      (indelminer.c:31-44) on first use and each contig is uploaded the first time a read is
 *      anchored on it (one small context per contig, because `char** sequences` carries no count).
 *   2. one added line after read_reference (indelminer.c:771): indelgpu_host_init(sequences,
 *      hdr->n_targets) uploads the whole reference into one context up front.
 * Both process one read per call (a 1-element batch).  The batched driver -- collect candidates,
 * one indelgpu_realign_batch per READCHUNK, replay evidence in read order -- is described in
 * INTEGRATION.md.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "alignment.h"      /* the reference's header: evidence, readaln, constants */
#include "slinklist.h"
#include "errors.h"

#include "indelgpu.h"

/* the flags alignment.c reads (alignment.c:3-9; defined in indelminer.c:31-44) */
extern uint klength;
extern uint numgaps;
extern uint maxdelsize;
extern uint ethreshold;

typedef struct {
    indelgpu_ctx* ctx;
} contig_slot;

static indelgpu_ctx* g_all = NULL;       /* mode 2: one context holding every contig       */
static contig_slot* g_slots = NULL;      /* mode 1: lazily created, one context per contig */
static int g_nslots = 0;
static uint32_t* g_segs = NULL;          /* result buffer of the 1-element batch            */
static int64_t g_segcap = 0;

static void gpu_die(const char* what)
{
    /* the reference's convention for errors on this path: message + exit (errors.c:15-27) */
    fatalf("libindelgpu: %s: %s", what, indelgpu_last_error());
}

static indelgpu_ctx* make_ctx(void)
{
    indelgpu_params p;
    indelgpu_default_params(&p);
    p.klength = (int32_t)klength;
    p.numgaps = (int32_t)numgaps;
    p.maxdelsize = (int32_t)maxdelsize;
    p.ethreshold = (int32_t)ethreshold;
    const char* dev = getenv("INDELGPU_DEVICE");
    indelgpu_ctx* c = indelgpu_create(dev ? atoi(dev) : 0, &p);
    if (c == NULL) gpu_die("indelgpu_create");
    return c;
}

/* optional: upload the whole reference once (call after read_reference, shared.c:46-82) */
void indelgpu_host_init(char** const sequences, const int ncontigs)
{
    int64_t* lengths = ckalloc(sizeof(int64_t) * (size_t)ncontigs);
    for (int i = 0; i < ncontigs; i++) lengths[i] = (int64_t)strlen(sequences[i]);
    g_all = make_ctx();
    if (indelgpu_set_reference(g_all, ncontigs, (const char* const*)sequences, lengths) != 0)
        gpu_die("indelgpu_set_reference");
    ckfree(lengths);
}

static indelgpu_ctx* ctx_for_contig(char** const sequences, const int32_t tid, int32_t* ptid)
{
    if (g_all != NULL) { *ptid = tid; return g_all; }
    if (tid >= g_nslots) {
        const int n = tid + 16;
        g_slots = ckrealloc(g_slots, sizeof(contig_slot) * (size_t)n);
        for (int i = g_nslots; i < n; i++) g_slots[i].ctx = NULL;
        g_nslots = n;
    }
    if (g_slots[tid].ctx == NULL) {
        const char* seq = sequences[tid];
        const int64_t len = (int64_t)strlen(seq);        /* once per contig, not per read (alignment.c:771) */
        g_slots[tid].ctx = make_ctx();
        if (indelgpu_set_reference(g_slots[tid].ctx, 1, &seq, &len) != 0) gpu_die("indelgpu_set_reference");
    }
    *ptid = 0;
    return g_slots[tid].ctx;
}

evidence* attempt_pe_alignment(char** const sequences,
                               const int32_t tid,
                               const int32_t position,
                               const int* const range,
                               readaln* const rln)
{
    forceassert(range[0] <= range[1]);                   /* alignment.c:773 */
    const char* read = rln->segments->sequence;          /* the single S segment (readaln.c:258-264) */
    const int64_t readlength = (int64_t)strlen(read);

    int32_t dtid = 0;
    indelgpu_ctx* ctx = ctx_for_contig(sequences, tid, &dtid);

    const int64_t need = indelgpu_seg_bound(1, readlength);
    if (need > g_segcap) {
        g_segs = ckrealloc(g_segs, sizeof(uint32_t) * (size_t)need);
        g_segcap = need;
    }
    const int64_t off[2] = {0, readlength};
    const int32_t range1 = range[1];
    indelgpu_batch in = {1, (const uint8_t*)read, off, &dtid, &position, &range1};
    int32_t status = 0, nseg = 0, rstart = 0;
    int64_t segoff = 0;
    indelgpu_result out;
    memset(&out, 0, sizeof(out));
    out.status = &status; out.nseg = &nseg; out.rstart = &rstart; out.seg_off = &segoff;
    out.segs = g_segs; out.seg_capacity = g_segcap;
    if (indelgpu_realign_batch(ctx, &in, &out) != 0) gpu_die("indelgpu_realign_batch");

    if (nseg == 0) return NULL;                          /* every NULL exit of alignment.c:568-751 */

    /* update_readsegs' list construction (readaln.c:355-457) from the stitched words */
    int refindx = rstart, readindx = 0;
    readseg* readsegs = NULL;
    for (int i = 0; i < nseg; i++) {
        readseg* rsg = new_readseg(read, g_segs[segoff + i], &refindx, &readindx);
        sladdhead(&readsegs, rsg);
    }
    free_readsegs(&rln->segments);
    slreverse(&readsegs);
    rln->segments = readsegs;

    /* add_evidence_from_segment(rln, NULL) (alignment.c:449-476) */
    evidence* allevidence = NULL;
    for (readseg* iter = rln->segments; iter; iter = iter->next) {
        if (iter->op == BAM_CDEL) {
            evidence* evdnc = new_evidence(rln, iter, DELETION, SPLIT_READ);
            sladdhead(&allevidence, evdnc);
        } else if (iter->op == BAM_CINS) {
            evidence* evdnc = new_evidence(rln, iter, INSERTION, SPLIT_READ);
            sladdhead(&allevidence, evdnc);
        }
    }
    free_readsegs(&rln->segments);
    return allevidence;
}
