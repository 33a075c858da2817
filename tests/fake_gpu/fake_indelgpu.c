/* TEST INFRASTRUCTURE ONLY -- never shipped, never linked by the product.
 *
 * A stand-in for libindelgpu.so that answers the C ABI of include/indelgpu.h with the CPU oracle
 * (oracle/indel_oracle.c).  It exists so that the HOST logic of the glue (host/indelgpu_inline.c's
 * producer / consumer, the prefetch-cache keys, the two-pass support loop, the BAM handle cache) can be
 * tested in the build container, which has no GPU: tests/test_inline_host_logic.py links the reference
 * program + the glue against this file and requires byte-identical VCFs.  The GPU tests (-m gpu) run
 * the same programs against the real library.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/indelgpu.h"
#include "../../oracle/indel_oracle.h"

struct indelgpu_ctx {
    orc_params p;
    int ncontigs;
    char** seq; int64_t* len;
};

static char g_err[256] = "";
const char* indelgpu_last_error(void) { return g_err; }
int indelgpu_version(void) { return INDELGPU_VERSION; }
int indelgpu_device_count(void) { return 1; }

void indelgpu_default_params(indelgpu_params* p)
{
    p->klength = 6; p->numgaps = 0; p->maxdelsize = 1000; p->ethreshold = 10;
    p->match = 1; p->mismatch = -10; p->gapopen = 10; p->gapextend = 10;
}

indelgpu_ctx* indelgpu_create(int device, const indelgpu_params* p)
{
    (void)device;
    indelgpu_params def;
    if (!p) { indelgpu_default_params(&def); p = &def; }
    indelgpu_ctx* c = calloc(1, sizeof(*c));
    c->p.klength = p->klength; c->p.numgaps = p->numgaps; c->p.maxdelsize = p->maxdelsize; c->p.ethreshold = p->ethreshold;
    c->p.match = p->match; c->p.mismatch = p->mismatch; c->p.gapopen = p->gapopen; c->p.gapextend = p->gapextend;
    return c;
}

void indelgpu_destroy(indelgpu_ctx* c)
{
    if (!c) return;
    for (int i = 0; i < c->ncontigs; i++) free(c->seq[i]);
    free(c->seq); free(c->len); free(c);
}

void* indelgpu_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void indelgpu_host_free(void* p) { free(p); }

int indelgpu_set_reference(indelgpu_ctx* c, int32_t n, const char* const* sequences, const int64_t* lengths)
{
    c->ncontigs = n;
    c->seq = calloc((size_t)n, sizeof(char*)); c->len = calloc((size_t)n, sizeof(int64_t));
    for (int i = 0; i < n; i++) {
        c->seq[i] = malloc((size_t)lengths[i] + 1);
        memcpy(c->seq[i], sequences[i], (size_t)lengths[i]);
        c->seq[i][lengths[i]] = '\0';
        c->len[i] = lengths[i];
    }
    return 0;
}

int64_t indelgpu_seg_bound(int32_t n, int64_t total_read_bases) { return 2 * total_read_bases + 8LL * n + 16; }

int indelgpu_realign_batch(indelgpu_ctx* c, const indelgpu_batch* h, indelgpu_result* o)
{
    orc_result* r = malloc(sizeof(orc_result));
    int64_t used = 0;
    for (int i = 0; i < h->n; i++) {
        const int t = h->tid[i];
        if (t < 0 || t >= c->ncontigs) { snprintf(g_err, sizeof(g_err), "bad contig"); free(r); return INDELGPU_EINVAL; }
        orc_realign_read(&c->p, c->seq[t], (int)c->len[t], h->position[i], h->range1[i],
                         (const char*)h->read_bases + h->read_off[i], (int)(h->read_off[i + 1] - h->read_off[i]), r, NULL);
        o->status[i] = r->status; o->nseg[i] = r->nseg; o->rstart[i] = r->nseg > 0 ? r->seg_start[0] : 0;
        o->seg_off[i] = used;
        if (used + r->nseg > o->seg_capacity) { snprintf(g_err, sizeof(g_err), "segment buffer too small"); free(r); return INDELGPU_ELIMIT; }
        for (int s = 0; s < r->nseg; s++) o->segs[used++] = ((uint32_t)r->seg_len[s] << 4) | (uint32_t)r->seg_op[s];
    }
    o->seg_count = used;
    free(r);
    return 0;
}

int indelgpu_indel_support_batch(indelgpu_ctx* c, int32_t n, const uint8_t* tg, const int64_t* toff, const uint8_t* q,
                                 const int64_t* qoff, int32_t* subs, int32_t* indels, int32_t* aligned, int64_t* cells)
{
    (void)c;
    long long cc = 0;
    for (int i = 0; i < n; i++) {
        int s, g, a;
        orc_indel_support_dp((const char*)tg + toff[i], (int)(toff[i + 1] - toff[i]), (const char*)q + qoff[i],
                             (int)(qoff[i + 1] - qoff[i]), &s, &g, &a, &cc);
        subs[i] = s; indels[i] = g; aligned[i] = a;
    }
    if (cells) *cells = cc;
    return 0;
}
