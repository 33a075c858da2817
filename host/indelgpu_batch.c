/* see indelgpu_batch.h */
#include "indelgpu_batch.h"

#include <stdlib.h>
#include <string.h>

static void* pinned(size_t bytes) { return indelgpu_host_alloc(bytes ? bytes : 1); }

igb_batch* igb_create(int32_t max_reads, int64_t max_bases)
{
    if (max_reads <= 0 || max_bases <= 0) return NULL;
    igb_batch* b = (igb_batch*)calloc(1, sizeof(igb_batch));
    if (!b) return NULL;
    b->cap_reads = max_reads; b->cap_bases = max_bases;
    b->seg_cap = indelgpu_seg_bound(max_reads, max_bases);
    b->bases    = (uint8_t*)pinned((size_t)max_bases + 16);
    b->off      = (int64_t*)pinned(sizeof(int64_t) * ((size_t)max_reads + 1));
    b->tid      = (int32_t*)pinned(sizeof(int32_t) * (size_t)max_reads);
    b->position = (int32_t*)pinned(sizeof(int32_t) * (size_t)max_reads);
    b->range1   = (int32_t*)pinned(sizeof(int32_t) * (size_t)max_reads);
    b->status   = (int32_t*)pinned(sizeof(int32_t) * (size_t)max_reads);
    b->nseg     = (int32_t*)pinned(sizeof(int32_t) * (size_t)max_reads);
    b->rstart   = (int32_t*)pinned(sizeof(int32_t) * (size_t)max_reads);
    b->seg_off  = (int64_t*)pinned(sizeof(int64_t) * (size_t)max_reads);
    b->segs     = (uint32_t*)pinned(sizeof(uint32_t) * (size_t)b->seg_cap);
    if (!b->bases || !b->off || !b->tid || !b->position || !b->range1 || !b->status || !b->nseg ||
        !b->rstart || !b->seg_off || !b->segs) { igb_destroy(b); return NULL; }
    igb_clear(b);
    return b;
}

void igb_destroy(igb_batch* b)
{
    if (!b) return;
    indelgpu_host_free(b->bases); indelgpu_host_free(b->off); indelgpu_host_free(b->tid);
    indelgpu_host_free(b->position); indelgpu_host_free(b->range1); indelgpu_host_free(b->status);
    indelgpu_host_free(b->nseg); indelgpu_host_free(b->rstart); indelgpu_host_free(b->seg_off);
    indelgpu_host_free(b->segs);
    free(b);
}

void igb_clear(igb_batch* b)
{
    b->n = 0; b->nbases = 0; b->seg_count = 0;
    b->off[0] = 0;
}

int igb_push(igb_batch* b, const char* read, int32_t readlen, int32_t tid, int32_t position, int32_t range1)
{
    if (readlen <= 0 || b->n >= b->cap_reads || b->nbases + readlen > b->cap_bases) return -1;
    memcpy(b->bases + b->nbases, read, (size_t)readlen);
    b->nbases += readlen;
    b->tid[b->n] = tid; b->position[b->n] = position; b->range1[b->n] = range1;
    b->n++;
    b->off[b->n] = b->nbases;
    return 0;
}

int igb_run(igb_batch* b, indelgpu_ctx* ctx)
{
    indelgpu_batch in;
    in.n = b->n; in.read_bases = b->bases; in.read_off = b->off;
    in.tid = b->tid; in.position = b->position; in.range1 = b->range1;
    indelgpu_result out;
    memset(&out, 0, sizeof(out));
    out.status = b->status; out.nseg = b->nseg; out.rstart = b->rstart; out.seg_off = b->seg_off;
    out.segs = b->segs; out.seg_capacity = b->seg_cap;
    const int rc = indelgpu_realign_batch(ctx, &in, &out);
    b->seg_count = out.seg_count;
    return rc;
}
