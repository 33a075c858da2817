/* Host-side batch builder in C for libindelgpu.so: packs candidate reads into the pinned SoA
 * buffers indelgpu_realign_batch takes (include/indelgpu.h) and owns the result buffers.
 * This is the packing half of the batched driver sketched in INTEGRATION.md section 3: fetch_func
 * (src/indelminer.c:411,486) pushes (read, mate contig, mate position, range[1]) instead of calling
 * attempt_pe_alignment, and consumes the results in the same order after igb_run().
 * Plain C99; no alignment code, no CPU fallback. */
#ifndef INDELGPU_BATCH_H
#define INDELGPU_BATCH_H

#include <stdint.h>

#include "indelgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct igb_batch {
    /* inputs (pinned) */
    int32_t  n, cap_reads;
    int64_t  nbases, cap_bases;
    uint8_t* bases;
    int64_t* off;           /* n + 1 */
    int32_t* tid;
    int32_t* position;
    int32_t* range1;
    /* outputs (pinned), valid after igb_run */
    int32_t*  status;
    int32_t*  nseg;
    int32_t*  rstart;
    int64_t*  seg_off;
    uint32_t* segs;
    int64_t   seg_cap, seg_count;
} igb_batch;

/* returns NULL when pinned memory cannot be allocated */
igb_batch* igb_create(int32_t max_reads, int64_t max_bases);
void       igb_destroy(igb_batch* b);
void       igb_clear(igb_batch* b);

/* 0 on success, -1 when the batch is full (run it, consume, clear, push again) */
int igb_push(igb_batch* b, const char* read, int32_t readlen, int32_t tid, int32_t position, int32_t range1);

/* one indelgpu_realign_batch over everything pushed; returns its error code */
int igb_run(igb_batch* b, indelgpu_ctx* ctx);

/* segment words of read i (BAM-style, len << 4 | op) and the reference position they start at */
static inline const uint32_t* igb_segments(const igb_batch* b, int32_t i, int32_t* nseg, int32_t* rstart)
{
    *nseg = b->nseg[i]; *rstart = b->rstart[i];
    return b->segs + b->seg_off[i];
}

#ifdef __cplusplus
}
#endif
#endif
