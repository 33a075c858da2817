/* Declarations shared by the C glue files of this directory (host/indelgpu_attempt.c,
 * host/indelgpu_inline.c, host/indelgpu_support.c, host/indelgpu_bamcache.c).  Internal to the glue: the
 * reference-facing symbols are attempt_pe_alignment (alignment.h:21-25), realign_with_indel
 * (indelgpu_support.h) and the redirected samtools entry points named below. */
#ifndef INDELGPU_GLUE_H
#define INDELGPU_GLUE_H

#include <pthread.h>
#include <stdint.h>

#include "bam.h"            /* the reference's bundled samtools-0.1.19 */
#include "indelgpu.h"

enum { MODE_DIRECT = 0, MODE_RECORD = 1, MODE_REPLAY = 2, MODE_INLINE = 3 };

/* operating mode of this process, decided once from $INDELGPU_MODE (indelgpu_attempt.c) */
int indelgpu_glue_mode(void);
/* 1 when $INDELGPU_MODE asks for the inline mode; unlike indelgpu_glue_mode() this never forks */
int indelgpu_glue_env_is_inline(void);
const char* indelgpu_glue_replay_path(void);
int indelgpu_glue_replay_is_temporary(void);
int indelgpu_glue_recording_runs(void);
void indelgpu_glue_wait_recording(void);

/* every libindelgpu call of the glue is made under this lock: the prefetching thread of the inline mode
 * and the main thread share the contexts */
extern pthread_mutex_t indelgpu_glue_gpu_mu;

/* the context that holds contig `tid` if one exists already (never creates one: the reference strings are
 * only known to attempt_pe_alignment); *ptid = the contig's index inside that context */
indelgpu_ctx* indelgpu_glue_ctx_peek(int32_t tid, int32_t* ptid);

/* ---- inline mode (row f2): host/indelgpu_inline.c -------------------------------------------------
 * indelminer.c is compiled with -Dbam_fetch=indelgpu_bam_fetch (no source change): the per-contig
 * bam_fetch(fp, idx, tid, beg, end, alndata, fetch_func) of indelminer.c:797,800 lands here. */
int indelgpu_bam_fetch(bamFile fp, const bam_index_t* idx, int tid, int beg, int end, void* data, bam_fetch_f func);

/* answers one attempt_pe_alignment call from the prefetched batch of the record being processed;
 * 1 = hit (outputs valid until the next record), 0 = not prefetched (compute it directly) */
int indelgpu_inline_lookup(int32_t tid, int32_t position, int32_t range1, const char* read, int32_t readlen,
                           int32_t* nseg, int32_t* rstart, const uint32_t** words);
/* a call that was computed directly: remember range[1] of the current record's read group */
void indelgpu_inline_learn(int32_t range1);

/* ---- cached BAM handles (rows f3 / f2b): host/indelgpu_bamcache.c ----------------------------------
 * shared.c and variant.c are compiled with -Dbgzf_open=indelgpu_bgzf_open -Dbgzf_close=indelgpu_bgzf_close
 * -Dbam_index_load=indelgpu_bam_index_load -Dbam_index_destroy=indelgpu_bam_index_destroy: the
 * open / load-index / fetch / close / destroy sequence calculate_cov_params (shared.c:178-212) and
 * is_indel_supported (variant.c:1561-1572) run once per printed variant reuses one handle and one index. */
/* shared.c only: -Dbam_fetch=indelgpu_bam_fetch_cov (host/indelgpu_inline.c): the per-variant coverage fetch of
 * calculate_cov_params, answered from the records the inline driver still holds when they cover the region */
int indelgpu_bam_fetch_cov(bamFile fp, const bam_index_t* idx, int tid, int beg, int end, void* data, bam_fetch_f func);
BGZF* indelgpu_bgzf_open(const char* path, const char* mode);
int indelgpu_bgzf_close(BGZF* fp);
bam_index_t* indelgpu_bam_index_load(const char* fn);
void indelgpu_bam_index_destroy(bam_index_t* idx);

#endif
