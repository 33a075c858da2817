/* Rows f3 / f2b of SURVEY.md section 8: take the per-variant BAM re-opening out of the critical path.
 *
 * print_vcf_output calls calculate_cov_params once per printed variant (src/variant.c:303-306), and
 * calculate_cov_params opens the BAM, loads the whole .bai, fetches one small region, closes both again
 * (src/shared.c:183-184, :207-208); is_indel_supported does the same per known variant in annotate mode
 * (src/variant.c:1564-1570).  Once the alignments are free this dominates the wall time (about 8 k variants
 * at 16 Mb, 32 k at config 3).
 *
 * The build compiles shared.c and variant.c -- unchanged -- with
 *     -Dbgzf_open=indelgpu_bgzf_open  -Dbgzf_close=indelgpu_bgzf_close
 *     -Dbam_index_load=indelgpu_bam_index_load  -Dbam_index_destroy=indelgpu_bam_index_destroy
 * (bam_open / bam_close are macros over bgzf_open / bgzf_close, bam.h:96-98), so those four calls land
 * here: the first open of a file for reading really opens it, later ones hand out the same handle while
 * it is not in use, and "closing" it only marks it free; the index of a file is loaded once.  bam_fetch
 * positions the handle itself (bam_iter_read seeks to the first chunk, bam_index.c:663-670), so a reused
 * handle returns exactly the records a fresh one would.  Handles opened while the cached one is busy,
 * or for writing, are ordinary handles.  The cache belongs to the process: region workers are separate
 * processes (tools/e2e_regions.py), each with its own handle and index -- the only read-side parallelism
 * the bundled BGZF offers (bgzf.c:419-424 multithreads writing only).
 * Output is unchanged; there is nothing GPU-specific in this file.
 */
#include <stdlib.h>
#include <string.h>

#include "bam.h"            /* bundled samtools-0.1.19, compiled WITHOUT the -D flags above */
#include "bgzf.h"
#include "memalloc.h"

#include "indelgpu_glue.h"

enum { MAXCACHED = 8 };
typedef struct { char* path; BGZF* fp; int busy; bam_index_t* idx; } cached_bam;
static cached_bam g_files[MAXCACHED];
static int g_nfiles = 0;

static cached_bam* slot_for(const char* path, int create)
{
    for (int i = 0; i < g_nfiles; i++)
        if (strcmp(g_files[i].path, path) == 0) return &g_files[i];
    if (!create || g_nfiles == MAXCACHED) return NULL;
    cached_bam* c = &g_files[g_nfiles++];
    c->path = ckallocz(strlen(path) + 1);
    strcpy(c->path, path);
    c->fp = NULL; c->busy = 0; c->idx = NULL;
    return c;
}

BGZF* indelgpu_bgzf_open(const char* path, const char* mode)
{
    if (mode == NULL || strchr(mode, 'r') == NULL || getenv("INDELGPU_NO_BAM_CACHE") != NULL) return bgzf_open(path, mode);
    cached_bam* c = slot_for(path, 1);
    if (c == NULL || c->busy) return bgzf_open(path, mode);
    if (c->fp == NULL) {
        c->fp = bgzf_open(path, mode);
        /* every fetch of calculate_cov_params starts at the linear-index offset of its 16 kb window and reads ~1 500
         * records up to the variant; variants are printed in position order, so consecutive fetches inflate the same
         * BGZF blocks again and again (3.7 x the whole file on a 30x data set).  The bundled BGZF keeps inflated
         * blocks when asked (-DBGZF_CACHE, bgzf.c:248-300): 64 MB hold the ~500 kb of genome between two flushes. */
        if (c->fp != NULL) bgzf_set_cache_size(c->fp, 64 << 20);
    }
    if (c->fp != NULL) c->busy = 1;
    return c->fp;
}

int indelgpu_bgzf_close(BGZF* fp)
{
    for (int i = 0; i < g_nfiles; i++)
        if (g_files[i].fp == fp && fp != NULL) { g_files[i].busy = 0; return 0; }
    return bgzf_close(fp);
}

bam_index_t* indelgpu_bam_index_load(const char* fn)
{
    if (getenv("INDELGPU_NO_BAM_CACHE") != NULL) return bam_index_load(fn);
    cached_bam* c = slot_for(fn, 1);
    if (c == NULL) return bam_index_load(fn);
    if (c->idx == NULL) c->idx = bam_index_load(fn);
    return c->idx;
}

void indelgpu_bam_index_destroy(bam_index_t* idx)
{
    for (int i = 0; i < g_nfiles; i++)
        if (g_files[i].idx == idx && idx != NULL) return;        /* kept for the next variant */
    bam_index_destroy(idx);
}
