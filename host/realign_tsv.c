/* C driver over the batch builder: realigns the candidates of a TSV file against a one-line-per-contig
 * sequence file, entirely through the C ABI (no Python).  Used by tests/test_c_batch.py as the "host
 * side in C" end of the GPU parity tests, and as a usage example of the region-sharded multi-GPU
 * layout of DESIGN.md section 7: one worker thread + one indelgpu_ctx per GPU, fixed-size genomic
 * regions dealt round-robin to the workers, results merged back into input order.
 *
 *   realign_tsv [-k K] [-g G] [-s MAXDEL] [-n ETHR] [-G WORKERS] [-R REGION] contigs.txt candidates.tsv > segments.tsv
 *   contigs.txt     one contig per line (upper-case sequence, as read_reference leaves it)
 *   candidates.tsv  tid <TAB> position <TAB> range1 <TAB> read
 *   output          one line per candidate, input order: status <TAB> rstart <TAB> word,word,...
 *   -G WORKERS      worker threads (default 1); worker w uses device w mod indelgpu_device_count()
 *   -R REGION       region size in bases for the sharding (default 1000000)
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "indelgpu_batch.h"

typedef struct { int32_t tid, pos, rng, len; char* read; } cand;
typedef struct { int32_t status, rstart, nseg; uint32_t* words; } outcome;

typedef struct {
    int device;
    const indelgpu_params* p;
    int ncontigs; char** contigs; const int64_t* lens;
    const cand* c; outcome* out;
    const int32_t* mine; int32_t nmine;          /* indices into c[] / out[], in input order */
    int rc; char err[600];
} worker;

static char* read_line(FILE* f)
{
    size_t cap = 1 << 12, n = 0;
    char* s = (char*)malloc(cap);
    int c;
    while ((c = fgetc(f)) != EOF && c != '\n') {
        if (n + 2 > cap) { cap *= 2; s = (char*)realloc(s, cap); }
        s[n++] = (char)c;
    }
    if (c == EOF && n == 0) { free(s); return NULL; }
    s[n] = '\0';
    return s;
}

static void* work(void* arg)
{
    worker* w = (worker*)arg;
    w->rc = 1;
    indelgpu_ctx* ctx = indelgpu_create(w->device, w->p);
    if (!ctx) { snprintf(w->err, sizeof w->err, "indelgpu_create(%d): %s", w->device, indelgpu_last_error()); return NULL; }
    if (indelgpu_set_reference(ctx, w->ncontigs, (const char* const*)w->contigs, w->lens) != 0) {
        snprintf(w->err, sizeof w->err, "indelgpu_set_reference: %s", indelgpu_last_error());
        indelgpu_destroy(ctx); return NULL;
    }
    igb_batch* b = igb_create(4096, 4096 * 512);
    if (!b) { snprintf(w->err, sizeof w->err, "igb_create failed"); indelgpu_destroy(ctx); return NULL; }
    int32_t next = 0;
    while (next < w->nmine) {
        igb_clear(b);
        const int32_t first = next;
        while (next < w->nmine && b->n < b->cap_reads && b->nbases + 512 <= b->cap_bases) {
            const cand* c = &w->c[w->mine[next]];
            if (igb_push(b, c->read, c->len, c->tid, c->pos, c->rng) != 0) {
                snprintf(w->err, sizeof w->err, "read %d too long for the batch", (int)w->mine[next]); goto done;
            }
            next++;
        }
        if (igb_run(b, ctx) != 0) { snprintf(w->err, sizeof w->err, "igb_run: %s", indelgpu_last_error()); goto done; }
        for (int32_t i = 0; i < b->n; i++) {
            outcome* o = &w->out[w->mine[first + i]];
            const uint32_t* words = igb_segments(b, i, &o->nseg, &o->rstart);
            o->status = b->status[i];
            o->words = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(o->nseg > 0 ? o->nseg : 1));
            memcpy(o->words, words, sizeof(uint32_t) * (size_t)o->nseg);
        }
    }
    w->rc = 0;
done:
    igb_destroy(b);
    indelgpu_destroy(ctx);
    return NULL;
}

int main(int argc, char** argv)
{
    indelgpu_params p;
    indelgpu_default_params(&p);
    int a = 1, nworkers = 1;
    long region = 1000000;
    for (; a + 1 < argc && argv[a][0] == '-'; a += 2) {
        const long v = atol(argv[a + 1]);
        switch (argv[a][1]) {
            case 'k': p.klength = (int)v; break;
            case 'g': p.numgaps = (int)v; break;
            case 's': p.maxdelsize = (int)v; break;
            case 'n': p.ethreshold = (int)v; break;
            case 'G': nworkers = (int)v; break;
            case 'R': region = v; break;
            default: fprintf(stderr, "unknown flag %s\n", argv[a]); return 2;
        }
    }
    if (argc - a != 2 || nworkers < 1 || nworkers > 64 || region < 1) {
        fprintf(stderr, "usage: realign_tsv [-k K] [-g G] [-s MAXDEL] [-n ETHR] [-G WORKERS] [-R REGION] contigs.txt candidates.tsv\n");
        return 2;
    }
    const int ndev = indelgpu_device_count();
    if (ndev < 1) { fprintf(stderr, "no CUDA device visible; libindelgpu has no CPU path\n"); return 1; }

    FILE* fc = fopen(argv[a], "r");
    if (!fc) { perror(argv[a]); return 1; }
    int ncontigs = 0, capc = 16;
    char** contigs = (char**)malloc(sizeof(char*) * capc);
    int64_t* lens = (int64_t*)malloc(sizeof(int64_t) * capc);
    for (char* s; (s = read_line(fc)) != NULL;) {
        if (ncontigs == capc) { capc *= 2; contigs = (char**)realloc(contigs, sizeof(char*) * capc); lens = (int64_t*)realloc(lens, sizeof(int64_t) * capc); }
        contigs[ncontigs] = s; lens[ncontigs] = (int64_t)strlen(s); ncontigs++;
    }
    fclose(fc);
    /* regions are numbered contig after contig, so that neighbouring regions go to different workers */
    int64_t* region_base = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncontigs + 1));
    region_base[0] = 0;
    for (int t = 0; t < ncontigs; t++) region_base[t + 1] = region_base[t] + (lens[t] + region - 1) / region;

    FILE* ft = fopen(argv[a + 1], "r");
    if (!ft) { perror(argv[a + 1]); return 1; }
    int32_t n = 0, cap = 1 << 12;
    cand* c = (cand*)malloc(sizeof(cand) * (size_t)cap);
    for (char* s; (s = read_line(ft)) != NULL;) {
        int tid, pos, rng, used = 0;
        if (sscanf(s, "%d\t%d\t%d\t%n", &tid, &pos, &rng, &used) != 3) { fprintf(stderr, "bad line: %s\n", s); return 1; }
        if (n == cap) { cap *= 2; c = (cand*)realloc(c, sizeof(cand) * (size_t)cap); }
        c[n].tid = tid; c[n].pos = pos; c[n].rng = rng;
        c[n].read = s + used; c[n].len = (int32_t)strlen(s + used);   /* the line stays allocated */
        n++;
    }
    fclose(ft);

    outcome* out = (outcome*)calloc((size_t)(n > 0 ? n : 1), sizeof(outcome));
    int32_t* owner_list = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t* count = (int32_t*)calloc((size_t)nworkers + 1, sizeof(int32_t));
    int32_t* owner = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; i++) {
        int64_t r = 0;
        if (c[i].tid >= 0 && c[i].tid < ncontigs && c[i].pos >= 0) r = region_base[c[i].tid] + c[i].pos / region;
        owner[i] = (int32_t)(r % nworkers);
        count[owner[i] + 1]++;
    }
    for (int w = 0; w < nworkers; w++) count[w + 1] += count[w];
    {
        int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)nworkers);
        for (int w = 0; w < nworkers; w++) fill[w] = count[w];
        for (int32_t i = 0; i < n; i++) owner_list[fill[owner[i]]++] = i;
        free(fill);
    }

    worker* ws = (worker*)calloc((size_t)nworkers, sizeof(worker));
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nworkers);
    for (int w = 0; w < nworkers; w++) {
        ws[w].device = w % ndev; ws[w].p = &p;
        ws[w].ncontigs = ncontigs; ws[w].contigs = contigs; ws[w].lens = lens;
        ws[w].c = c; ws[w].out = out;
        ws[w].mine = owner_list + count[w]; ws[w].nmine = count[w + 1] - count[w];
        if (pthread_create(&th[w], NULL, work, &ws[w]) != 0) { fprintf(stderr, "pthread_create failed\n"); return 1; }
    }
    int bad = 0;
    for (int w = 0; w < nworkers; w++) {
        pthread_join(th[w], NULL);
        if (ws[w].rc != 0) { fprintf(stderr, "worker %d (device %d): %s\n", w, ws[w].device, ws[w].err); bad = 1; }
    }
    if (bad) return 1;
    for (int32_t i = 0; i < n; i++) {
        printf("%d\t%d\t", out[i].status, out[i].rstart);
        for (int t = 0; t < out[i].nseg; t++) printf(t ? ",%u" : "%u", out[i].words[t]);
        putchar('\n');
    }
    if (nworkers > 1) {
        fprintf(stderr, "realign_tsv: %d candidates,", (int)n);
        for (int w = 0; w < nworkers; w++) fprintf(stderr, " worker %d (device %d): %d", w, ws[w].device, (int)ws[w].nmine);
        fputc('\n', stderr);
    }
    return 0;
}
