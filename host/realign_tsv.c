/* Small C driver over the batch builder: realigns the candidates of a TSV file against a one-line-
 * per-contig sequence file, entirely through the C ABI (no Python).  Used by tests/test_c_batch.py
 * as the "host side in C" end of the GPU parity tests, and as a usage example.
 *
 *   realign_tsv [-k K] [-g G] [-s MAXDEL] [-n ETHR] contigs.txt candidates.tsv > segments.tsv
 *   contigs.txt     one contig per line (upper-case sequence, as read_reference leaves it)
 *   candidates.tsv  tid <TAB> position <TAB> range1 <TAB> read
 *   output          one line per candidate: status <TAB> rstart <TAB> word,word,...
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "indelgpu_batch.h"

static char* read_line(FILE* f)
{
    size_t cap = 1 << 16, n = 0;
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

int main(int argc, char** argv)
{
    indelgpu_params p;
    indelgpu_default_params(&p);
    int a = 1;
    for (; a + 1 < argc && argv[a][0] == '-'; a += 2) {
        const int v = atoi(argv[a + 1]);
        switch (argv[a][1]) {
            case 'k': p.klength = v; break;
            case 'g': p.numgaps = v; break;
            case 's': p.maxdelsize = v; break;
            case 'n': p.ethreshold = v; break;
            default: fprintf(stderr, "unknown flag %s\n", argv[a]); return 2;
        }
    }
    if (argc - a != 2) { fprintf(stderr, "usage: realign_tsv [-k K] [-g G] [-s MAXDEL] [-n ETHR] contigs.txt candidates.tsv\n"); return 2; }

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

    indelgpu_ctx* ctx = indelgpu_create(0, &p);
    if (!ctx) { fprintf(stderr, "indelgpu_create: %s\n", indelgpu_last_error()); return 1; }
    if (indelgpu_set_reference(ctx, ncontigs, (const char* const*)contigs, lens) != 0) {
        fprintf(stderr, "indelgpu_set_reference: %s\n", indelgpu_last_error()); return 1;
    }

    igb_batch* b = igb_create(4096, 4096 * 512);
    if (!b) { fprintf(stderr, "igb_create failed\n"); return 1; }
    FILE* ft = fopen(argv[a + 1], "r");
    if (!ft) { perror(argv[a + 1]); return 1; }
    int eof = 0;
    while (!eof) {
        igb_clear(b);
        for (;;) {                                              /* fill one batch */
            char* s = read_line(ft);
            if (!s) { eof = 1; break; }
            int tid, pos, rng, used = 0;
            if (sscanf(s, "%d\t%d\t%d\t%n", &tid, &pos, &rng, &used) != 3) { fprintf(stderr, "bad line: %s\n", s); return 1; }
            const char* read = s + used;
            if (igb_push(b, read, (int32_t)strlen(read), tid, pos, rng) != 0) { fprintf(stderr, "read too long for the batch\n"); return 1; }
            free(s);
            if (b->n == b->cap_reads || b->nbases + 512 > b->cap_bases) break;
        }
        if (b->n == 0) break;
        if (igb_run(b, ctx) != 0) { fprintf(stderr, "igb_run: %s\n", indelgpu_last_error()); return 1; }
        for (int i = 0; i < b->n; i++) {
            int32_t ns, rs;
            const uint32_t* w = igb_segments(b, i, &ns, &rs);
            printf("%d\t%d\t", b->status[i], rs);
            for (int t = 0; t < ns; t++) printf(t ? ",%u" : "%u", w[t]);
            putchar('\n');
        }
    }
    fclose(ft);
    igb_destroy(b);
    indelgpu_destroy(ctx);
    return 0;
}
