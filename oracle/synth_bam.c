/* TEST / BENCH INFRASTRUCTURE ONLY.  Synthetic paired-end BAM for end-to-end runs of the indelMINER
 * program at the sizes BASELINE.json names (D2 = config 3, D3 = config 4, D4 = config 5),
 * written directly through the reference's bundled samtools-0.1.19 API (libbam.a, compiled where it lies
 * by oracle/Makefile).  The C twin of tests/synth_bam.py: that one is fine at 400 kb, this one writes the
 * 64 Mb / 30x data set of config 3 (12.8 M records) in a couple of minutes.
 *
 * Per contig: uniform ACGT reference; homozygous 1..max_indel bp insertions and deletions planted every
 * ~spacing bases (50/50); 2 x read_len pairs drawn from the mutated genome, insert ~ N(mean, sd)
 * (Irwin-Hall, integer arithmetic only), sub_rate substitutions.  There is no read aligner in the image,
 * so every read gets the record an aligner could have produced: reads that do not touch an indel are plain
 * <read_len>M; reads across an indel get, with probabilities 0.4 / 0.4 / 0.2, the true I/D CIGAR, a soft clip
 * at the indel (longer side kept), or flag 0x4 with the mate mapped.  MAPQ 60, MQ:C:60, no RG tag (read
 * group "generic", indelminer.c:370).  Records are written in coordinate order with a total order on
 * (position, pair, end), and every random draw comes from a counter-based generator keyed by
 * (seed, contig, pair), so the file is a pure function of the arguments: the build container and the GPU
 * box produce the same bytes, which is what lets a VCF computed by the reference in one place be compared
 * with a GPU run in the other (tests/golden/cfg3_reference.json).
 *
 *   synth_bam PREFIX --contigs N --length L [--lengths l1,l2,...] --depth D --readlen M --insert MEAN,SD
 *             --spacing S --maxindel K --subrate R --seed X [--keep F] [--readseed Y] [--level Z] [--rg N]
 * --bigdel K (default 0) turns K of every contig's sites into 1.5-3 kb deletions; pairs that span one come out with an
 * insert beyond the proper range and are flagged improper (0x2 clear): the reference's paired-end evidence path.
 * --rg N (default 0 = no RG tag) deals the pairs to N read groups rg0 .. rgN-1 (RG:Z tag, @RG header lines) and gives
 * every group its own IL line in the config (the maximum grows by 10 per group): exercises the per-read-group range.
 * writes PREFIX.fa, PREFIX.bam, PREFIX.bam.bai, PREFIX.config and prints one JSON line of counts.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bam.h"

static uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
typedef struct { uint64_t s; } rng_t;
static uint64_t rnd(rng_t* r) { r->s += 0x9E3779B97F4A7C15ULL; return mix64(r->s); }
static uint32_t rnd_below(rng_t* r, uint32_t n) { return (uint32_t)(((rnd(r) >> 32) * (uint64_t)n) >> 32); }
static int rnd_prob(rng_t* r, uint32_t per_million) { return rnd_below(r, 1000000u) < per_million; }

typedef struct { int64_t pos; int len; int isdel; int keep; int64_t ins_off; } site_t;
typedef struct { int n; int len[8]; char op[8]; } ops_t;            /* M I D S */
typedef struct { int unmapped; int64_t pos0; ops_t ops; int rev; char bases[1024]; } end_t;

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

typedef struct {
    int64_t L, S;                 /* reference / sample length */
    char* ref; char* samp; int32_t* rcmap;        /* rcmap[s] = reference coordinate of sample base s, -1 = inserted */
    int M; int imean, isd; uint32_t sub_ppm;
    uint64_t seed; int contig;
} genome_t;

static void ops_push(ops_t* o, int n, char op)
{
    if (n <= 0) return;
    if (o->n > 0 && o->op[o->n - 1] == op) { o->len[o->n - 1] += n; return; }
    if (o->n == 8) { fprintf(stderr, "synth_bam: CIGAR too long\n"); exit(1); }
    o->len[o->n] = n; o->op[o->n] = op; o->n++;
}

/* the record an aligner could have produced for sample[s0 .. s0+M) */
static void make_end(const genome_t* g, rng_t* r, int64_t s0, int rev, end_t* e)
{
    const int M = g->M;
    e->rev = rev; e->unmapped = 0; e->ops.n = 0;
    for (int i = 0; i < M; i++) {
        char c = g->samp[s0 + i];
        if (rnd_prob(r, g->sub_ppm)) c = ACGT[rnd_below(r, 4)];
        e->bases[i] = c;
    }
    const int32_t first = g->rcmap[s0], last = g->rcmap[s0 + M - 1];
    if (first >= 0 && last - first == M - 1) { e->pos0 = first; ops_push(&e->ops, M, 'M'); return; }
    int64_t prev = -1; int any = 0;
    for (int i = 0; i < M; i++) {
        const int32_t rc = g->rcmap[s0 + i];
        if (rc < 0) ops_push(&e->ops, 1, 'I');
        else {
            if (!any) { e->pos0 = rc; any = 1; }
            else if (rc > prev + 1) ops_push(&e->ops, (int)(rc - prev - 1), 'D');
            ops_push(&e->ops, 1, 'M');
            prev = rc;
        }
    }
    if (!any) { e->unmapped = 1; return; }
    ops_t* o = &e->ops;
    if (o->op[0] == 'I') o->op[0] = 'S';                              /* inserted bases at a read end are soft-clipped */
    if (o->op[o->n - 1] == 'I') o->op[o->n - 1] = 'S';
    int i0 = -1;
    for (int i = 0; i < o->n; i++) if (o->op[i] == 'I' || o->op[i] == 'D') { i0 = i; break; }
    if (i0 < 0) return;
    const uint32_t u = rnd_below(r, 10);
    if (u < 4) return;                                                /* the aligner found the indel */
    if (u >= 8) { e->unmapped = 1; return; }
    int left = 0, right = 0;                                          /* soft clip at the indel, keeping the longer side */
    for (int i = 0; i < i0; i++) if (o->op[i] != 'D') left += o->len[i];
    for (int i = i0 + 1; i < o->n; i++) if (o->op[i] != 'D') right += o->len[i];
    if (o->op[i0] == 'I') right += o->len[i0];
    ops_t n; n.n = 0;
    if (left >= right) {
        for (int i = 0; i < i0; i++) ops_push(&n, o->len[i], o->op[i]);
        ops_push(&n, M - left, 'S');
    } else {
        int consumed_ref = 0, kept = 0;
        for (int i = 0; i <= i0; i++) if (o->op[i] == 'M' || o->op[i] == 'D') consumed_ref += o->len[i];
        for (int i = i0 + 1; i < o->n; i++) if (o->op[i] != 'D') kept += o->len[i];
        e->pos0 += consumed_ref;
        ops_push(&n, M - kept, 'S');
        for (int i = i0 + 1; i < o->n; i++) ops_push(&n, o->len[i], o->op[i]);
    }
    *o = n;
}

static int64_t ref_end(const end_t* e)
{
    int64_t x = e->pos0;
    for (int i = 0; i < e->ops.n; i++) if (e->ops.op[i] == 'M' || e->ops.op[i] == 'D') x += e->ops.len[i];
    return x;
}

/* both ends of pair k; returns 0 when the pair is dropped */
static int make_pair(const genome_t* g, uint64_t k, end_t* a, end_t* b, int* first_left)
{
    rng_t r; r.s = mix64(g->seed ^ mix64(((uint64_t)g->contig << 40) ^ k));
    const int M = g->M;
    const int64_t fstart = (int64_t)(rnd(&r) % (uint64_t)(g->S - 1000));
    int64_t acc = 0;
    for (int t = 0; t < 12; t++) acc += (int64_t)rnd_below(&r, 1 << 20);          /* Irwin-Hall: mean 6 * 2^20, sd 2^20 */
    int64_t isz = g->imean + ((acc - 6 * (1LL << 20)) * g->isd) / (1LL << 20);
    if (isz < 2 * M + 10) isz = 2 * M + 10;
    if (isz > g->imean + 4 * g->isd) isz = g->imean + 4 * g->isd;
    *first_left = (int)(rnd(&r) & 1);
    if (fstart + isz >= g->S) return 0;
    make_end(g, &r, fstart, 0, a);
    make_end(g, &r, fstart + isz - M, 1, b);
    if (a->unmapped && b->unmapped) return 0;
    return 1;
}

typedef struct { int64_t pos; uint64_t pair; int which; } skey_t;
static int key_cmp(const void* x, const void* y)
{
    const skey_t* a = x; const skey_t* b = y;
    if (a->pos != b->pos) return a->pos < b->pos ? -1 : 1;
    if (a->pair != b->pair) return a->pair < b->pair ? -1 : 1;
    return a->which - b->which;                                        /* total order: the sort algorithm cannot matter */
}

static int op_code(char c) { return c == 'M' ? BAM_CMATCH : c == 'I' ? BAM_CINS : c == 'D' ? BAM_CDEL : BAM_CSOFT_CLIP; }

static int g_nrg = 0;
static int g_proper_max = 0;           /* > 0: pairs with a longer template are flagged improper (--bigdel) */

static void write_record(bamFile out, int tid, uint64_t pair, const end_t* e, const end_t* mate, int is_first, int M)
{
    static uint8_t data[4096];
    bam1_t b; memset(&b, 0, sizeof(b));
    char name[32];
    const int lq = snprintf(name, sizeof(name), "p%d_%llu", tid, (unsigned long long)pair) + 1;
    uint32_t flag = 0x1 | (is_first ? 0x40 : 0x80);
    if (e->rev) flag |= 0x10;
    if (mate->rev) flag |= 0x20;
    if (e->unmapped) flag |= 0x4;
    if (mate->unmapped) flag |= 0x8;
    const int64_t pos = e->unmapped ? mate->pos0 : e->pos0;
    const int64_t pnext = mate->unmapped ? pos : mate->pos0;
    int64_t tlen = 0, end = pos + 1;
    if (!e->unmapped) end = ref_end(e);
    if (!e->unmapped && !mate->unmapped) {
        const int64_t ee = ref_end(e), me = ref_end(mate);
        const int64_t lo = e->pos0 < mate->pos0 ? e->pos0 : mate->pos0, hi = ee > me ? ee : me;
        tlen = e->pos0 <= mate->pos0 ? hi - lo : -(hi - lo);
        if (g_proper_max <= 0 || hi - lo <= g_proper_max) flag |= 0x2;
    }
    uint8_t* p = data;
    memcpy(p, name, (size_t)lq); p += lq;
    const int nc = e->unmapped ? 0 : e->ops.n;
    for (int i = 0; i < nc; i++) { const uint32_t c = ((uint32_t)e->ops.len[i] << BAM_CIGAR_SHIFT) | (uint32_t)op_code(e->ops.op[i]); memcpy(p, &c, 4); p += 4; }
    memset(p, 0, (size_t)(M + 1) / 2);
    for (int i = 0; i < M; i++) p[i >> 1] |= (uint8_t)(bam_nt16_table[(int)e->bases[i]] << ((~i & 1) << 2));
    p += (M + 1) / 2;
    memset(p, 40, (size_t)M); p += M;                                  /* quality 'I' */
    p[0] = 'M'; p[1] = 'Q'; p[2] = 'C'; p[3] = 60; p += 4;
    int l_aux = 4;
    if (g_nrg > 0) {
        const int n = snprintf((char*)p, 16, "RGZrg%d", (int)(pair % (uint64_t)g_nrg)) + 1;      /* tag, type, NUL-terminated name */
        p += n; l_aux += n;
    }
    b.core.tid = tid; b.core.pos = (int32_t)pos; b.core.bin = bam_reg2bin((uint32_t)pos, (uint32_t)end);
    b.core.qual = e->unmapped ? 0 : 60; b.core.l_qname = (uint8_t)lq; b.core.flag = flag; b.core.n_cigar = (uint16_t)nc;
    b.core.l_qseq = M; b.core.mtid = tid; b.core.mpos = (int32_t)pnext; b.core.isize = (int32_t)tlen;
    b.data = data; b.data_len = (int)(p - data); b.m_data = (int)sizeof(data); b.l_aux = l_aux;
    bam_write1(out, &b);
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: synth_bam PREFIX [options]; see the head of oracle/synth_bam.c\n"); return 2; }
    const char* prefix = argv[1];
    int ncontigs = 1; int64_t length = 1000000; const char* lengths = NULL;
    double depth = 20; int M = 150, imean = 500, isd = 50, spacing = 2000, maxindel = 50, level = 1;
    double subrate = 0.01, keep = 1.0; uint64_t seed = 7, readseed = 0; int have_readseed = 0; int bigdel = 0;
    for (int i = 2; i + 1 < argc; i += 2) {
        const char* o = argv[i]; const char* v = argv[i + 1];
        if (!strcmp(o, "--contigs")) ncontigs = atoi(v);
        else if (!strcmp(o, "--length")) length = atoll(v);
        else if (!strcmp(o, "--lengths")) lengths = v;
        else if (!strcmp(o, "--depth")) depth = atof(v);
        else if (!strcmp(o, "--readlen")) M = atoi(v);
        else if (!strcmp(o, "--insert")) { if (sscanf(v, "%d,%d", &imean, &isd) != 2) return 2; }
        else if (!strcmp(o, "--spacing")) spacing = atoi(v);
        else if (!strcmp(o, "--maxindel")) maxindel = atoi(v);
        else if (!strcmp(o, "--subrate")) subrate = atof(v);
        else if (!strcmp(o, "--seed")) seed = strtoull(v, NULL, 10);
        else if (!strcmp(o, "--keep")) keep = atof(v);
        else if (!strcmp(o, "--readseed")) { readseed = strtoull(v, NULL, 10); have_readseed = 1; }
        else if (!strcmp(o, "--level")) level = atoi(v);
        else if (!strcmp(o, "--rg")) g_nrg = atoi(v);
        else if (!strcmp(o, "--bigdel")) bigdel = atoi(v);
        else { fprintf(stderr, "synth_bam: unknown option %s\n", o); return 2; }
    }
    if (M > 1000 || M < 20 || ncontigs < 1 || ncontigs > 512) return 2;
    if (bigdel > 0) g_proper_max = imean + 4 * isd;
    int64_t* clen = calloc((size_t)ncontigs, sizeof(int64_t));
    for (int c = 0; c < ncontigs; c++) clen[c] = length;
    if (lengths) { const char* p = lengths; for (int c = 0; c < ncontigs && p; c++) { clen[c] = atoll(p); p = strchr(p, ','); if (p) p++; } }

    char path[4096], mode[8];
    snprintf(path, sizeof(path), "%s.fa", prefix);
    FILE* fa = fopen(path, "w");
    snprintf(path, sizeof(path), "%s.config", prefix);
    FILE* cfg = fopen(path, "w");
    if (!fa || !cfg) { fprintf(stderr, "synth_bam: cannot write %s.*\n", prefix); return 1; }
    if (g_nrg <= 0) fprintf(cfg, "IL generic %d %d\n", imean - 6 * isd, imean + 4 * isd);
    for (int r = 0; r < g_nrg; r++) fprintf(cfg, "IL rg%d %d %d\n", r, imean - 6 * isd, imean + 4 * isd + 10 * r);

    bam_header_t* h = bam_header_init();
    h->n_targets = ncontigs;
    h->target_name = calloc((size_t)ncontigs, sizeof(char*));
    h->target_len = calloc((size_t)ncontigs, sizeof(uint32_t));
    char* text = malloc(64 + 64 * (size_t)ncontigs + 32 * (size_t)(g_nrg > 0 ? g_nrg : 0));
    int tl = sprintf(text, "@HD\tVN:1.0\tSO:coordinate\n");
    for (int c = 0; c < ncontigs; c++) {
        char nm[32]; snprintf(nm, sizeof(nm), "chr%d", c + 1);
        h->target_name[c] = strdup(nm); h->target_len[c] = (uint32_t)clen[c];
        tl += sprintf(text + tl, "@SQ\tSN:%s\tLN:%lld\n", nm, (long long)clen[c]);
    }
    for (int r = 0; r < g_nrg; r++) tl += sprintf(text + tl, "@RG\tID:rg%d\tSM:s\n", r);
    h->text = text; h->l_text = (uint32_t)tl;
    snprintf(path, sizeof(path), "%s.bam", prefix);
    snprintf(mode, sizeof(mode), "w%d", level);
    bamFile out = bam_open(path, mode);
    if (!out) { fprintf(stderr, "synth_bam: cannot write %s\n", path); return 1; }
    bam_header_write(out, h);

    long long nrec = 0, npairs_all = 0, nsites_all = 0, ndel = 0, nins = 0, ncand = 0;
    for (int c = 0; c < ncontigs; c++) {
        const int64_t L = clen[c];
        genome_t g; memset(&g, 0, sizeof(g));
        g.L = L; g.M = M; g.imean = imean; g.isd = isd; g.sub_ppm = (uint32_t)(subrate * 1e6 + 0.5);
        g.contig = c;
        g.ref = malloc((size_t)L + 1);
        rng_t rr; rr.s = mix64(seed ^ (0xABCD0000ULL + (uint64_t)c));
        for (int64_t i = 0; i < L; i += 32) {                           /* 32 bases per draw */
            uint64_t w = rnd(&rr);
            for (int t = 0; t < 32 && i + t < L; t++, w >>= 2) g.ref[i + t] = ACGT[w & 3];
        }
        g.ref[L] = '\0';
        fprintf(fa, ">chr%d\n", c + 1);
        for (int64_t i = 0; i < L; i += 60) { const int n = (int)(L - i < 60 ? L - i : 60); fwrite(g.ref + i, 1, (size_t)n, fa); fputc('\n', fa); }
        fprintf(cfg, "RC chr%d %d\n", c + 1, (int)depth);

        /* sites (drawn for every site so that subsets agree), then the mutated genome */
        const int64_t nsites = L > 4000 ? (L - 4000) / spacing : 0;
        site_t* sites = calloc((size_t)nsites + 1, sizeof(site_t));
        rng_t rs; rs.s = mix64(seed ^ (0x5117E000ULL + (uint64_t)c));
        rng_t rk; rk.s = mix64((seed + 1000) ^ (0x5117E000ULL + (uint64_t)c));
        int64_t instotal = 0;
        for (int64_t i = 0; i < nsites; i++) {
            sites[i].pos = 2000 + i * spacing + rnd_below(&rs, (uint32_t)(spacing / 2));
            sites[i].len = 1 + (int)rnd_below(&rs, (uint32_t)maxindel);
            sites[i].isdel = (int)(rnd(&rs) & 1);
            sites[i].ins_off = instotal; instotal += sites[i].len;
            sites[i].keep = rnd_below(&rk, 1000000u) < (uint32_t)(keep * 1e6 + 0.5);
        }
        if (bigdel > 0 && nsites > 2 * bigdel) {                        /* its own draws: the default data set does not change */
            rng_t rb; rb.s = mix64(seed ^ (0xB16DE100ULL + (uint64_t)c));
            const int64_t every = nsites / bigdel;
            for (int64_t i = every / 2; i < nsites; i += every) { sites[i].isdel = 1; sites[i].len = 1500 + (int)rnd_below(&rb, 1500); }
        }
        char* insbases = malloc((size_t)instotal + 1);
        for (int64_t i = 0; i < instotal; i++) insbases[i] = ACGT[rnd_below(&rs, 4)];
        g.samp = malloc((size_t)L + (size_t)instotal + 16);
        g.rcmap = malloc(sizeof(int32_t) * ((size_t)L + (size_t)instotal + 16));
        int64_t S = 0, prev = 0;
        for (int64_t i = 0; i < nsites; i++) {
            if (!sites[i].keep || sites[i].pos < prev) continue;           /* a site inside a big deletion is gone */
            for (int64_t p = prev; p < sites[i].pos; p++) { g.samp[S] = g.ref[p]; g.rcmap[S++] = (int32_t)p; }
            if (sites[i].isdel) { prev = sites[i].pos + sites[i].len; ndel++; }
            else { for (int t = 0; t < sites[i].len; t++) { g.samp[S] = insbases[sites[i].ins_off + t]; g.rcmap[S++] = -1; } prev = sites[i].pos; nins++; }
            nsites_all++;
        }
        for (int64_t p = prev; p < L; p++) { g.samp[S] = g.ref[p]; g.rcmap[S++] = (int32_t)p; }
        g.S = S;
        g.seed = have_readseed ? mix64(readseed) : mix64(seed ^ 0x7EAD5ULL);

        const uint64_t npairs = (uint64_t)(depth * (double)L / (2.0 * M));
        skey_t* keys = malloc(sizeof(skey_t) * (size_t)(2 * npairs + 2));
        size_t nk = 0;
        end_t a, b; int fl;
        for (uint64_t k = 0; k < npairs; k++) {
            if (!make_pair(&g, k, &a, &b, &fl)) continue;
            npairs_all++;
            keys[nk].pos = a.unmapped ? b.pos0 : a.pos0; keys[nk].pair = k; keys[nk].which = 0; nk++;
            keys[nk].pos = b.unmapped ? a.pos0 : b.pos0; keys[nk].pair = k; keys[nk].which = 1; nk++;
        }
        qsort(keys, nk, sizeof(skey_t), key_cmp);
        for (size_t i = 0; i < nk; i++) {
            make_pair(&g, keys[i].pair, &a, &b, &fl);
            const end_t* e = keys[i].which ? &b : &a; const end_t* m = keys[i].which ? &a : &b;
            write_record(out, c, keys[i].pair, e, m, (keys[i].which == 0) == (fl != 0), M);
            if (e->unmapped || e->ops.n > 1) ncand++;
            nrec++;
        }
        free(keys); free(g.ref); free(g.samp); free(g.rcmap); free(sites); free(insbases);
    }
    bam_close(out);
    fclose(fa); fclose(cfg);
    snprintf(path, sizeof(path), "%s.bam", prefix);
    if (bam_index_build(path) != 0) { fprintf(stderr, "synth_bam: indexing %s failed\n", path); return 1; }
    printf("{\"contigs\": %d, \"records\": %lld, \"pairs\": %lld, \"sites\": %lld, \"deletions\": %lld, \"insertions\": %lld, \"non_plain_records\": %lld}\n",
           ncontigs, nrec, npairs_all, nsites_all, ndel, nins, ncand);
    return 0;
}
