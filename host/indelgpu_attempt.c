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
 * Both process one read per call (a 1-element batch) unless INDELGPU_MODE selects the batched
 * record / replay operation described further down (and in INTEGRATION.md).
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

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

/* segment words -> rln->segments -> evidence list: what the reference does once its alignments are known */
static evidence* consume_segments(readaln* const rln, const char* read, const int32_t nseg, const int32_t rstart,
                                  const uint32_t* words)
{
    if (nseg == 0) return NULL;                          /* every NULL exit of alignment.c:568-751 */

    /* update_readsegs' list construction (readaln.c:355-457) from the stitched words */
    int refindx = rstart, readindx = 0;
    readseg* readsegs = NULL;
    for (int i = 0; i < nseg; i++) {
        readseg* rsg = new_readseg(read, words[i], &refindx, &readindx);
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

/* ---- batched operation with an UNCHANGED caller: record / replay ------------------------------------
 * attempt_pe_alignment is a pure function and fetch_func decides to call it from the BAM record alone
 * (flags, CIGAR, mate quality; indelminer.c:384-492), never from earlier alignment results, so two runs
 * of the same command make the same calls in the same order:
 *   INDELGPU_MODE=record  every call is queued and answered NULL (the VCF of this run is discarded); at
 *                         exit the queue is realigned contig by contig with ONE indelgpu_realign_batch
 *                         each and the results are written to $INDELGPU_REPLAY_FILE;
 *   INDELGPU_MODE=replay  every call is answered from that file, in order.
 * The second run prints the VCF; the GPU sees whole-contig batches instead of one read at a time.
 *   INDELGPU_MODE=auto    both in ONE run: at the first call the process forks; the child is the recording
 *                         run (stdout discarded), the parent waits for it and continues as the replay run.
 *                         Parent and child share the offsets of the open files (the BAM), so the parent
 *                         puts every regular file's offset back where it was before it goes on. */
enum { MODE_DIRECT = 0, MODE_RECORD = 1, MODE_REPLAY = 2 };
static int g_mode = -1;

typedef struct { int32_t tid, position, range1, readlen; int64_t base_off; } cand;
static cand* g_cands = NULL;  static int64_t g_ncands = 0, g_capcands = 0;
static char* g_bases = NULL;  static int64_t g_nbases = 0, g_capbases = 0;

static const char* replay_path(void)
{
    const char* p = getenv("INDELGPU_REPLAY_FILE");
    return p ? p : "indelgpu_replay.bin";
}

static void flush_recorded(void)
{
    FILE* f = fopen(replay_path(), "wb");
    if (f == NULL) fatalf("libindelgpu: cannot write %s", replay_path());
    const int64_t magic = 0x31594C5052474449LL;          /* "IDGRPLY1" */
    fwrite(&magic, 8, 1, f); fwrite(&g_ncands, 8, 1, f);
    /* results in call order */
    int32_t* nseg = ckallocz(sizeof(int32_t) * (size_t)(g_ncands + 1));
    int32_t* rstart = ckallocz(sizeof(int32_t) * (size_t)(g_ncands + 1));
    uint32_t** words = ckallocz(sizeof(uint32_t*) * (size_t)(g_ncands + 1));
    for (int t = 0; t < g_nslots || (g_all != NULL && t == 0); t++) {
        /* the candidates of contig t (or of every contig when one context holds them all), in call order */
        int64_t n = 0, nb = 0;
        for (int64_t i = 0; i < g_ncands; i++)
            if (g_all != NULL || g_cands[i].tid == t) { n++; nb += g_cands[i].readlen; }
        if (n == 0) { if (g_all != NULL) break; continue; }
        indelgpu_ctx* ctx = g_all != NULL ? g_all : g_slots[t].ctx;
        uint8_t* bases = ckalloc((size_t)nb + 16);
        int64_t* off = ckalloc(sizeof(int64_t) * (size_t)(n + 1));
        int32_t* tid = ckalloc(sizeof(int32_t) * (size_t)n);
        int32_t* pos = ckalloc(sizeof(int32_t) * (size_t)n);
        int32_t* rng = ckalloc(sizeof(int32_t) * (size_t)n);
        int64_t* which = ckalloc(sizeof(int64_t) * (size_t)n);
        int64_t k = 0, b = 0;
        for (int64_t i = 0; i < g_ncands; i++) {
            if (!(g_all != NULL || g_cands[i].tid == t)) continue;
            memcpy(bases + b, g_bases + g_cands[i].base_off, (size_t)g_cands[i].readlen);
            off[k] = b; b += g_cands[i].readlen;
            tid[k] = g_all != NULL ? g_cands[i].tid : 0; pos[k] = g_cands[i].position; rng[k] = g_cands[i].range1;
            which[k] = i; k++;
        }
        off[n] = b;
        const int64_t segcap = indelgpu_seg_bound((int32_t)n, nb);
        int32_t* st = ckalloc(sizeof(int32_t) * (size_t)n);
        int32_t* ns = ckalloc(sizeof(int32_t) * (size_t)n);
        int32_t* rs = ckalloc(sizeof(int32_t) * (size_t)n);
        int64_t* so = ckalloc(sizeof(int64_t) * (size_t)n);
        uint32_t* sg = ckalloc(sizeof(uint32_t) * (size_t)segcap);
        indelgpu_batch in = {(int32_t)n, bases, off, tid, pos, rng};
        indelgpu_result out;
        memset(&out, 0, sizeof(out));
        out.status = st; out.nseg = ns; out.rstart = rs; out.seg_off = so; out.segs = sg; out.seg_capacity = segcap;
        if (indelgpu_realign_batch(ctx, &in, &out) != 0) gpu_die("indelgpu_realign_batch (record mode)");
        for (k = 0; k < n; k++) {
            const int64_t i = which[k];
            nseg[i] = ns[k]; rstart[i] = rs[k];
            if (ns[k] > 0) {
                words[i] = ckalloc(sizeof(uint32_t) * (size_t)ns[k]);
                memcpy(words[i], sg + so[k], sizeof(uint32_t) * (size_t)ns[k]);
            }
        }
        ckfree(bases); ckfree(off); ckfree(tid); ckfree(pos); ckfree(rng); ckfree(which);
        ckfree(st); ckfree(ns); ckfree(rs); ckfree(so); ckfree(sg);
        if (g_all != NULL) break;
    }
    for (int64_t i = 0; i < g_ncands; i++) {
        fwrite(&g_cands[i].position, 4, 1, f); fwrite(&g_cands[i].readlen, 4, 1, f);
        fwrite(&nseg[i], 4, 1, f); fwrite(&rstart[i], 4, 1, f);
        if (nseg[i] > 0) fwrite(words[i], 4, (size_t)nseg[i], f);
    }
    if (fclose(f) != 0) fatalf("libindelgpu: error writing %s", replay_path());
    fprintf(stderr, "libindelgpu: %lld candidate reads realigned in batches, results in %s\n", (long long)g_ncands, replay_path());
}

static char g_auto_path[64] = "";    /* temporary replay file of INDELGPU_MODE=auto */
static int32_t* g_replay = NULL;     /* the whole replay file after its header, as 32-bit words */
static int64_t g_replay_words = 0, g_replay_pos = 0, g_replay_left = 0;
static int g_replay_loaded = 0;

static void load_replay(void)
{
    FILE* f = fopen(replay_path(), "rb");
    if (f == NULL) fatalf("libindelgpu: cannot read %s (run with INDELGPU_MODE=record first)", replay_path());
    int64_t hdr[2];
    if (fread(hdr, 8, 2, f) != 2 || hdr[0] != 0x31594C5052474449LL) fatalf("libindelgpu: %s is not a replay file", replay_path());
    g_replay_left = hdr[1];
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f) - 16;
    fseek(f, 16, SEEK_SET);
    g_replay = ckalloc((size_t)bytes + 4);
    g_replay_words = bytes / 4;
    if (bytes > 0 && fread(g_replay, 1, (size_t)bytes, f) != (size_t)bytes) fatalf("libindelgpu: short read on %s", replay_path());
    fclose(f);
    if (g_auto_path[0] != '\0') unlink(g_auto_path);     /* the temporary file of INDELGPU_MODE=auto */
}

/* INDELGPU_MODE=auto: fork the recording run; returns the mode this process continues in */

static int fork_recording_run(void)
{
    /* offsets of every open regular file: the child will move them */
    enum { MAXFD = 256 };
    int fds[MAXFD]; off_t offs[MAXFD]; int nfd = 0;
    DIR* d = opendir("/proc/self/fd");
    if (d != NULL) {
        for (struct dirent* e; (e = readdir(d)) != NULL && nfd < MAXFD;) {
            const int fd = atoi(e->d_name);
            struct stat st;
            if (e->d_name[0] < '0' || e->d_name[0] > '9' || fd == dirfd(d)) continue;
            if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) continue;
            fds[nfd] = fd; offs[nfd] = lseek(fd, 0, SEEK_CUR); nfd++;
        }
        closedir(d);
    }
    if (getenv("INDELGPU_REPLAY_FILE") == NULL) {
        snprintf(g_auto_path, sizeof(g_auto_path), "/tmp/indelgpu_replay_%d.bin", (int)getpid());
        setenv("INDELGPU_REPLAY_FILE", g_auto_path, 1);
    }
    fflush(stdout); fflush(stderr);
    const pid_t pid = fork();
    if (pid < 0) fatalf("libindelgpu: fork failed");
    if (pid == 0) {                                      /* child: the recording run, its VCF goes nowhere */
        const int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, STDOUT_FILENO); close(nul); }
        return MODE_RECORD;
    }
    int status = 0;
    if (waitpid(pid, &status, 0) != pid || !WIFEXITED(status) || WEXITSTATUS(status) != 0)
        fatalf("libindelgpu: the recording run failed (status %d)", status);
    for (int i = 0; i < nfd; i++) if (offs[i] != (off_t)-1) lseek(fds[i], offs[i], SEEK_SET);
    return MODE_REPLAY;
}

/* The operating mode of this process (0 direct, 1 record, 2 replay), decided once -- INDELGPU_MODE=auto
 * forks here, at the first GPU-bound call of either glue file (host/indelgpu_support.c shares it). */
int indelgpu_glue_mode(void)
{
    if (g_mode < 0) {
        const char* m = getenv("INDELGPU_MODE");
        g_mode = (m && strcmp(m, "record") == 0) ? MODE_RECORD : (m && strcmp(m, "replay") == 0) ? MODE_REPLAY : MODE_DIRECT;
        if (m && strcmp(m, "auto") == 0) g_mode = fork_recording_run();
    }
    return g_mode;
}

/* path of the replay file (INDELGPU_REPLAY_FILE, or the temporary one of auto mode) */
const char* indelgpu_glue_replay_path(void) { return replay_path(); }
/* 1 when this process created the replay file itself (auto mode) and should remove it after loading */
int indelgpu_glue_replay_is_temporary(void) { return g_auto_path[0] != '\0'; }

evidence* attempt_pe_alignment(char** const sequences,
                               const int32_t tid,
                               const int32_t position,
                               const int* const range,
                               readaln* const rln)
{
    forceassert(range[0] <= range[1]);                   /* alignment.c:773 */
    const char* read = rln->segments->sequence;          /* the single S segment (readaln.c:258-264) */
    const int64_t readlength = (int64_t)strlen(read);

    if (g_replay_loaded == 0 && indelgpu_glue_mode() == MODE_REPLAY) { load_replay(); g_replay_loaded = 1; }

    if (g_mode == MODE_REPLAY) {
        if (g_replay_left <= 0 || g_replay_pos + 4 > g_replay_words) fatalf("libindelgpu: replay file exhausted (different command line than the recording run?)");
        const int32_t* r = g_replay + g_replay_pos;
        if (r[0] != position || r[1] != (int32_t)readlength) fatalf("libindelgpu: replay out of step at position %d", position);
        const int32_t nseg = r[2], rstart = r[3];
        g_replay_pos += 4 + nseg; g_replay_left--;
        return consume_segments(rln, read, nseg, rstart, (const uint32_t*)(r + 4));
    }

    int32_t dtid = 0;
    indelgpu_ctx* ctx = ctx_for_contig(sequences, tid, &dtid);   /* uploads the contig the first time it is seen */

    if (g_mode == MODE_RECORD) {
        /* exit handlers run in reverse order of registration: this one must be registered AFTER the CUDA
         * runtime registered its own (first context creation, just above), or the GPU is gone when it runs */
        static int registered = 0;
        if (!registered) { atexit(flush_recorded); registered = 1; }
        if (g_ncands == g_capcands) { g_capcands = g_capcands ? 2 * g_capcands : 4096; g_cands = ckrealloc(g_cands, sizeof(cand) * (size_t)g_capcands); }
        if (g_nbases + readlength > g_capbases) { g_capbases = 2 * (g_capbases + readlength) + 4096; g_bases = ckrealloc(g_bases, (size_t)g_capbases); }
        cand* c = &g_cands[g_ncands++];
        c->tid = tid; c->position = position; c->range1 = range[1]; c->readlen = (int32_t)readlength; c->base_off = g_nbases;
        memcpy(g_bases + g_nbases, read, (size_t)readlength);
        g_nbases += readlength;
        return NULL;                                     /* rln untouched, as on a failed alignment */
    }

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
    return consume_segments(rln, read, nseg, rstart, g_segs + segoff);
}
