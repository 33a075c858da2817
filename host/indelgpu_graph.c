/* Row f4 of SURVEY.md section 8 (the evidence graph): add_node without the scan of every earlier node.
 *
 * process_evidence (src/indelminer.c:117-209) turns every piece of evidence of a chunk into a graph node, and
 * add_node (src/graph.c:82-149) compares each new node with ALL earlier ones to decide where edges go: O(n^2) per
 * chunk -- with the alignments and the per-variant BAM fetches out of the way, a quarter of what is left of the
 * main thread's time.  But an edge can only join
 *   - two SPLIT_READ nodes with the same variant class and the SAME breakpoints (graph.c:120-125), or
 *   - two PAIRED_READ nodes (graph.c:96-119, interval tests);
 * every other pair is skipped (graph.c:126-140).
 *
 * The build compiles graph.c -- unchanged -- with -Dadd_node=indelgpu_reference_add_node and links this file's
 * add_node instead.  It keeps, per graph, the chain of earlier SPLIT_READ nodes with the same (class, b1, b2) and the
 * chain of all PAIRED_READ nodes, most recent first -- the order graph.c's scan visits them in -- and then lets the
 * REFERENCE's own add_node do the work on exactly those nodes: the candidates are linked into a temporary list, the
 * graph's node list is pointed at it for the call (so the reference's tests, its make_edge / edge_exists, its
 * hashtable insertion and its sladdhead run as they are), and the real list is restored around the new node.
 * The edges, their order in every node's edge array and in the graph's edge list are the reference's.  Evidence of
 * any other type sends the whole graph back to the reference's full scan.  Output unchanged (md5-identical VCFs on
 * every config, tests/test_e2e_configs.py); INDELGPU_NO_GRAPH_INDEX=1 turns it off.  Nothing GPU-specific here. */
#include <stdlib.h>
#include <string.h>

#include "graph.h"          /* the reference's header: node, edge, graph, evidence */
#include "asserts.h"
#include "memalloc.h"

void indelgpu_reference_add_node(graph* g, node* const n);      /* graph.c:82-149, renamed by the build */

typedef struct { node* n; int prev; } chain_entry;              /* prev: the earlier entry of the same chain, -1 = none */
typedef struct { int32_t vclass, b1, b2; int last; } key_slot;   /* last: the most recent entry with this key, -1 = empty slot */

static graph* g_graph = NULL;
static chain_entry* g_ent = NULL;  static int g_nent = 0, g_capent = 0;
static key_slot* g_slots = NULL;   static int g_nslots = 0, g_used = 0;
static int g_last_paired = -1;     /* chain of PAIRED_READ nodes */
static int g_plain = 1;            /* 0: an evidence type this file does not know was seen: full scans from then on */
static node** g_tmp = NULL; static node** g_saved = NULL; static int g_captmp = 0;

static void reset_index(graph* g)
{
    g_graph = g; g_nent = 0; g_used = 0; g_last_paired = -1; g_plain = 1;
    if (g_nslots == 0) { g_nslots = 1 << 12; g_slots = ckalloc(sizeof(key_slot) * (size_t)g_nslots); }
    for (int i = 0; i < g_nslots; i++) g_slots[i].last = -1;
}

static unsigned hash_key(int32_t c, int32_t b1, int32_t b2)
{
    unsigned h = (unsigned)c * 0x9E3779B1u ^ ((unsigned)b1 * 0x85EBCA6Bu) ^ ((unsigned)b2 * 0xC2B2AE35u);
    return h ^ (h >> 15);
}

static key_slot* find_slot(int32_t c, int32_t b1, int32_t b2)
{
    unsigned i = hash_key(c, b1, b2) & (unsigned)(g_nslots - 1);
    while (g_slots[i].last >= 0 && !(g_slots[i].vclass == c && g_slots[i].b1 == b1 && g_slots[i].b2 == b2)) i = (i + 1) & (unsigned)(g_nslots - 1);
    return &g_slots[i];
}

static void grow_slots(void)
{
    key_slot* old = g_slots; const int oldn = g_nslots;
    g_nslots *= 2;
    g_slots = ckalloc(sizeof(key_slot) * (size_t)g_nslots);
    for (int i = 0; i < g_nslots; i++) g_slots[i].last = -1;
    for (int i = 0; i < oldn; i++)
        if (old[i].last >= 0) *find_slot(old[i].vclass, old[i].b1, old[i].b2) = old[i];
    ckfree(old);
}

static int push_entry(node* n, int prev)
{
    if (g_nent == g_capent) { g_capent = g_capent ? 2 * g_capent : 4096; g_ent = ckrealloc(g_ent, sizeof(chain_entry) * (size_t)g_capent); }
    g_ent[g_nent].n = n; g_ent[g_nent].prev = prev;
    return g_nent++;
}

void add_node(graph* g, node* const n)
{
    static int disabled = -1;
    if (disabled < 0) disabled = getenv("INDELGPU_NO_GRAPH_INDEX") != NULL;
    if (g != g_graph || g->node_list == NULL) reset_index(g);          /* a graph starts empty (new_graph, graph.c:4-9) */
    const evidence* e = n->val;
    if (e->type != SPLIT_READ && e->type != PAIRED_READ) g_plain = 0;
    if (disabled || !g_plain) { indelgpu_reference_add_node(g, n); return; }

    /* graph.c:93 asserts the sort order against every earlier node; the list is sorted iff every neighbour pair is */
    node* const head = g->node_list;
    if (head != NULL) forceassert(((evidence*)head->val)->b1 <= e->b1);

    /* the earlier nodes an edge could go to, most recent first */
    key_slot* slot = NULL;
    int chain;
    if (e->type == SPLIT_READ) {
        if (2 * (g_used + 1) > g_nslots) grow_slots();
        slot = find_slot((int32_t)e->variantclass, e->b1, e->b2);
        chain = slot->last;
    } else chain = g_last_paired;
    int k = 0;
    for (int c = chain; c >= 0; c = g_ent[c].prev) {
        if (k == g_captmp) {
            g_captmp = g_captmp ? 2 * g_captmp : 256;
            g_tmp = ckrealloc(g_tmp, sizeof(node*) * (size_t)g_captmp);
            g_saved = ckrealloc(g_saved, sizeof(node*) * (size_t)g_captmp);
        }
        g_tmp[k++] = g_ent[c].n;
    }
    /* link them into a list of their own, let the reference's add_node see only that list, restore */
    for (int i = 0; i < k; i++) { g_saved[i] = g_tmp[i]->next; g_tmp[i]->next = (i + 1 < k) ? g_tmp[i + 1] : NULL; }
    g->node_list = k > 0 ? g_tmp[0] : NULL;
    indelgpu_reference_add_node(g, n);                                 /* tests, make_edge, hashtable, sladdhead: the reference's */
    for (int i = 0; i < k; i++) g_tmp[i]->next = g_saved[i];
    n->next = head;                                                    /* what sladdhead(&g->node_list, n) does on the full list */
    g->node_list = n;

    if (e->type == SPLIT_READ) {
        if (slot->last < 0) { slot->vclass = (int32_t)e->variantclass; slot->b1 = e->b1; slot->b2 = e->b2; g_used++; }
        slot->last = push_entry(n, slot->last);
    } else g_last_paired = push_entry(n, g_last_paired);
}
