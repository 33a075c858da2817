/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * Reaches the reference's `static` realign_with_indel (src/variant.c:1246-1424) without editing the
 * reference: this TU #include's the unmodified variant.c (found through -iquote $(REF)/src at build
 * time; nothing is copied into this repo) and exports a wrapper that builds the three structs the
 * function reads from plain arguments.  Built by oracle/Makefile into oracle/_ref/libref_variant.so.
 */
#include <string.h>

#include "variant.c"

/* the globals variant.c reads (variant.c:3-6; defined in indelminer.c) */
uint maxdelsize = 1000;
uint maxpedelsize = 1000000;
uint minbalance = 30;
bool call_all_indels = FALSE;

int refshim_realign_with_indel(const char* reference, int rstart, int rstop,
                               const char* query, int qstart, int qstop,
                               int vtype, unsigned vstart, unsigned vstop, const char* alternate,
                               int* subs, int* indels, int* aligned)
{
    readseg seg;
    memset(&seg, 0, sizeof(seg));
    seg.sequence = (char*)query;
    readaln rln;
    memset(&rln, 0, sizeof(rln));
    rln.segments = &seg;
    knownvariant v;
    memset(&v, 0, sizeof(v));
    v.type = (varianttype)vtype; v.start = vstart; v.stop = vstop; v.alternate = (char*)alternate;
    realign_with_indel(reference, rstart, rstop, &rln, qstart, qstop, &v, subs, indels, aligned);
    return 0;
}
