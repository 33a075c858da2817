/* Prototype of the GPU-backed realign_with_indel (host/indelgpu_support.c).  In the reference the
 * function is file-local (src/variant.c:1246); a build that binds the GPU version drops that definition
 * and declares this one instead (oracle/Makefile target gpuprog_annotate). */
#ifndef INDELGPU_SUPPORT_H
#define INDELGPU_SUPPORT_H

#include "variant.h"        /* the reference's header: readaln, knownvariant */

void realign_with_indel(const char* const reference, const int rstart, const int rstop,
                        const readaln* const rln, const int qstart, const int qstop,
                        const knownvariant* const variant,
                        int* const alnsubs, int* const alnindels, int* const alnaligned);

#endif
