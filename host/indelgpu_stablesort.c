/* Row f4 of SURVEY.md section 8 (the deterministic tie order): slsort (src/slinklist.c:121-147) sorts evidence, graph
 * nodes and variants with libc qsort, whose order of EQUAL elements is unspecified -- the one token in which the
 * reference built on glibc 2.39 differs from its own golden file (BF=52,48 against BF=48,52 at POS 1000 of
 * test_data/indelminer.expected.vcf) comes from there.  Compiling slinklist.c -- unchanged --
 * with -Dqsort=indelgpu_stable_qsort binds slsort to the merge sort below, which keeps equal elements in list order
 * whatever the C library: the output no longer depends on the libc.  Opt-in build flavour (oracle/Makefile target
 * `gpuprog_stable`): the default build keeps libc qsort, because parity is measured against the reference as built here.
 * Nothing GPU-specific in this file. */
#include <stdlib.h>
#include <string.h>

static void merge_sort(char* a, char* tmp, size_t n, size_t size, int (*cmp)(const void*, const void*))
{
    if (n < 2) return;
    const size_t h = n / 2;
    merge_sort(a, tmp, h, size, cmp);
    merge_sort(a + h * size, tmp, n - h, size, cmp);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) {
        if (cmp(a + j * size, a + i * size) < 0) memcpy(tmp + (k++) * size, a + (j++) * size, size);   /* strictly smaller: ties keep their order */
        else memcpy(tmp + (k++) * size, a + (i++) * size, size);
    }
    while (i < h) memcpy(tmp + (k++) * size, a + (i++) * size, size);
    while (j < n) memcpy(tmp + (k++) * size, a + (j++) * size, size);
    memcpy(a, tmp, n * size);
}

void indelgpu_stable_qsort(void* base, size_t n, size_t size, int (*cmp)(const void*, const void*))
{
    if (n < 2 || size == 0) return;
    char* tmp = malloc(n * size);
    if (tmp == NULL) { qsort(base, n, size, cmp); return; }
    merge_sort((char*)base, tmp, n, size, cmp);
    free(tmp);
}
