/* Minimal stand-in for R's C API, ONLY so tests can type-check r/src/rglue.c against
 * include/derfinder_b200.h in an image without R (gcc -fsyntax-only). Not R, not shipped. */
#ifndef R_STUB_H
#define R_STUB_H
#include <stddef.h>
char* R_alloc(size_t n, int size);
#endif
