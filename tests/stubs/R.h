/* Stub of <R.h> for the syntax check of r/src/ruvnb_shim.c (tests/test_r_shim.py): R is not installed in this image. */
#ifndef STUB_R_H
#define STUB_R_H
#include <stddef.h>
char* R_alloc(size_t n, int size);
#endif
