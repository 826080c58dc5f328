// Force-included when compiling the UNMODIFIED reference sources on Linux (oracle/ref_witness/Makefile):
// the two MSVC-only CRT calls they use.  Nothing here changes arithmetic.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
static inline int fopen_s(FILE **f, const char *name, const char *mode) { *f = fopen(name, mode); return *f ? 0 : 1; }
#define sprintf_s(buf, ...) snprintf(buf, sizeof(buf), __VA_ARGS__)
