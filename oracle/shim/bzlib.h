/* stand-in for <bzlib.h>: declarations of the four libbz2 entry points cqf/reader.h and cqf/CQF_mt.h name.
 * Linked against the image's libbz2.so.1 when present (oracle/Makefile), so -f b works too. */
#pragma once
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef void BZFILE;
#define BZ_OK 0
#define BZ_STREAM_END 4
BZFILE* BZ2_bzReadOpen(int* bzerror, FILE* f, int verbosity, int small, void* unused, int nUnused);
int BZ2_bzRead(int* bzerror, BZFILE* b, void* buf, int len);
void BZ2_bzReadClose(int* bzerror, BZFILE* b);
void BZ2_bzReadGetUnused(int* bzerror, BZFILE* b, void** unused, int* nUnused);
#ifdef __cplusplus
}
#endif
