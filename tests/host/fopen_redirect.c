/*
 * fopen_redirect.c -- test helper (LD_PRELOAD): the reference's enabled test case reads its input from an absolute path
 * on its author's machine (reference test/test.cpp:118, "/home/kreimer/data.csv") and exits 0 when the file is missing.
 * To run that test UNCHANGED on real data without touching the file system outside the repository, fopen() of exactly
 * that path is redirected to $VISO_TEST_DATA_CSV.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char* redirect(const char* path)
{
    const char* to = getenv("VISO_TEST_DATA_CSV");
    return (to && path && !strcmp(path, "/home/kreimer/data.csv")) ? to : path;
}

FILE* fopen(const char* path, const char* mode)
{
    static FILE* (*real)(const char*, const char*);
    if (!real) real = (FILE * (*)(const char*, const char*)) dlsym(RTLD_NEXT, "fopen");
    return real(redirect(path), mode);
}

FILE* fopen64(const char* path, const char* mode)
{
    static FILE* (*real)(const char*, const char*);
    if (!real) real = (FILE * (*)(const char*, const char*)) dlsym(RTLD_NEXT, "fopen64");
    return real(redirect(path), mode);
}
