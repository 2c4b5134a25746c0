/* The 13 OpenImageDenoise C entry points Utils::OIDN_denoise uses (source/utils.cpp:144-196), implemented over the library's own
 * GPU denoise stage (b200rt_denoise, csrc/denoise.cu). The reference does not ship the OIDN binaries; with this file its unmodified
 * main.cpp links, runs end to end, and the three "denoised" PNGs it writes (main.cpp:118-125) are filtered on the GPU. Not OIDN:
 * an edge-avoiding a-trous filter in OIDN's place. Only what utils.cpp calls is implemented: one float3 "color" image
 * filtered in place into "output" (the reference binds both names to the same buffer, utils.cpp:158-159). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200rt.h"

typedef struct { void* data; size_t bytes; } nbuf;
typedef struct { nbuf* color; nbuf* output; size_t w, h; } nfilter;
static int g_dev;
static const char* g_err;

void* oidnNewDevice(int type) { (void)type; return &g_dev; }
void oidnCommitDevice(void* d) { (void)d; }
void* oidnNewBuffer(void* d, size_t bytes) { (void)d; nbuf* b = (nbuf*)malloc(sizeof(nbuf)); b->data = malloc(bytes); b->bytes = bytes; return b; }
void* oidnGetBufferData(void* b) { return ((nbuf*)b)->data; }
void* oidnNewFilter(void* d, const char* type) { (void)d; (void)type; return calloc(1, sizeof(nfilter)); }
void oidnSetFilterImage(void* f, const char* name, void* buf, int fmt, size_t w, size_t h, size_t off, size_t ps, size_t rs)
{
    nfilter* flt = (nfilter*)f;
    (void)fmt; (void)off; (void)ps; (void)rs;
    if (!strcmp(name, "color")) flt->color = (nbuf*)buf;
    else if (!strcmp(name, "output")) flt->output = (nbuf*)buf;
    flt->w = w; flt->h = h;
}
void oidnSetFilterBool(void* f, const char* name, int v) { (void)f; (void)name; (void)v; }
void oidnCommitFilter(void* f) { (void)f; }
void oidnExecuteFilter(void* f)
{
    nfilter* flt = (nfilter*)f;
    if (!flt->color || !flt->output) { g_err = "color / output image not set"; return; }
    if (b200rt_denoise((const float*)flt->color->data, 3, (int)flt->w, (int)flt->h, 1.0f, 0, 0.0f, (float*)flt->output->data)) g_err = b200rt_last_error();
}
int oidnGetDeviceError(void* d, const char** msg) { (void)d; if (msg) *msg = g_err; const int e = g_err != 0; g_err = 0; return e; }
void oidnReleaseBuffer(void* b) { if (b) { free(((nbuf*)b)->data); free(b); } }
void oidnReleaseFilter(void* f) { free(f); }
void oidnReleaseDevice(void* d) { (void)d; }
