/* Pass-through stand-in for the 13 OpenImageDenoise C entry points the reference's Utils::OIDN_denoise uses
 * (utils.cpp:144-196). The reference does not ship the OIDN binaries; with this file its unmodified main.cpp links and
 * runs end to end, the "denoised" outputs simply equal the noisy render. Build scaffolding for the drop-in demo only. */
#include <stdlib.h>

typedef struct { void* data; } nbuf;
static int g_dev, g_filter;

void* oidnNewDevice(int type) { (void)type; return &g_dev; }
void oidnCommitDevice(void* d) { (void)d; }
void* oidnNewBuffer(void* d, size_t bytes) { (void)d; nbuf* b = (nbuf*)malloc(sizeof(nbuf)); b->data = malloc(bytes); return b; }
void* oidnGetBufferData(void* b) { return ((nbuf*)b)->data; }
void* oidnNewFilter(void* d, const char* type) { (void)d; (void)type; return &g_filter; }
void oidnSetFilterImage(void* f, const char* name, void* buf, int fmt, size_t w, size_t h, size_t off, size_t ps, size_t rs)
{ (void)f; (void)name; (void)buf; (void)fmt; (void)w; (void)h; (void)off; (void)ps; (void)rs; }
void oidnSetFilterBool(void* f, const char* name, int v) { (void)f; (void)name; (void)v; }
void oidnCommitFilter(void* f) { (void)f; }
void oidnExecuteFilter(void* f) { (void)f; }   /* input and output share one buffer in the reference: identity */
int oidnGetDeviceError(void* d, const char** msg) { (void)d; if (msg) *msg = 0; return 0; }
void oidnReleaseBuffer(void* b) { if (b) { free(((nbuf*)b)->data); free(b); } }
void oidnReleaseFilter(void* f) { (void)f; }
void oidnReleaseDevice(void* d) { (void)d; }
