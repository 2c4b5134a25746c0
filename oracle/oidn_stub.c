/* TEST INFRASTRUCTURE ONLY.
 * Utils::OIDN_denoise (reference source/utils.cpp:144-196) references 13 OpenImageDenoise C symbols; the OIDN
 * binaries are not shipped with the reference (.gitignore:14-16) and the denoiser is post-processing outside the
 * hot path, never called by the oracle driver. These stubs only let the shared library resolve at load time. */
#include <stdio.h>
#include <stdlib.h>

#define OIDN_STUB(name) void* name(void) { fprintf(stderr, "oracle: " #name " called - OIDN is not available\n"); abort(); return 0; }

OIDN_STUB(oidnCommitDevice)
OIDN_STUB(oidnCommitFilter)
OIDN_STUB(oidnExecuteFilter)
OIDN_STUB(oidnGetBufferData)
OIDN_STUB(oidnGetDeviceError)
OIDN_STUB(oidnNewBuffer)
OIDN_STUB(oidnNewDevice)
OIDN_STUB(oidnNewFilter)
OIDN_STUB(oidnReleaseBuffer)
OIDN_STUB(oidnReleaseDevice)
OIDN_STUB(oidnReleaseFilter)
OIDN_STUB(oidnSetFilterBool)
OIDN_STUB(oidnSetFilterImage)
