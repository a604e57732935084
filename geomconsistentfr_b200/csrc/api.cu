// libgfr_b200: version / error strings.
#include "gfr_common.cuh"

extern "C" int gfr_version(void) { return 100; }   // 0.1.0

extern "C" const char* gfr_error_string(int code) {
  switch (code) {
    case GFR_OK: return "ok";
    case GFR_E_NULL: return "gfr: required pointer is NULL";
    case GFR_E_SHAPE: return "gfr: unsupported shape";
    case GFR_E_ARG: return "gfr: bad argument";
    case GFR_E_UNSUPPORTED: return "gfr: unsupported configuration";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gfr: unknown error";
  }
}
