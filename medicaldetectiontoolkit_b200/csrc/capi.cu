// Library-level entry points of libmdt_b200.so (version, error text, launch counter).
#include "mdt_common.cuh"

namespace mdt {
unsigned long long g_launch_count = 0;
}

extern "C" {

int mdt_version(void) { return 100; }  // 0.1.0

unsigned long long mdt_launch_count(void) { return mdt::g_launch_count; }

const char *mdt_error_string(int code) {
    switch (code) {
        case MDT_OK: return "success";
        case MDT_EINVAL: return "mdt: invalid argument";
        case MDT_EWORKSPACE: return "mdt: workspace too small";
        case MDT_EUNSUPPORTED: return "mdt: unsupported shape or size";
        case MDT_EDRIVER: return "mdt: CUDA driver entry point unavailable or tensor-map encode failed";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "mdt: unknown error";
}

}  // extern "C"
