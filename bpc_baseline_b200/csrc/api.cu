// Library-level entry points of libbpc_b200: ABI version, error strings, launch counter.
#include "common.cuh"

namespace bpc {
std::atomic<unsigned long long> g_launches{0};
}

extern "C" int bpc_abi_version(void) { return BPC_ABI_VERSION; }

extern "C" const char* bpc_error_string(int code) {
    switch (code) {
        case BPC_OK: return "ok";
        case BPC_EINVAL: return "invalid argument (size, null pointer or unsupported value)";
        case BPC_EALIGN: return "pointer is not aligned as documented in bpc_b200.h";
        case BPC_EWORKSPACE: return "workspace too small";
        case BPC_EUNSUPPORTED: return "this variant does not exist for the configuration (see include/bpc_b200.h)";
        case BPC_ETOOBIG: return "problem exceeds a documented limit (BPC_MAX_DET, BPC_MAX_ROI_WIDTH, shared memory)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

extern "C" unsigned long long bpc_launch_count(void) { return bpc::g_launches.load(std::memory_order_relaxed); }
