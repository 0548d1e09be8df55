// capi.cu -- status strings and the thread-local CUDA error slot of the C ABI.
#include "common.cuh"

namespace papc {
thread_local int g_last_cuda_error = 0;
thread_local unsigned long long g_launch_count = 0;
}

extern "C" const char *papc_status_string(int status) {
    switch (status) {
        case PAPC_OK: return "PAPC_OK";
        case PAPC_EINVAL: return "PAPC_EINVAL: invalid argument";
        case PAPC_EWORKSPACE: return "PAPC_EWORKSPACE: workspace missing or too small";
        case PAPC_ECUDA: return "PAPC_ECUDA: CUDA runtime error";
        case PAPC_EUNSUPPORTED: return "PAPC_EUNSUPPORTED: outside what this build implements";
        default: return "PAPC: unknown status";
    }
}

extern "C" int papc_abi_version(void) { return PAPC_ABI_VERSION; }

extern "C" int papc_last_cuda_error(void) { return papc::g_last_cuda_error; }

extern "C" uint64_t papc_launch_count(void) { return papc::g_launch_count; }
