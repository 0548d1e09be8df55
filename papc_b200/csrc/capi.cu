// capi.cu -- status strings and the thread-local CUDA error slot of the C ABI.
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace papc {
thread_local int g_last_cuda_error = 0;
thread_local unsigned long long g_launch_count = 0;
thread_local Prof g_prof;

ProfScope::ProfScope(cudaStream_t stream, const char *name, long long M, int cin, int cout,
                     double flops, double bytes)
    : st(stream) {
    Prof &p = g_prof;
    if (!p.on || p.n >= p.cap) return;
    ProfRec *q = &p.rec[p.n];
    if (cudaEventCreate(&q->e0) != cudaSuccess) return;
    if (cudaEventCreate(&q->e1) != cudaSuccess) { cudaEventDestroy(q->e0); return; }
    snprintf(q->name, sizeof(q->name), "%s", name);
    q->M = M; q->cin = cin; q->cout = cout; q->flops = flops; q->bytes = bytes;
    cudaEventRecord(q->e0, st);
    r = q;
    ++p.n;
}
ProfScope::~ProfScope() {
    if (r != nullptr) cudaEventRecord(r->e1, st);
}
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

// ---- launch profiler -------------------------------------------------------------------------
extern "C" int papc_prof_enable(int on) {
    papc::Prof &p = papc::g_prof;
    if (on && p.rec == nullptr) {
        p.cap = 8192;
        p.rec = static_cast<papc::ProfRec *>(calloc((size_t)p.cap, sizeof(papc::ProfRec)));
        if (p.rec == nullptr) { p.cap = 0; return PAPC_EINVAL; }
    }
    p.on = on != 0;
    return PAPC_OK;
}

extern "C" int papc_prof_reset(void) {
    papc::Prof &p = papc::g_prof;
    for (int i = 0; i < p.n; ++i) {
        cudaEventDestroy(p.rec[i].e0);
        cudaEventDestroy(p.rec[i].e1);
    }
    p.n = 0;
    return PAPC_OK;
}

extern "C" int papc_prof_count(void) { return papc::g_prof.n; }

extern "C" int papc_prof_get(int i, char *name, int name_cap, int64_t *M, int32_t *cin,
                             int32_t *cout, double *flops, double *bytes, float *ms) {
    papc::Prof &p = papc::g_prof;
    if (i < 0 || i >= p.n) return PAPC_EINVAL;
    const papc::ProfRec &r = p.rec[i];
    if (name != nullptr && name_cap > 0) snprintf(name, (size_t)name_cap, "%s", r.name);
    if (M) *M = r.M;
    if (cin) *cin = r.cin;
    if (cout) *cout = r.cout;
    if (flops) *flops = r.flops;
    if (bytes) *bytes = r.bytes;
    if (ms) {
        PAPC_CUDA_TRY(cudaEventSynchronize(r.e1));
        PAPC_CUDA_TRY(cudaEventElapsedTime(ms, r.e0, r.e1));
    }
    return PAPC_OK;
}
