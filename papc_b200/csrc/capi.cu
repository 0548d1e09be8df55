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

// ------------------------------------------------------------------ the step's ONE exchange over peer memory
// the batch-sharded forward ends with an all-gather of the per-shard [B/G, C] features (128 KiB per
// rank).  NCCL's all-gather costs ~30 us of latency at 8 ranks for these bytes; here every rank stores its rows
// straight into every peer's (symmetric-memory) result buffer over NVLink / NVSwitch -- plain 16-byte peer stores,
// one launch -- and a signal-pad barrier (the caller's, torch symmetric memory) closes the exchange.
namespace papc {
struct PeerPtrs { float *p[16]; };
__global__ void __launch_bounds__(256)
p2p_allgather_kernel(const float *__restrict__ local, long long n, const PeerPtrs peers, int rank, int world) {
    const int peer = blockIdx.y;
    float *dst = peers.p[peer] + (long long)rank * n;
    const long long n4 = n / 4;
    const float4 *src4 = reinterpret_cast<const float4 *>(local);
    float4 *dst4 = reinterpret_cast<float4 *>(dst);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst4[i] = src4[i];
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = local[i];
}
}  // namespace papc

extern "C" int papc_p2p_allgather_f32(const float *local, int64_t n_per_rank, void *const *peer_bufs_host, int rank,
                                      int world, papc_stream_t stream) {
    using namespace papc;
    if (!local || !peer_bufs_host || n_per_rank < 0 || world < 1 || world > 16 || rank < 0 || rank >= world) return PAPC_EINVAL;
    if ((reinterpret_cast<uintptr_t>(local) & 15u) != 0 || (n_per_rank * 4) % 16 != 0) return PAPC_EINVAL;
    PeerPtrs pp{};
    for (int i = 0; i < world; ++i) {
        if (!peer_bufs_host[i] || (reinterpret_cast<uintptr_t>(peer_bufs_host[i]) & 15u) != 0) return PAPC_EINVAL;
        pp.p[i] = reinterpret_cast<float *>(peer_bufs_host[i]);
    }
    if (n_per_rank == 0) return PAPC_OK;
    cudaStream_t st = as_stream(stream);
    long long blocks = (n_per_rank / 4 + 255) / 256;
    if (blocks > 16) blocks = 16;           // 8 peers x 16 blocks: enough stores in flight for 128 KiB per peer
    if (blocks < 1) blocks = 1;
    ProfScope prof(st, "p2p_allgather", n_per_rank, world, 0, 0.0, 4.0 * n_per_rank * (world + 1));
    p2p_allgather_kernel<<<dim3((unsigned)blocks, (unsigned)world), 256, 0, st>>>(local, n_per_rank, pp, rank, world);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}
