// sa_mlp_tc.cuh -- interface between the MLP driver (sa_mlp.cu) and the tcgen05 layer kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace papc {
namespace tc {

struct TcArgs {
    // gathered source (layer 0 of the fused path)
    const float *xyz, *new_xyz, *feats;
    const int32_t *idx;
    int N, S, K, D;
    // plain source: x [M,cin], optional act(v) = relu(in_scale*v + in_shift)
    const float *x, *in_scale, *in_shift;
    long long M;
    int cin, cout;
    const float *bias;
    float *y;
    float *pool_max, *pool_min;
    double *stats_partial;
    long long partial_rows;
    int vec_y;
    // filled by launch()
    const float *wimg;
    int KC, kpad, stages;
};

struct TcProblem {
    int cin, cout, K, D;
    bool gather, pool;
};

// Can the tensor-core kernel run this layer (shape / shared-memory limits)?
bool eligible(const TcProblem &p);
// Bytes of the pre-split, pre-swizzled W image the kernel stages from.
size_t wimg_bytes(int cin, int cout);
// Builds the W image into `wimg` and launches the layer.  Returns a papc_status.
int launch(TcArgs a, const float *W, float *wimg, bool gather, int order, cudaStream_t st);

}  // namespace tc
}  // namespace papc
