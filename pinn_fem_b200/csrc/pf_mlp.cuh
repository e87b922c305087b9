// Shared definitions of the material-network kernels.
#pragma once
#include <stdint.h>

#define PF_MLP_MAX_LAYERS 8

// Flat parameter layout of SimpleNN in nn.Module.parameters() order:
// W_0[w][in], b_0[w], W_1[w][w], b_1[w], ..., W_out[1][w], b_out[1].
struct PfMlpDesc {
    int in_dim, L, w, wp;  // wp = w rounded up to a multiple of 4
    int n_params;
    int w_off[PF_MLP_MAX_LAYERS + 1];
    int b_off[PF_MLP_MAX_LAYERS + 1];
};

int pf_mlp_make_desc(int input_dim, int hidden_layers, int width, PfMlpDesc* d);

// softplus(z) * scale with torch's defaults (beta 1, threshold 20), fem/properties.py:153-156
__host__ __device__ __forceinline__ double pf_mlp_output(double z, double scale, int positive) {
    if (!positive) return z * scale;
    const double sp = z > 20.0 ? z : log1p(exp(z));
    return sp * scale;
}

// d(value)/dz
__host__ __device__ __forceinline__ double pf_mlp_output_grad(double z, double scale, int positive) {
    if (!positive || z > 20.0) return scale;
    return scale / (1.0 + exp(-z));
}
