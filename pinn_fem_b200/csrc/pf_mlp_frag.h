// pf_mlp_frag.cu: register-resident DMMA kernels of the material networks (large point sets, batched problems).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pf_mlp.cuh"

// shapes the kernels cover: input_dim <= 3, width <= 24, 1-3 hidden layers for the backward pass
bool pf_mlp_frag_supported(const PfMlpDesc& d, bool backward);
// doubles of the activation record of ONE problem: hidden activations in fragment order + d value / d z per point
int64_t pf_mlp_frag_acts_len(const PfMlpDesc& d, int64_t n);
// rows of partial gradients per problem the backward kernel writes (part: [B][chunks][n_params])
int pf_mlp_frag_chunks(const PfMlpDesc& d, int64_t n, int64_t B, int sm_count);
// out[pt * ldb + p] = value of problem p (theta + p * theta_stride) at point pt; acts may be NULL (not saved)
int pf_mlp_frag_forward(const PfMlpDesc& d, const double* theta, int64_t theta_stride, int64_t B, int64_t n,
                        const double* X, const double* centroid, double load_factor, double scale, int positive,
                        double* out, int64_t ldb, double* acts, int64_t acts_stride, int sm_count, cudaStream_t st);
// g_theta[p * gt_stride + q] = sum_pt g_out[pt * ldb + p] d value / d theta_q, from the record the forward left
int pf_mlp_frag_backward(const PfMlpDesc& d, const double* theta, int64_t theta_stride, int64_t B, int64_t n,
                         const double* X, const double* centroid, double load_factor, const double* g_out, int64_t ldb,
                         const double* acts, int64_t acts_stride, double* part, double* g_theta, int64_t gt_stride,
                         int sm_count, cudaStream_t st);
