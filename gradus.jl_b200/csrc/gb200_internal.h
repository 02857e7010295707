// gb200_internal.h -- launchers shared between gb200_trace.cu (kernels) and gb200_api.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
struct GbParams;
cudaError_t gb200_launch_trace(const GbParams& P, int sm_count, cudaStream_t stream, int* blocks_out);
cudaError_t gb200_launch_hist(const double* g, const double* f, int64_t n, const double* bins, int nbins, int right_closed,
                              double* partial, int nblocks, double* out, cudaStream_t stream);
cudaError_t gb200_launch_bucket2d(const double* x, const double* y, const double* w, int64_t n, const double* xb, int nx, const double* yb, int ny,
                                  double scale, long long* acc, double* out, int blocks, cudaStream_t stream);
cudaError_t gb200_launch_hist128_finish(const unsigned long long* acc, int nbins, double scale, double* out, cudaStream_t stream);
cudaError_t gb200_launch_dfma(double* d_out, int blocks, int iters, cudaStream_t stream);
cudaError_t gb200_launch_dfma_mix(double* d_out, int blocks, int iters, int mix, cudaStream_t stream);
cudaError_t gb200_launch_path(const GbParams& P, const double* d_u0, int cap, double* d_lambda, double* d_u, int* d_meta, cudaStream_t stream);
cudaError_t gb200_launch_debug_rhs(const GbParams& P, long long n, const double* d_u, double* d_du, cudaStream_t stream);
cudaError_t gb200_launch_debug_math(long long n, const double* d_x, double* d_out5, cudaStream_t stream);
cudaError_t gb200_launch_debug_math_lo(long long n, const double* d_x, double* d_out2, cudaStream_t stream);

// forward-mode (dual-number) traces and the single-ray path recorder: gb200_dual.cu
struct GbDualIO {
    int64_t n;
    int32_t npartials, norm_partials;
    const double *alpha, *beta, *dalpha, *dbeta, *height; // device
    int32_t* status;
    double* lambda;
    double* x[4];
    double* v[4];
    double *g, *dg, *rho, *drho;
    int32_t *naccept, *nreject, *flags;
};
cudaError_t gb200_launch_dual(const GbParams& P, const GbDualIO& io, cudaStream_t stream);
// closest approach to a target point (optimize_for_target's objective): the generic integrator over the rays of P
cudaError_t gb200_launch_target(const GbParams& P, cudaStream_t stream);
