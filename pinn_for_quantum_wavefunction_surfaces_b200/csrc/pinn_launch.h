// Launch parameters and host-side launchers shared by pinn_kernels.cu and pinn_capi.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pinn_common.cuh"

namespace pinn {

struct StepParams {
  const void *x, *y, *z, *R;
  const uint8_t* mask;
  const Wts* wts;
  const double* weights;  // {w_pde, w_bc1, w_bc2}
  double* partials;       // [gridDim.x][NPART]
  float* E_out;
  float *psi, *lap, *hpsi, *res;  // fields mode outputs (nullable)
  long long n;
  int in_f64;
  float bcut;
  VariantCoef vc;
  int base_grads, gate_grads;  // reverse sweeps wanted (fine-tune mode clears both)
};

cudaError_t launch_step(int nev, bool train, const StepParams& p, int grid, cudaStream_t st);
int step_groups();
// tcgen05 engine (pinn_step_tc.cu): super-tiles of 128 points, one persistent CTA per SM
cudaError_t launch_step_tc(int nev, bool train, const StepParams& p, int grid, cudaStream_t st);
cudaError_t launch_prep(const float* theta, Wts* out, cudaStream_t st);
cudaError_t launch_count(const StepParams& p, unsigned long long* counts, double* weights, cudaStream_t st);
cudaError_t launch_reduce(const double* partials, int nrows, const double* weights, uint32_t grad_mask, double* dtheta,
                          double* sums, const float* E_out, long long n, cudaStream_t st);

}  // namespace pinn
