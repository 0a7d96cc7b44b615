// Launch parameters of the training-loop kernels (pinn_train.cu), shared with pinn_capi.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pinn_launch.h"

namespace pinn {

struct SampleParams {
  long long n;
  unsigned long long seed;
  unsigned long long* batch_counter;  // device: index of the batch being drawn (advanced by the sampler itself)
  float xL, xR, yL, yR, zL, zR, RL, RR;
  float cutoff, bcutoff;
  float *x, *y, *z, *R;
  uint8_t* mask;
  unsigned long long* counts;  // device, 2: sizes of the two boundary sets
  long long index_offset;      // data-parallel shard: this rank draws points [index_offset, index_offset + n) of the batch
};

struct AdamParams {
  double *theta, *m, *v;        // device float64 [1521]
  const double* grad;           // dLtot/dtheta
  const double* sums;           // the 8 loss sums of the step
  float* theta32;               // refreshed float32 copy the next step kernel reads
  unsigned long long* step;     // device: optimizer steps done
  double *best_loss, *best_theta;
  long long* best_step;
  double* hist;                 // [hist_cap][4] {Ltot, Lpde, Lbc, E} or NULL
  long long hist_cap, n;
  double lr, beta1, beta2, eps, best_after;
  uint32_t grad_mask;
  int best_mode, hist_mean_E;
};

// dp.world > 1: the set sizes are summed over the ranks (same exchange buffers as the gradient sum) and weights become
// {1/(world*n), 1/|set1|, 1/|set2|} of the GLOBAL batch
cudaError_t launch_sample(const SampleParams& s, double* weights, const DpArgs& dp, cudaStream_t st);
cudaError_t launch_adam(const AdamParams& a, cudaStream_t st);
cudaError_t launch_enet_curve(const float* theta, const double* R, int n, double* E, double* dE, double* d2E, double* gate,
                              cudaStream_t st);

}  // namespace pinn
