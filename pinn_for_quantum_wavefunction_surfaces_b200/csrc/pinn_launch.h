// Launch parameters and host-side launchers shared by the kernel files (pinn_step_tc.cu, pinn_reduce.cu, pinn_train.cu) and pinn_capi.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "pinn_common.cuh"

namespace pinn {

// Dense-grid mode of the inference kernel (poc/main.py:438-517, 639-676): point p of the n = nx*ny*nz
// grid is (i,j,k) = meshgrid 'ij' order of three linspaces, R is one value, and instead of per-point stores the kernel
// accumulates the quadrature sums  S = sum_p w_p * {psi H psi, psi^2, lcao H lcao, lcao^2, (dV/dR) psi^2}.
struct GridDesc {
  int on;
  int nx, ny, nz;
  double x0, dx, y0, dy, z0, dz;  // linspace start / step
  double R;
  const double *wx, *wy, *wz;     // 1-D quadrature weights (Simpson), device
  double* partials;               // [gridDim.x][8]
};

struct StepParams {
  const void *x, *y, *z, *R;
  const uint8_t* mask;
  const Wts* wts;         // A/B builds, FFMA engine: weight image built by prep_weights_kernel
  const float* theta;     // the raw parameter vector; every CTA builds its own operand images in shared memory
  const double* weights;  // {w_pde, w_bc1, w_bc2}
  double* partials;       // [gridDim.x][NPART]
  float* E_out;
  float *psi, *lap, *hpsi, *res;  // fields mode outputs (nullable)
  long long n;
  int in_f64;
  float bcut;
  VariantCoef vc;
  int base_grads, gate_grads;  // reverse sweeps wanted (fine-tune mode clears both)
  GridDesc grid;
  // HOST pointers, read by the launchers only: when set, theta (1521 float; the *_host entry) and / or the three loss
  // weights travel inside the kernel parameters instead of through a preceding host-to-device copy
  const float* theta_inline;
  const double* weights_inline;
  int w_by_value;         // set by the launcher: the loss weights are in the parameter block behind this struct
  // the caller's 16 parameter tensors instead of a packed vector (pinn_loss_fwd_bwd_tensors): device pointers in
  // canonical (state_dict) order, float32 or float64, nn.Linear (out,in) or train.py (in,out) layout; every CTA gathers
  // them while it builds its operand images - no packing kernel, no packed copy in front of the launch
  const void* theta_tensors[16];
  int theta_from_tensors, tensors_f64, tensors_in_out;
  int E_f64;              // E_out points at float64 (the reference's dtype) instead of float32
};

// Launch as a programmatic dependent of the kernel in front of it in the stream: the grid may be set up (and, where
// resources allow, its CTAs made resident) while that kernel drains.  Every kernel launched this way executes
// griddepcontrol.wait (pdl_wait) before its first global-memory access, which returns once the kernel in front has
// completed and its writes are visible - stream semantics are unchanged, only launch latency is hidden.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

#ifdef PINN_AB_BUILD  // FFMA engine (pinn_step_ffma.cu), A/B builds only
cudaError_t launch_step(int nev, bool train, const StepParams& p, int grid, cudaStream_t st);
cudaError_t launch_prep(const float* theta, Wts* out, cudaStream_t st);
#endif
// tcgen05 engine (pinn_step_tc.cu): super-tiles of 128 points, one persistent CTA per SM
cudaError_t launch_step_tc(int nev, bool train, const StepParams& p, int grid, cudaStream_t st);
cudaError_t launch_grid_finish(const double* partials, int nrows, double* out, cudaStream_t st);
cudaError_t launch_count(const StepParams& p, unsigned long long* counts, double* weights, cudaStream_t st);
// Data-parallel exchange fused into the reduction kernel: every rank owns one exchange buffer that all
// peers of the box can write through NVLink peer memory (cudaIpc / peer access).
//   rows [2 slots][DP_MAX_WORLD][NPART] x 16 B : rank r deposits its reduced row in slot (step & 1), index r, of EVERY
//        rank; each float64 is two 8-byte words {32 data bits, 32-bit step number} (pinn_reduce.cu: reduce_partials_kernel)
//   ctl  8 x u64 {exchanges completed, blocks done, status, set-count exchanges completed, ...}
//   cnt  [2 slots][DP_MAX_WORLD][2] x 8 B : the sampler's boundary-set sizes, same {data, step} words (pinn_train.cu)
constexpr int DP_MAX_WORLD = 8;
// How long a rank polls for a peer's contribution before it declares the run failed (SM cycles, ~30 s): long enough for any
// host-side skew between the ranks' launches (one rank writing a checkpoint, drawing a batch on the CPU ...), short
// enough that a dead peer does not hang the GPU.
constexpr long long DP_TIMEOUT_CYCLES = 60000000000ll;
constexpr int DP_BLOCKS = NPART / 32;
constexpr size_t DP_ROWS_BYTES = 2ull * DP_MAX_WORLD * NPART * 16;
constexpr size_t DP_CTL_BYTES = 64;
constexpr size_t DP_CNT_BYTES = 2ull * DP_MAX_WORLD * 2 * 8;
constexpr size_t DP_BUFFER_BYTES = DP_ROWS_BYTES + DP_CTL_BYTES + DP_CNT_BYTES;
struct DpArgs {
  int world = 0;  // <= 1: no exchange
  int rank = 0;
  unsigned char* peer[DP_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // exchange buffers, [rank] = own
  long long timeout_cycles = DP_TIMEOUT_CYCLES;  // pinn_dp_set_timeout
};
// weights: device pointer, or NULL with weights_inline (host, 3 double) carried in the kernel parameters
// adam (optional): the optimizer step of the device-resident trainer fused behind the reduction (pinn_train.h)
// presample (optional): extra blocks of the same launch draw the trainer's next batch (pinn_sample.cuh)
struct AdamParams;
struct SampleParams;
// E_f64: E_out holds float64; dtheta_in_out: the 2-D tensors of dtheta are written in train.py's (in,out) layout
cudaError_t launch_reduce(const double* partials, int nrows, const double* weights, const double* weights_inline,
                          uint32_t grad_mask, double* dtheta, double* sums, const float* E_out, long long n, const DpArgs& dp,
                          cudaStream_t st, const AdamParams* adam = nullptr, unsigned long long* adam_ticket = nullptr,
                          const SampleParams* presample = nullptr, int E_f64 = 0, int dtheta_in_out = 0);
// boundary index sets (int64 row indices, device) -> per-point mask bytes (bit 0: set 1, bit 1: set 2); mask 4-byte aligned
cudaError_t launch_mask_from_index_sets(const long long* idx1, long long n1, const long long* idx2, long long n2, uint8_t* mask,
                                        long long n, cudaStream_t st);

}  // namespace pinn
