// Private to the library: the handle behind the C ABI (include/pinn_b200.h) and the error helpers.
#pragma once
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pinn_b200.h"
#include "pinn_common.cuh"
#include "pinn_launch.h"

using namespace pinn;

struct pinn_handle {
  int device = 0;
  int sm_count = 0;
  Wts* wts = nullptr;              // prepared weight image
  float* theta_dev = nullptr;      // upload block of the *_host entry: 1536 float ...
  double* weights_dev = nullptr;   // ... followed by the 3 loss weights (also the output of the set-count kernels)
  unsigned long long* counts = nullptr;
  double* partials = nullptr;      // [max_rows][NPART]
  int max_rows = 0;
  double* grid_partials = nullptr;  // [sm_count + 1][8] dense-grid quadrature rows
  unsigned long long* batch_counter = nullptr;  // device counter used by the stand-alone pinn_sample entry
  // *_host entry
  void* stage_dev = nullptr;       // coordinates + mask
  size_t stage_bytes = 0;
  double* out_pinned = nullptr;    // mapped page-locked: the reduction kernel of the *_host entry writes it directly
  double* out_mapped = nullptr;    // its device alias
  float* theta_pinned = nullptr;
  double* weights_pinned = nullptr;
  cudaStream_t s_copy = nullptr, s_main = nullptr;
  cudaEvent_t ev_copy = nullptr;
  cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr};  // *_host entry: copy of chunk c complete
  // data-parallel exchange (pinn_dp_*): own buffer, the peers' buffers as mapped on this device, and whether calls reduce
  unsigned char* dp_buf = nullptr;
  DpArgs dp;
  bool dp_on = false;
  bool dp_opened[DP_MAX_WORLD] = {false, false, false, false, false, false, false, false};  // peer[r] came from cudaIpcOpenMemHandle
  double host_us[4] = {0, 0, 0, 0};   // last pinn_loss_fwd_bwd_host call: submit, wait, copy-out, total (microseconds)
  int64_t launches = 0;
  int engine = PINN_ENGINE_TCGEN05;  // which implementation of the fused step kernel runs
  bool host_inline_params = true;     // *_host entry: theta + loss weights inside the kernel parameters (PINN_B200_HOST_INLINE=0: upload)
  bool host_zero_copy = true;         // pinn_loss_fwd_bwd_host reads page-locked inputs in place (PINN_B200_HOST_ZEROCOPY=0: stage)
  bool profiling = false;          // pinn_profile_begin/collect: CUDA events around the fused step kernel
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::string err;
  std::mutex mu;
  std::mutex host_mu;  // pinn_loss_fwd_bwd_host calls on one handle share its staging / mapped buffers: one at a time
};

// Every entry point works on the handle's device and leaves the caller's current device as it found it.
struct DevGuard {
  int prev = -1, dev;
  explicit DevGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
  DevGuard(const DevGuard&) = delete;
  DevGuard& operator=(const DevGuard&) = delete;
};

// The training evaluation behind pinn_loss_fwd_bwd / pinn_loss_fwd_bwd_host / the trainer (pinn_capi.cu).
// theta / weights: device pointers, or NULL with theta_inline / weights_inline = HOST arrays that travel inside the
// kernel parameters (tcgen05 engine).  adam (optional): optimizer step fused behind the reduction (the trainer).
// presample (optional): the trainer's next batch drawn by extra blocks of the reduction kernel.
namespace pinn { struct AdamParams; struct SampleParams; }
int loss_fwd_bwd_impl(pinn_handle* h, int variant, int64_t n, const void* x, const void* y, const void* z, const void* R,
                      int in_dtype, const uint8_t* mask, const float* theta, const double* weights, const float* theta_inline,
                      const double* weights_inline, uint32_t grad_mask, float bcutoff, double* sums, double* dtheta,
                      float* E_out, cudaStream_t st, const pinn::AdamParams* adam = nullptr,
                      unsigned long long* adam_ticket = nullptr, const pinn::SampleParams* presample = nullptr);

extern std::string g_create_err;

inline int fail(pinn_handle* h, int code, const char* what) {
  char buf[512];
  if (code > 0)
    snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorString((cudaError_t)code),
             cudaGetErrorName((cudaError_t)code));
  else
    snprintf(buf, sizeof(buf), "%s", what);
  if (h) h->err = buf; else g_create_err = buf;
  return code;
}
#define CU(h, call)                                   \
  do {                                                \
    cudaError_t e__ = (call);                         \
    if (e__ != cudaSuccess) return fail(h, (int)e__, #call); \
  } while (0)

