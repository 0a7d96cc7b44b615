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

// Scratch memory of the calls ENQUEUED ON ONE STREAM.  Kernels of one stream run in order, so they may share it; calls
// that arrive on different streams (torch's current stream, the *_host entry's private stream, every trainer's own
// stream) may overlap on the device and therefore get one workspace each (ws_for, pinn_capi.cu).
struct pinn_workspace {
  cudaStream_t stream = nullptr;
  double* partials = nullptr;       // [rows][NPART] per-CTA rows of the step kernel
  int rows = 0;
  double* weights = nullptr;        // 4 doubles: the output of the set-count kernels when the caller passes no loss weights
  unsigned long long* counts = nullptr;  // [0..1] boundary-set sizes, [2] batch index of pinn_sample, [3] its block ticket
  double* grid_partials = nullptr;  // [sm_count + 1][8] dense-grid quadrature rows
  uint64_t last_use = 0;
  bool in_graph = false;            // a call on this stream was captured into a CUDA graph: the addresses above must stay valid
};

struct pinn_handle {
  int device = 0;
  int sm_count = 0;
  std::vector<pinn_workspace> ws;  // one per stream that has called in (guarded by mu)
  uint64_t ws_clock = 0;
#ifdef PINN_AB_BUILD
  Wts* wts = nullptr;              // FFMA engine (A/B builds only): prepared weight image
#endif
  float* theta_dev = nullptr;      // upload block of the *_host entry: 1536 float ...
  double* weights_dev = nullptr;   // ... followed by the 3 loss weights
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
  cudaStream_t dp_stream = nullptr;  // the exchange buffers carry ONE sequence of steps: the stream of the last exchanging call
  bool dp_stream_set = false;
  bool dp_opened[DP_MAX_WORLD] = {false, false, false, false, false, false, false, false};  // peer[r] came from cudaIpcOpenMemHandle
  double host_us[4] = {0, 0, 0, 0};   // last pinn_loss_fwd_bwd_host call: submit, wait, copy-out, total (microseconds)
  int64_t launches = 0;
  // The product library has ONE engine (tcgen05) and no run-time knobs.  -DPINN_AB_BUILD (tools/build_ab.sh) adds the
  // FFMA engine and two environment switches of the *_host entry for A/B measurements.
  int engine = PINN_ENGINE_TCGEN05;
  bool host_inline_params = true;     // *_host entry: theta + loss weights inside the kernel parameters
  bool host_zero_copy = true;         // pinn_loss_fwd_bwd_host reads page-locked inputs in place
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

// The training evaluation behind pinn_loss_fwd_bwd[_tensors] / pinn_loss_fwd_bwd_host / the trainer (pinn_capi.cu).
namespace pinn { struct AdamParams; struct SampleParams; }
struct LossCall {
  int variant = 0;
  int64_t n = 0;
  const void *x = nullptr, *y = nullptr, *z = nullptr, *R = nullptr;  // device (or mapped page-locked host) columns
  int in_dtype = PINN_F32;
  const uint8_t* mask = nullptr;
  // theta: exactly one of - packed float32 device vector; HOST float32[1521] that travels inside the kernel parameters;
  // the caller's 16 tensors (device pointers in canonical order, float32/float64, (out,in) or train.py (in,out) layout)
  const float* theta = nullptr;
  const float* theta_inline = nullptr;
  const void* const* theta_tensors = nullptr;
  int tensors_f64 = 0, tensors_in_out = 0;
  // loss weights: device pointer, or HOST double[3] carried by value, or neither (sets counted on the device first)
  const double* weights = nullptr;
  const double* weights_inline = nullptr;
  uint32_t grad_mask = 0xFFFFu;
  float bcutoff = 17.5f;
  double *sums = nullptr, *dtheta = nullptr;
  void* E_out = nullptr;
  int E_f64 = 0;          // E_out is float64
  int dtheta_in_out = 0;  // 2-D tensors of dtheta in train.py's (in,out) layout
  const pinn::AdamParams* adam = nullptr;       // optimizer step fused behind the reduction (the trainer)
  unsigned long long* adam_ticket = nullptr;
  const pinn::SampleParams* presample = nullptr;  // the trainer's next batch drawn by extra blocks of the reduction kernel
};
int loss_fwd_bwd_impl(pinn_handle* h, const LossCall& c, cudaStream_t st);

// The workspace of stream `st` with room for at least `rows` partial rows (created / grown on first use; the handle mutex
// is held by the caller).  NULL + error message on failure.
pinn_workspace* ws_for(pinn_handle* h, cudaStream_t st, int rows);
void ws_release(pinn_handle* h, cudaStream_t st);

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

