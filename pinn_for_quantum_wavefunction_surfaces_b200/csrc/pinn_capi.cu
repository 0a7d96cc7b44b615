// C ABI of the PINN hot path (see include/pinn_b200.h for the contract).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pinn_handle.h"
#include "pinn_train.h"

std::string g_create_err;

// upload block of the *_host entry: theta as float (padded to 1536) followed by the 3 loss weights
constexpr int HOST_IN_THETA = 1536;
constexpr size_t HOST_IN_BYTES = HOST_IN_THETA * sizeof(float) + 4 * sizeof(double);

static const int kOffsets[17] = {O_W1, O_B1, O_W2, O_B2, O_WO, O_BO, O_WE1, O_BE1, O_WE2, O_BE2, O_WE, O_BE,
                                 O_WGL, O_BGL, O_WG, O_BG, NTHETA};

extern "C" {

int pinn_version(void) { return 100; }
int pinn_theta_size(void) { return NTHETA; }
void pinn_theta_offsets(int* out) { memcpy(out, kOffsets, sizeof(kOffsets)); }

// pinn_create: a failing CUDA call releases what was built so far; the message goes to the create-error slot
#define CREATE_CU(call)                                    \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) {                              \
      pinn_destroy(h);                                     \
      return fail(nullptr, (int)e__, "pinn_create: " #call); \
    }                                                      \
  } while (0)

int pinn_create(int device, pinn_handle** out) {
  if (!out) return fail(nullptr, PINN_EINVAL, "pinn_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return fail(nullptr, (int)e, "cudaGetDeviceCount");
  if (device < 0 || device >= ndev) return fail(nullptr, PINN_EINVAL, "pinn_create: no such CUDA device");
  cudaDeviceProp prop;
  CU(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, PINN_ENOTSUP, "pinn_create: this library only carries sm_100a code (B200); no fallback exists");
  DevGuard dev_guard(device);
  pinn_handle* h = new pinn_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
#ifdef PINN_AB_BUILD  // A/B builds only (tools/build_ab.sh): the product has one engine and reads no environment
  if (const char* e = getenv("PINN_B200_ENGINE")) {
    if (!strcmp(e, "ffma")) h->engine = PINN_ENGINE_FFMA;
    else if (!strcmp(e, "tcgen05")) h->engine = PINN_ENGINE_TCGEN05;
  }
  if (const char* e = getenv("PINN_B200_HOST_ZEROCOPY")) h->host_zero_copy = strcmp(e, "0") != 0;
  if (const char* e = getenv("PINN_B200_HOST_INLINE")) h->host_inline_params = strcmp(e, "0") != 0;
  CREATE_CU(cudaMalloc(&h->wts, sizeof(Wts)));
#endif
  // theta (1536 float) and the 3 loss weights share one block so that the *_host entry uploads both with one copy
  CREATE_CU(cudaMalloc(&h->theta_dev, HOST_IN_BYTES));
  h->weights_dev = reinterpret_cast<double*>(h->theta_dev + HOST_IN_THETA);
  CREATE_CU(cudaHostAlloc(&h->out_pinned, NPART * sizeof(double), cudaHostAllocMapped));
  CREATE_CU(cudaHostGetDevicePointer(&h->out_mapped, h->out_pinned, 0));
  CREATE_CU(cudaMallocHost(&h->theta_pinned, HOST_IN_BYTES));
  h->weights_pinned = reinterpret_cast<double*>(h->theta_pinned + HOST_IN_THETA);
  CREATE_CU(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
  CREATE_CU(cudaStreamCreateWithFlags(&h->s_main, cudaStreamNonBlocking));
  CREATE_CU(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  for (int c = 0; c < 4; c++) CREATE_CU(cudaEventCreateWithFlags(&h->ev_chunk[c], cudaEventDisableTiming));
  // the workspace of the *_host entry's stream holds the rows of up to 4 chunk launches
  if (!ws_for(h, h->s_main, 4 * h->sm_count)) {
    const std::string msg = h->err;
    pinn_destroy(h);
    g_create_err = "pinn_create: " + msg;
    return PINN_EINVAL;
  }
  *out = h;
  return 0;
}

}  // extern "C"

static void ws_free(pinn_workspace& w) {
  cudaFree(w.partials); cudaFree(w.weights); cudaFree(w.counts); cudaFree(w.grid_partials);
  w = pinn_workspace();
}

// The owner of `st` is about to destroy it (a trainer): drain it and give the workspace back.
void ws_release(pinn_handle* h, cudaStream_t st) {
  for (pinn_workspace& c : h->ws)
    if (c.stream == st && c.partials) {
      cudaStreamSynchronize(st);
      ws_free(c);
    }
}

static bool stream_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); return false; }
  return cap != cudaStreamCaptureStatusNone;
}

pinn_workspace* ws_for(pinn_handle* h, cudaStream_t st, int rows) {
  constexpr size_t MAX_WS = 32;
  pinn_workspace* w = nullptr;
  for (pinn_workspace& c : h->ws)
    if (c.stream == st && c.partials) { w = &c; break; }
  if (w && w->rows >= rows) {
    // a CUDA graph captured from this stream carries the workspace's addresses: from then on it is never recycled
    if (!w->in_graph && stream_is_capturing(st)) w->in_graph = true;
    w->last_use = ++h->ws_clock;
    return w;
  }
  // first call on this stream (or more rows than before): allocate.  Not possible while the stream is being captured
  // into a CUDA graph - the caller has to make one plain call first.
  if (stream_is_capturing(st)) {
    fail(h, PINN_EINVAL, "first call on a stream under CUDA-graph capture: call once on this stream outside the capture first");
    return nullptr;
  }
  if (!w) {
    for (pinn_workspace& c : h->ws)
      if (!c.partials) { w = &c; break; }  // a released slot
    if (!w && h->ws.size() >= MAX_WS) {  // streams come and go in the caller: recycle the least recently used workspace
      size_t lru = h->ws.size();
      for (size_t i = 0; i < h->ws.size(); i++)
        if (!h->ws[i].in_graph && h->ws[i].stream != h->s_main && (lru == h->ws.size() || h->ws[i].last_use < h->ws[lru].last_use)) lru = i;
      if (lru == h->ws.size()) {
        fail(h, PINN_EINVAL, "more than 32 streams with captured graphs or trainers on one handle");
        return nullptr;
      }
      cudaDeviceSynchronize();      // its last kernels may still be running
      ws_free(h->ws[lru]);
      w = &h->ws[lru];
    } else if (!w) {
      h->ws.emplace_back();
      w = &h->ws.back();
    }
  } else {
    if (w->in_graph) {
      fail(h, PINN_EINVAL, "this stream's workspace is referenced by a captured graph and cannot grow");
      return nullptr;
    }
    cudaStreamSynchronize(st);
    ws_free(*w);
  }
  w->stream = st;
  cudaError_t e = cudaMalloc(&w->partials, (size_t)rows * NPART * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&w->weights, 4 * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&w->counts, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(w->counts, 0, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&w->grid_partials, (size_t)(h->sm_count + 1) * 8 * sizeof(double));
  if (e != cudaSuccess) {
    ws_free(*w);
    fail(h, (int)e, "workspace allocation (cudaMalloc)");
    return nullptr;
  }
  w->rows = rows;
  w->last_use = ++h->ws_clock;
  return w;
}

// With the data-parallel exchange enabled the handle's exchange buffers carry one sequence of steps: calls must reach the
// device in one order on all ranks.  A call on another stream than the previous exchanging call first waits (on the
// host) for that stream to drain.
static int dp_serialize(pinn_handle* h, cudaStream_t st) {
  if (!h->dp_on || h->dp.world <= 1) return 0;
  if (h->dp_stream_set && h->dp_stream != st) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    if (cap != cudaStreamCaptureStatusNone)
      return fail(h, PINN_EINVAL, "data-parallel exchange: the previous exchanging call used another stream; cannot switch streams under capture");
    CU(h, cudaStreamSynchronize(h->dp_stream));
  }
  h->dp_stream = st;
  h->dp_stream_set = true;
  return 0;
}

extern "C" {

static void dp_release(pinn_handle* h) {
  for (int r = 0; r < DP_MAX_WORLD; r++) {
    if (h->dp_opened[r] && h->dp.peer[r]) cudaIpcCloseMemHandle(h->dp.peer[r]);
    h->dp_opened[r] = false;
    h->dp.peer[r] = nullptr;
  }
  if (h->dp_buf) cudaFree(h->dp_buf);
  h->dp_buf = nullptr;
  h->dp = DpArgs();
  h->dp_on = false;
  h->dp_stream_set = false;
}

int pinn_destroy(pinn_handle* h) {
  if (!h) return 0;
  DevGuard dev_guard(h->device);
  dp_release(h);
  cudaDeviceSynchronize();
#ifdef PINN_AB_BUILD
  cudaFree(h->wts);
#endif
  for (pinn_workspace& w : h->ws) ws_free(w);
  cudaFree(h->theta_dev); cudaFree(h->stage_dev);
  cudaFreeHost(h->out_pinned); cudaFreeHost(h->theta_pinned);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->s_main) cudaStreamDestroy(h->s_main);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  for (int c = 0; c < 4; c++) if (h->ev_chunk[c]) cudaEventDestroy(h->ev_chunk[c]);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  delete h;
  return 0;
}

const char* pinn_last_error(pinn_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
int64_t pinn_launch_count(pinn_handle* h) { return h ? h->launches : 0; }

int pinn_set_engine(pinn_handle* h, int engine) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (engine != PINN_ENGINE_FFMA && engine != PINN_ENGINE_TCGEN05) return fail(h, PINN_EINVAL, "pinn_set_engine: unknown engine");
#ifndef PINN_AB_BUILD
  if (engine != PINN_ENGINE_TCGEN05)
    return fail(h, PINN_ENOTSUP, "pinn_set_engine: this library carries the tcgen05 engine only (the FFMA engine exists in A/B builds, tools/build_ab.sh)");
#endif
  h->engine = engine;
  return 0;
}
int pinn_get_engine(pinn_handle* h) { return h ? h->engine : PINN_EINVAL; }

int pinn_host_timing(pinn_handle* h, double* out4) {
  if (!h || !out4) return PINN_EINVAL;
  memcpy(out4, h->host_us, sizeof(h->host_us));
  return 0;
}

int pinn_profile_begin(pinn_handle* h) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  h->profiling = true;
  h->ev_used = 0;
  return 0;
}

int pinn_profile_collect(pinn_handle* h, double* total_ms, int* launches) {
  if (!h || !total_ms || !launches) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  DevGuard dev_guard(h->device);
  double tot = 0.0;
  for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
    CU(h, cudaEventSynchronize(h->ev_pool[i + 1]));
    float ms = 0.0f;
    CU(h, cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]));
    tot += ms;
  }
  *total_ms = tot;
  *launches = (int)(h->ev_used / 2);
  h->profiling = false;
  h->ev_used = 0;
  return 0;
}

// Device-visible alias of a page-locked host pointer, or false when the memory is pageable / not host memory.
static bool map_pinned(const void* p, const void** dev) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
  *dev = a.devicePointer;
  return true;
}

// ---------------------------------------------------------------------------------------------
// data-parallel exchange over NVLink peer memory (include/pinn_b200.h: pinn_dp_*)
// ---------------------------------------------------------------------------------------------
int pinn_dp_init(pinn_handle* h, int rank, int world, void* ipc_handle_out) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world)
    return fail(h, PINN_EINVAL, "pinn_dp_init: need 0 <= rank < world <= 8");
  DevGuard dev_guard(h->device);
  dp_release(h);
  CU(h, cudaMalloc(&h->dp_buf, DP_BUFFER_BYTES));
  CU(h, cudaMemset(h->dp_buf, 0, DP_BUFFER_BYTES));
  CU(h, cudaDeviceSynchronize());
  h->dp.rank = rank;
  h->dp.world = world;
  h->dp.peer[rank] = h->dp_buf;
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == PINN_DP_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t ih;
    CU(h, cudaIpcGetMemHandle(&ih, h->dp_buf));
    memcpy(ipc_handle_out, &ih, sizeof(ih));
  }
  return 0;
}

int pinn_dp_connect(pinn_handle* h, const void* all_handles) {
  if (!h || !all_handles) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->dp_buf) return fail(h, PINN_EINVAL, "pinn_dp_connect: call pinn_dp_init first");
  DevGuard dev_guard(h->device);
  for (int r = 0; r < h->dp.world; r++) {
    if (r == h->dp.rank) continue;
    cudaIpcMemHandle_t ih;
    memcpy(&ih, (const char*)all_handles + (size_t)r * PINN_DP_HANDLE_BYTES, sizeof(ih));
    void* ptr = nullptr;
    CU(h, cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    h->dp.peer[r] = (unsigned char*)ptr;
    h->dp_opened[r] = true;
  }
  h->dp_on = true;
  return 0;
}

int pinn_dp_connect_local(pinn_handle* h, pinn_handle* const* peers) {
  if (!h || !peers) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->dp_buf) return fail(h, PINN_EINVAL, "pinn_dp_connect_local: call pinn_dp_init first");
  DevGuard dev_guard(h->device);
  for (int r = 0; r < h->dp.world; r++) {
    if (r == h->dp.rank) continue;
    pinn_handle* q = peers[r];
    if (!q || !q->dp_buf || q->dp.world != h->dp.world || q->dp.rank != r)
      return fail(h, PINN_EINVAL, "pinn_dp_connect_local: peer handle is not initialised for this rank/world");
    if (q->device != h->device) {
      int can = 0;
      CU(h, cudaDeviceCanAccessPeer(&can, h->device, q->device));
      if (!can) return fail(h, PINN_ENOTSUP, "pinn_dp_connect_local: no peer access between the two devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(q->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(h, (int)e, "cudaDeviceEnablePeerAccess");
    }
    h->dp.peer[r] = q->dp_buf;
  }
  h->dp_on = true;
  return 0;
}

int pinn_dp_enable(pinn_handle* h, int on) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (on) {
    for (int r = 0; r < h->dp.world; r++)
      if (!h->dp.peer[r]) return fail(h, PINN_EINVAL, "pinn_dp_enable: the exchange is not connected");
    if (h->dp.world < 1) return fail(h, PINN_EINVAL, "pinn_dp_enable: the exchange is not initialised");
  }
  h->dp_on = on != 0;
  return 0;
}

int pinn_dp_set_timeout(pinn_handle* h, double seconds) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!(seconds > 0.0) || seconds > 3600.0) return fail(h, PINN_EINVAL, "pinn_dp_set_timeout: need 0 < seconds <= 3600");
  if (h->dp.world < 1) return fail(h, PINN_EINVAL, "pinn_dp_set_timeout: the exchange is not initialised");
  h->dp.timeout_cycles = (long long)(seconds * 2.0e9);  // SM cycles at ~2 GHz: a bound, not a stopwatch
  return 0;
}

int pinn_dp_status(pinn_handle* h, int64_t* exchanges) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!h->dp_buf) return fail(h, PINN_EINVAL, "pinn_dp_status: the exchange is not initialised");
  DevGuard dev_guard(h->device);
  unsigned long long ctl[3];
  CU(h, cudaMemcpy(ctl, h->dp_buf + DP_ROWS_BYTES, sizeof(ctl), cudaMemcpyDeviceToHost));
  if (exchanges) *exchanges = (int64_t)ctl[0];
  if (ctl[2]) return fail(h, PINN_ETIMEDOUT, "pinn_dp: a peer did not deliver its partial sums within the time-out");
  return 0;
}

int pinn_dp_shutdown(pinn_handle* h) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  DevGuard dev_guard(h->device);
  CU(h, cudaDeviceSynchronize());
  dp_release(h);
  return 0;
}

static int variant_coef(int variant, VariantCoef* vc, int* nev) {
  if (variant == PINN_VARIANT_POC) { *vc = {1.0f, -0.5f, -1.0f, -1.0f}; *nev = 2; return 0; }
  if (variant == PINN_VARIANT_TRAINPY) { *vc = {2.0f, 1.0f, 1.0f, 1.0f}; *nev = 1; return 0; }
  return PINN_EINVAL;
}

static int grid_for(pinn_handle* h, long long n) {
  // one persistent CTA per SM; a CTA works on super-tiles of 128 points
  const long long want = (n + 127) / 128;
  return (int)(want < h->sm_count ? (want < 1 ? 1 : want) : h->sm_count);
}

// The step kernel builds its operand images from p.theta itself (A/B builds: the FFMA engine reads an image prepared by
// a small kernel).
static int enqueue_prep(pinn_handle* h, const float* theta, StepParams& p, cudaStream_t st) {
  p.theta = theta;
#ifdef PINN_AB_BUILD
  p.wts = h->wts;
  if (h->engine == PINN_ENGINE_TCGEN05) return 0;
  CU(h, launch_prep(theta, h->wts, st));
  h->launches++;
#else
  (void)st;
#endif
  return 0;
}

static cudaError_t launch_step_any(pinn_handle* h, int nev, bool train, const StepParams& p, int grid, cudaStream_t st) {
#ifdef PINN_AB_BUILD
  if (h->engine != PINN_ENGINE_TCGEN05) return launch_step(nev, train, p, grid, st);
#endif
  (void)h;
  return launch_step_tc(nev, train, p, grid, st);
}

// Enqueue the fused step kernel for points [first, first + cnt) of a batch; its per-CTA rows go to partial rows
// [row0, row0 + *rows).  The handle mutex is held by the caller.
static int enqueue_step_chunk(pinn_handle* h, pinn_workspace* ws, int nev, StepParams p, int64_t first, int64_t cnt, int row0,
                              int* rows, cudaStream_t st) {
  const size_t es = p.in_f64 ? 8 : 4;
  p.x = (const char*)p.x + first * es; p.y = (const char*)p.y + first * es;
  p.z = (const char*)p.z + first * es; p.R = (const char*)p.R + first * es;
  if (p.mask) p.mask += first;
  if (p.E_out) p.E_out += first;
  p.n = cnt;
  p.partials = ws->partials + (size_t)row0 * NPART;
  const int grid = grid_for(h, cnt);
  if (row0 + grid > ws->rows) return fail(h, PINN_EINVAL, "internal: partial-row workspace exceeded");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->profiling) {
    while (h->ev_pool.size() < h->ev_used + 2) {
      cudaEvent_t e;
      CU(h, cudaEventCreate(&e));
      h->ev_pool.push_back(e);
    }
    e0 = h->ev_pool[h->ev_used++];
    e1 = h->ev_pool[h->ev_used++];
    CU(h, cudaEventRecord(e0, st));
  }
  CU(h, launch_step_any(h, nev, true, p, grid, st));
  if (e1) CU(h, cudaEventRecord(e1, st));
  h->launches++;
  *rows = grid;
  return 0;
}

// The training evaluation on device buffers.  theta / weights: device pointers, or (tcgen05 engine, used by the *_host
// entry) NULL with theta_inline / weights_inline = HOST arrays that travel inside the kernel parameters.
}  // extern "C"

int loss_fwd_bwd_impl(pinn_handle* h, const LossCall& c, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(h->mu);
  StepParams p{};
  int nev = 0;
  if (variant_coef(c.variant, &p.vc, &nev)) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd: unknown variant");
  if (c.n <= 0) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd: n must be positive");
  if (!c.x || !c.y || !c.z || !c.R || (!c.theta && !c.theta_inline && !c.theta_tensors) || !c.sums || !c.dtheta)
    return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd: NULL pointer argument");
  if (c.in_dtype != PINN_F32 && c.in_dtype != PINN_F64) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd: bad in_dtype");
  DevGuard dev_guard(h->device);
  pinn_workspace* ws = ws_for(h, st, h->sm_count);
  if (!ws) return PINN_EINVAL;
  if (int rc = dp_serialize(h, st)) return rc;
  p.x = c.x; p.y = c.y; p.z = c.z; p.R = c.R; p.mask = c.mask; p.n = c.n; p.in_f64 = c.in_dtype == PINN_F64;
  p.bcut = c.bcutoff; p.partials = ws->partials; p.E_out = static_cast<float*>(c.E_out); p.E_f64 = c.E_f64;
  p.base_grads = (c.grad_mask & 0x003Fu) != 0;
  p.gate_grads = (c.grad_mask & 0xF000u) != 0;
  const double* weights = c.weights;
  if (c.theta_inline) {
    p.theta_inline = c.theta_inline;
  } else if (c.theta_tensors) {
#ifdef PINN_AB_BUILD
    if (h->engine != PINN_ENGINE_TCGEN05) return fail(h, PINN_ENOTSUP, "the FFMA engine takes a packed theta only");
#endif
    for (int k = 0; k < 16; k++) {
      if (!c.theta_tensors[k]) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd_tensors: NULL parameter tensor");
      p.theta_tensors[k] = c.theta_tensors[k];
    }
    p.theta_from_tensors = 1; p.tensors_f64 = c.tensors_f64; p.tensors_in_out = c.tensors_in_out;
  } else {
    if (int rc = enqueue_prep(h, c.theta, p, st)) return rc;
  }
  if (c.weights_inline) {
    p.weights_inline = c.weights_inline;
  } else if (!weights) {
    CU(h, launch_count(p, ws->counts, ws->weights, st));
    h->launches += 2;
    weights = ws->weights;
  }
  p.weights = weights;
  int grid = 0;
  int rc = enqueue_step_chunk(h, ws, nev, p, 0, c.n, 0, &grid, st);
  if (rc) return rc;
  CU(h, launch_reduce(ws->partials, grid, weights, c.weights_inline, c.grad_mask, c.dtheta, c.sums, p.E_out, c.n,
                      h->dp_on ? h->dp : DpArgs(), st, c.adam, c.adam_ticket, c.presample, c.E_f64, c.dtheta_in_out));
  h->launches++;
  return 0;
}

extern "C" {

int pinn_loss_fwd_bwd(pinn_handle* h, int variant, int64_t n, const void* x, const void* y, const void* z,
                      const void* R, int in_dtype, const uint8_t* mask, const float* theta, const double* weights,
                      uint32_t grad_mask, float bcutoff, double* sums, double* dtheta, float* E_out, void* stream) {
  if (!h) return PINN_EINVAL;
  if (!theta) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd: NULL pointer argument");
  LossCall c;
  c.variant = variant; c.n = n; c.x = x; c.y = y; c.z = z; c.R = R; c.in_dtype = in_dtype; c.mask = mask;
  c.theta = theta; c.weights = weights; c.grad_mask = grad_mask; c.bcutoff = bcutoff; c.sums = sums; c.dtheta = dtheta;
  c.E_out = E_out;
  return loss_fwd_bwd_impl(h, c, (cudaStream_t)stream);
}

int pinn_loss_fwd_bwd_tensors(pinn_handle* h, int variant, int64_t n, const void* x, const void* y, const void* z,
                              const void* R, int in_dtype, const uint8_t* mask, const void* const* params, int param_dtype,
                              int param_layout, const double* weights_host, uint32_t grad_mask, float bcutoff, double* sums,
                              double* dtheta, void* E_out, int E_dtype, void* stream) {
  if (!h) return PINN_EINVAL;
  if (!params) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd_tensors: NULL pointer argument");
  if ((param_dtype != PINN_F32 && param_dtype != PINN_F64) || (E_dtype != PINN_F32 && E_dtype != PINN_F64) ||
      (param_layout != PINN_VARIANT_POC && param_layout != PINN_VARIANT_TRAINPY))
    return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd_tensors: bad dtype / layout code");
  // canonical order of the 16 tensors; train.py keeps the gate (L1a..L2b) in front of the E-net (train.py:108-109)
  static const int kTrainpyToCanonical[16] = {0, 1, 2, 3, 4, 5, 12, 13, 14, 15, 6, 7, 8, 9, 10, 11};
  const void* canon[16];
  for (int k = 0; k < 16; k++) canon[param_layout == PINN_VARIANT_TRAINPY ? kTrainpyToCanonical[k] : k] = params[k];
  LossCall c;
  c.variant = variant; c.n = n; c.x = x; c.y = y; c.z = z; c.R = R; c.in_dtype = in_dtype; c.mask = mask;
  c.theta_tensors = canon; c.tensors_f64 = param_dtype == PINN_F64; c.tensors_in_out = param_layout == PINN_VARIANT_TRAINPY;
  c.weights_inline = weights_host; c.grad_mask = grad_mask; c.bcutoff = bcutoff; c.sums = sums; c.dtheta = dtheta;
  c.E_out = E_out; c.E_f64 = E_dtype == PINN_F64; c.dtheta_in_out = c.tensors_in_out;
  return loss_fwd_bwd_impl(h, c, (cudaStream_t)stream);
}

int pinn_mask_from_index_sets(pinn_handle* h, int64_t n, const int64_t* idx1, int64_t n1, const int64_t* idx2, int64_t n2,
                              uint8_t* mask, void* stream) {
  if (!h) return PINN_EINVAL;
  if (n <= 0 || n1 < 0 || n2 < 0 || !mask || (n1 > 0 && !idx1) || (n2 > 0 && !idx2) || ((uintptr_t)mask & 3u))
    return fail(h, PINN_EINVAL, "pinn_mask_from_index_sets: bad argument (mask must be 4-byte aligned with room for n rounded up to 4)");
  DevGuard dev_guard(h->device);
  CU(h, launch_mask_from_index_sets(reinterpret_cast<const long long*>(idx1), n1, reinterpret_cast<const long long*>(idx2), n2,
                                    mask, n, (cudaStream_t)stream));
  std::lock_guard<std::mutex> lk(h->mu);
  h->launches += 1;
  return 0;
}

int pinn_fields(pinn_handle* h, int variant, int64_t n, const void* x, const void* y, const void* z, const void* R,
                int in_dtype, const float* theta, float* psi, float* lap, float* hpsi, float* res, float* E,
                void* stream) {
  if (!h) return PINN_EINVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  StepParams p{};
  int nev = 0;
  if (variant_coef(variant, &p.vc, &nev)) return fail(h, PINN_EINVAL, "pinn_fields: unknown variant");
  if (n <= 0) return fail(h, PINN_EINVAL, "pinn_fields: n must be positive");
  if (!x || !y || !z || !R || !theta) return fail(h, PINN_EINVAL, "pinn_fields: NULL pointer argument");
  if (in_dtype != PINN_F32 && in_dtype != PINN_F64) return fail(h, PINN_EINVAL, "pinn_fields: bad in_dtype");
  DevGuard dev_guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  p.x = x; p.y = y; p.z = z; p.R = R; p.n = n; p.in_f64 = in_dtype == PINN_F64;
  p.psi = psi; p.lap = lap; p.hpsi = hpsi; p.res = res; p.E_out = E;
  if (int rc = enqueue_prep(h, theta, p, st)) return rc;
  CU(h, launch_step_any(h, nev, false, p, grid_for(h, n), st));
  h->launches++;
  return 0;
}

int pinn_loss_fwd_bwd_host(pinn_handle* h, int variant, int64_t n, const void* x, const void* y, const void* z,
                           const void* R, int in_dtype, const uint8_t* mask, const double* theta_host,
                           const double* weights_host, uint32_t grad_mask, float bcutoff, double* sums_host,
                           double* dtheta_host, float* E_out_host) {
  if (!h) return PINN_EINVAL;
  if (n <= 0) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd_host: n must be positive");
  if (!x || !y || !z || !R || !theta_host || !sums_host || !dtheta_host)
    return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd_host: NULL pointer argument");
  if (in_dtype != PINN_F32 && in_dtype != PINN_F64)
    return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd_host: bad in_dtype");
  std::lock_guard<std::mutex> host_lock(h->host_mu);
  const auto t_enter = std::chrono::steady_clock::now();
  DevGuard dev_guard(h->device);
  const size_t es = in_dtype == PINN_F64 ? 8 : 4;
  // Page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) inputs are read by the kernel in place: the
  // step kernel requests the coordinates of its next super-tile one tile ahead, which hides the PCIe latency, and
  // the 16 B/point crossing the bus overlap the whole kernel instead of preceding it.  Pageable inputs are staged.
  const void* mapped[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const bool zero_copy = h->host_zero_copy && map_pinned(x, &mapped[0]) && map_pinned(y, &mapped[1]) &&
                         map_pinned(z, &mapped[2]) && map_pinned(R, &mapped[3]) && (!mask || map_pinned(mask, &mapped[4]));
  const size_t col = zero_copy ? 0 : ((size_t)n * es + 255) & ~(size_t)255;
  const size_t mcol = zero_copy ? 0 : ((size_t)n + 255) & ~(size_t)255;
  const size_t ecol = ((size_t)n * 4 + 255) & ~(size_t)255;
  const size_t need = 4 * col + mcol + ecol;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (need > h->stage_bytes) {  // grows geometrically; steady-state steps do not allocate
      CU(h, cudaStreamSynchronize(h->s_main));
      if (h->stage_dev) CU(h, cudaFree(h->stage_dev));
      h->stage_dev = nullptr;
      const size_t cap = need + need / 4;
      CU(h, cudaMalloc(&h->stage_dev, cap));
      h->stage_bytes = cap;
    }
  }
  char* base = (char*)h->stage_dev;
  cudaStream_t st = h->s_main;
  for (int i = 0; i < NTHETA; i++) h->theta_pinned[i] = (float)theta_host[i];
  const double* wdev = nullptr;
  if (weights_host) {
    memcpy(h->weights_pinned, weights_host, 3 * sizeof(double));
    wdev = h->weights_dev;
  }
  // tcgen05 engine with known weights: theta and the weights ride in the kernel parameters, nothing is uploaded first
  const bool inline_params = weights_host && h->engine == PINN_ENGINE_TCGEN05 && h->host_inline_params;
  if (!inline_params)
    CU(h, cudaMemcpyAsync(h->theta_dev, h->theta_pinned, weights_host ? HOST_IN_BYTES : NTHETA * sizeof(float),
                          cudaMemcpyHostToDevice, st));
  const float* th_dev = inline_params ? nullptr : h->theta_dev;
  const float* th_inl = inline_params ? h->theta_pinned : nullptr;
  const double* w_inl = inline_params ? h->weights_pinned : nullptr;
  if (inline_params) wdev = nullptr;
  // the 8 sums and 1521 gradients are written by the reduction kernel straight into mapped page-locked memory
  double* outp = h->out_mapped;
  uint8_t* mdev = mask ? (uint8_t*)(base + 4 * col) : nullptr;
  float* edev = (float*)(base + 4 * col + mcol);
  // With the weights known up front the batch is processed in up to 4 chunks so that the copy of chunk k+1 (copy stream)
  // overlaps the kernel of chunk k.  Chunks are whole rounds of super-tiles (sm_count x 128 points) so that no launch
  // ends on a partial wave; the first chunk is short (2 rounds) to start computing early.
  // Without weights the set sizes must be counted over the whole batch first: one chunk.
  const int64_t unit = (int64_t)h->sm_count * 128;
  const int64_t rounds = (n + unit - 1) / unit;
  int nchunk = 1;
  int64_t first[4] = {0, 0, 0, 0}, cnt[4] = {n, 0, 0, 0};
  if (weights_host && rounds >= 8) {
    nchunk = 4;
    const int64_t rest = rounds - 2;
    const int64_t r[4] = {2, rest / 3, rest / 3, rest - 2 * (rest / 3)};
    int64_t at = 0;
    for (int c = 0; c < 4; c++) {
      first[c] = at;
      cnt[c] = (c == 3) ? n - at : r[c] * unit;
      at += cnt[c];
    }
  }
  if (zero_copy) nchunk = 1;
  const void* src[4] = {x, y, z, R};
  for (int c = 0; c < nchunk && !zero_copy; c++) {
    cudaStream_t sc = nchunk == 1 ? st : h->s_copy;
    for (int k = 0; k < 4; k++)
      CU(h, cudaMemcpyAsync(base + k * col + first[c] * es, (const char*)src[k] + first[c] * es, (size_t)cnt[c] * es,
                            cudaMemcpyHostToDevice, sc));
    if (mask) CU(h, cudaMemcpyAsync(mdev + first[c], mask + first[c], (size_t)cnt[c], cudaMemcpyHostToDevice, sc));
    if (nchunk > 1) CU(h, cudaEventRecord(h->ev_chunk[c], sc));
  }
  if (zero_copy) {
    LossCall c;
    c.variant = variant; c.n = n; c.x = mapped[0]; c.y = mapped[1]; c.z = mapped[2]; c.R = mapped[3]; c.in_dtype = in_dtype;
    c.mask = (const uint8_t*)mapped[4]; c.theta = th_dev; c.weights = wdev; c.theta_inline = th_inl; c.weights_inline = w_inl;
    c.grad_mask = grad_mask; c.bcutoff = bcutoff; c.sums = outp; c.dtheta = outp + 8; c.E_out = edev;
    int rc = loss_fwd_bwd_impl(h, c, st);
    if (rc) return rc;
  } else if (nchunk == 1) {
    LossCall c;
    c.variant = variant; c.n = n; c.x = base; c.y = base + col; c.z = base + 2 * col; c.R = base + 3 * col; c.in_dtype = in_dtype;
    c.mask = mdev; c.theta = th_dev; c.weights = wdev; c.theta_inline = th_inl; c.weights_inline = w_inl;
    c.grad_mask = grad_mask; c.bcutoff = bcutoff; c.sums = outp; c.dtheta = outp + 8; c.E_out = edev;
    int rc = loss_fwd_bwd_impl(h, c, st);
    if (rc) return rc;
  } else {
    std::lock_guard<std::mutex> lk(h->mu);
    StepParams p{};
    int nev = 0;
    if (variant_coef(variant, &p.vc, &nev)) return fail(h, PINN_EINVAL, "pinn_loss_fwd_bwd_host: unknown variant");
    pinn_workspace* ws = ws_for(h, st, 4 * h->sm_count);
    if (!ws) return PINN_EINVAL;
    if (int rc = dp_serialize(h, st)) return rc;
    p.x = base; p.y = base + col; p.z = base + 2 * col; p.R = base + 3 * col; p.mask = mdev;
    p.in_f64 = in_dtype == PINN_F64; p.bcut = bcutoff; p.E_out = edev; p.weights = wdev;
    p.base_grads = (grad_mask & 0x003Fu) != 0;
    p.gate_grads = (grad_mask & 0xF000u) != 0;
    if (inline_params) { p.theta_inline = th_inl; p.weights_inline = w_inl; }
    else if (int rc = enqueue_prep(h, h->theta_dev, p, st)) return rc;
    int rows = 0;
    for (int c = 0; c < nchunk; c++) {
      CU(h, cudaStreamWaitEvent(st, h->ev_chunk[c], 0));
      int r = 0;
      int rc = enqueue_step_chunk(h, ws, nev, p, first[c], cnt[c], rows, &r, st);
      if (rc) return rc;
      rows += r;
    }
    CU(h, launch_reduce(ws->partials, rows, wdev, w_inl, grad_mask, outp + 8, outp, edev, n, h->dp_on ? h->dp : DpArgs(), st));
    h->launches++;
  }
  if (E_out_host) CU(h, cudaMemcpyAsync(E_out_host, edev, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  const auto t_submitted = std::chrono::steady_clock::now();
  CU(h, cudaStreamSynchronize(st));
  const auto t_done = std::chrono::steady_clock::now();
  memcpy(sums_host, h->out_pinned, 8 * sizeof(double));
  memcpy(dtheta_host, h->out_pinned + 8, NTHETA * sizeof(double));
  const auto t_exit = std::chrono::steady_clock::now();
  auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  h->host_us[0] = us(t_enter, t_submitted); h->host_us[1] = us(t_submitted, t_done); h->host_us[2] = us(t_done, t_exit);
  h->host_us[3] = us(t_enter, t_exit);
  return 0;
}

}  // extern "C"
