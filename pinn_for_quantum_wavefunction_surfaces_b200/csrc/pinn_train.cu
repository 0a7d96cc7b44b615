// Device-side pieces of the training loop around the fused step kernel:
//   * collocation sampler + clamp + boundary sets (train.py:26-39; poc/main.py:124-156, 390-393), Philox4x32-10,
//     one counter per point, so a batch never round-trips through the host;
//   * fused Adam over the 1521 parameters in float64 with the reference's best-model bookkeeping
//     (torch.optim.Adam semantics; train.py:58-60, 65-69; poc/main.py:403-417) and the per-step history;
//   * E(R), dE/dR, d2E/dR2 and the gate along an R grid (energy.py:26-33; poc/main.py:164-176, 1324-1332).
// Everything is launched on a caller-given stream and is CUDA-graph capturable (no host decisions inside a step).
#include "pinn_device.cuh"
#include "pinn_sample.cuh"
#include "pinn_train.h"

namespace pinn {

__global__ void __launch_bounds__(256) sample_kernel(const SampleParams s) {
  pdl_wait();
  sample_block(s, blockIdx.x, gridDim.x);
}

cudaError_t launch_sample(const SampleParams& s, bool zero_counts, cudaStream_t st) {
  if (zero_counts) {
    cudaError_t e = cudaMemsetAsync(s.counts, 0, 2 * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
  }
  long long blocks = (s.n + 255) / 256;
  if (blocks > 148 * 2) blocks = 148 * 2;  // two waves of 256 threads per SM: ~3.5 points per thread at 2^18, few blocks to launch and to count
  return launch_pdl(sample_kernel, dim3((unsigned)blocks), dim3(256), 0, st, s);
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam: lerp first moment, bias corrections, eps outside the square root; no weight decay, no amsgrad)
// in float64, the reference's parameter dtype, plus best-model bookkeeping and history
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) adam_kernel(const AdamParams a) {
  __shared__ int take_best_s;
  pdl_wait();
  const unsigned long long t = *a.step;  // optimizer steps done so far = index tt of this step in the reference loops
  if (threadIdx.x == 0) take_best_s = adam_take_best(a, t, a.sums[0]) ? 1 : 0;
  __syncthreads();
  const bool take_best = take_best_s != 0;
  const AdamCoef c = adam_coef(a, t);
  for (int i = threadIdx.x; i < NTHETA; i += blockDim.x) adam_update_entry(a, c, i, a.grad[i], take_best);
  __syncthreads();
  if (threadIdx.x == 0) adam_bookkeeping(a, t, a.sums, take_best);
}

cudaError_t launch_adam(const AdamParams& a, cudaStream_t st) {
  return launch_pdl(adam_kernel, dim3(1), dim3(1024), 0, st, a);
}

// ---------------------------------------------------------------------------------------------
// E(R), dE/dR, d2E/dR2 (forward-mode through the 1-32-32-1 E-net) and the gate g(R); one thread per R
// ---------------------------------------------------------------------------------------------
__global__ void enet_curve_kernel(const float* __restrict__ th, const double* __restrict__ R, int n, double* __restrict__ E,
                                  double* __restrict__ dE, double* __restrict__ d2E, double* __restrict__ gate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = R[i];
  double e1[NE], e1p[NE], e1pp[NE];
  for (int k = 0; k < NE; k++) {
    const double w = th[O_WE1 + k], u = w * r + (double)th[O_BE1 + k];
    const double s = 1.0 / (1.0 + exp(-u)), sp = s * (1.0 - s), spp = sp * (1.0 - 2.0 * s);
    e1[k] = s; e1p[k] = sp * w; e1pp[k] = spp * w * w;
  }
  double Ev = th[O_BE], Ep = 0.0, Epp = 0.0;
  for (int j = 0; j < NE; j++) {
    double u = th[O_BE2 + j], up = 0.0, upp = 0.0;
    for (int k = 0; k < NE; k++) {
      const double w = th[O_WE2 + j * NE + k];
      u += w * e1[k]; up += w * e1p[k]; upp += w * e1pp[k];
    }
    const double s = 1.0 / (1.0 + exp(-u)), sp = s * (1.0 - s), spp = sp * (1.0 - 2.0 * s);
    const double wE = th[O_WE + j];
    Ev += wE * s; Ep += wE * sp * up; Epp += wE * (spp * up * up + sp * upp);
  }
  if (E) E[i] = Ev;
  if (dE) dE[i] = Ep;
  if (d2E) d2E[i] = Epp;
  if (gate) {
    double g = th[O_BG];
    for (int k = 0; k < NL; k++) {
      const double u = (double)th[O_WGL + k] * r + (double)th[O_BGL + k];
      g += (double)th[O_WG + k] / (1.0 + exp(-u));
    }
    gate[i] = g;
  }
}

// ---------------------------------------------------------------------------------------------
// Roofline denominator, measured where and when the benchmark runs: a register-resident FFMA loop (16 independent
// accumulators per thread, 256 threads x 8 blocks per SM - the loop of tools/microbench/pipes.cu that reaches the
// FP32 pipe's rate).  bench.py reports the kernel against THIS number (MEASURED_PEAKS.json has no FP32 entry).
// ---------------------------------------------------------------------------------------------
constexpr int FFMA_PEAK_ITERS = 4096, FFMA_PEAK_ACC = 16;
// CLK = false: the timed kernel (nothing but the loop and one store per thread, exactly the microbenchmark's); CLK = true:
// one extra launch in which block 0 also times itself (SM cycles, nanoseconds) - the self-timing variant measured 4.6 %
// slower as a whole (0.577 vs 0.552 ms on one box) and is therefore not the one the rate is taken from.
template <bool CLK>
__global__ void ffma_peak_kernel(float* __restrict__ out, const float* __restrict__ in, double* __restrict__ clk) {
  long long clk0 = 0;
  unsigned long long ns0 = 0;
  if (CLK && blockIdx.x == 0 && threadIdx.x == 0) {
    clk0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
  }
  float a[FFMA_PEAK_ACC];
  const float b = in[0], c = in[1];
#pragma unroll
  for (int i = 0; i < FFMA_PEAK_ACC; i++) a[i] = in[2 + i] + threadIdx.x;
  for (int it = 0; it < FFMA_PEAK_ITERS; it++) {
#pragma unroll
    for (int i = 0; i < FFMA_PEAK_ACC; i++) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < FFMA_PEAK_ACC; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (CLK && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    clk[0] = (double)(clock64() - clk0);
    clk[1] = (double)(ns1 - ns0);
  }
}
cudaError_t measure_fp32_peak(int sm_count, cudaStream_t st, double* fma_per_s, double* ms_best, double* sm_mhz) {
  float* out = nullptr;
  const size_t nthreads = (size_t)sm_count * 8 * 256;
  cudaError_t e = cudaMalloc(&out, nthreads * sizeof(float) + 64 * sizeof(float) + 2 * sizeof(double));
  if (e != cudaSuccess) return e;
  float* in = out + nthreads;
  double* clk = reinterpret_cast<double*>(in + 64);
  float hin[64];
  for (int i = 0; i < 64; i++) hin[i] = 0.5f + 0.001f * i;
  hin[0] = 0.999f; hin[1] = 1e-3f;
  e = cudaMemcpyAsync(in, hin, sizeof(hin), cudaMemcpyHostToDevice, st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const dim3 grid(sm_count * 8), block(256);
  for (int i = 0; i < 3; i++) ffma_peak_kernel<false><<<grid, block, 0, st>>>(out, in, nullptr);
  float best = 1e30f;
  for (int r = 0; r < 8 && e == cudaSuccess; r++) {  // best of 8 x 10 launches (~45 ms in all)
    cudaEventRecord(e0, st);
    for (int i = 0; i < 10; i++) ffma_peak_kernel<false><<<grid, block, 0, st>>>(out, in, nullptr);
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10.0f;
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  double hclk[2] = {0.0, 1.0};
  if (e == cudaSuccess) {
    ffma_peak_kernel<true><<<grid, block, 0, st>>>(out, in, clk);
    e = cudaMemcpyAsync(hclk, clk, sizeof(hclk), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  cudaFree(out);
  if (e != cudaSuccess) return e;
  *ms_best = best;
  *sm_mhz = hclk[1] > 0.0 ? hclk[0] / hclk[1] * 1e3 : 0.0;  // effective SM clock of the self-timing launch (block 0)
  *fma_per_s = (double)FFMA_PEAK_ITERS * FFMA_PEAK_ACC * 256.0 * 8.0 * sm_count / (best * 1e-3);
  return cudaSuccess;
}

cudaError_t launch_enet_curve(const float* theta, const double* R, int n, double* E, double* dE, double* d2E, double* gate,
                              cudaStream_t st) {
  enet_curve_kernel<<<(n + 127) / 128, 128, 0, st>>>(theta, R, n, E, dE, d2E, gate);
  return cudaGetLastError();
}

}  // namespace pinn
