// Device-side pieces of the training loop around the fused step kernel:
//   * collocation sampler + clamp + boundary sets (train.py:26-39; poc/main.py:124-156, 390-393), Philox4x32-10,
//     one counter per point, so a batch never round-trips through the host;
//   * fused Adam over the 1521 parameters in float64 with the reference's best-model bookkeeping
//     (torch.optim.Adam semantics; train.py:58-60, 65-69; poc/main.py:403-417) and the per-step history;
//   * E(R), dE/dR, d2E/dR2 and the gate along an R grid (energy.py:26-33; poc/main.py:164-176, 1324-1332).
// Everything is launched on a caller-given stream and is CUDA-graph capturable (no host decisions inside a step).
#include "pinn_device.cuh"
#include "pinn_train.h"

namespace pinn {

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1) -> 4 x 32 random bits
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }  // [0,1), 24 bits

// weights {1/n, 1/|set1|, 1/|set2|} of the reference's means (empty set -> inf -> NaN loss, as in the reference),
// and the batch counter moves on.  Run by one warp of the block of sample_kernel that finishes last.  Data-parallel runs (dp.world > 1) first add the set sizes of all ranks:
// lane r stores this rank's two counts into peer r's exchange buffer as {count, step} words and polls its own buffer
// for peer r's (same protocol as the gradient sum in reduce_partials_kernel), lane 0 adds them in rank order.
__device__ __forceinline__ void sample_finish(const SampleParams& s) {
  __shared__ unsigned long long c[DP_MAX_WORLD][2];
  const DpArgs& dp = s.dp;
  const int lane = threadIdx.x;
  const long long n = s.n;
  double* w = s.weights;
  unsigned long long c1 = atomicAdd(&s.counts[0], 0ull), c2 = atomicAdd(&s.counts[1], 0ull);  // L2 values of the other blocks' atomics
  long long ntot = n;
  if (dp.world > 1) {
    unsigned char* own = dp.peer[dp.rank];
    unsigned long long* ctl = reinterpret_cast<unsigned long long*>(own + DP_ROWS_BYTES);
    const unsigned long long step64 = ld_acquire_sys(&ctl[3]) + 1;
    const unsigned int step = (unsigned int)step64;
    const size_t slot = (size_t)(step64 & 1ull) * DP_MAX_WORLD;
    if (lane < dp.world) {
      const int r = lane;
      if (r == dp.rank) {
        c[r][0] = c1; c[r][1] = c2;
      } else {
        unsigned int* dst = reinterpret_cast<unsigned int*>(dp.peer[r] + DP_ROWS_BYTES + DP_CTL_BYTES) + (slot + dp.rank) * 4;
        st_relaxed_sys_v2(dst, (unsigned int)c1, step);
        st_relaxed_sys_v2(dst + 2, (unsigned int)c2, step);
        const unsigned int* src = reinterpret_cast<const unsigned int*>(own + DP_ROWS_BYTES + DP_CTL_BYTES) + (slot + r) * 4;
        const long long t0 = clock64();
        uint2 a, b;
        for (;;) {
          a = ld_relaxed_sys_v2(src);
          b = ld_relaxed_sys_v2(src + 2);
          if (a.y == step && b.y == step) break;
          if (clock64() - t0 > 6000000000ll) { ctl[2] = 1ull; break; }
        }
        c[r][0] = a.x; c[r][1] = b.x;
      }
    }
    __syncwarp();
    if (lane == 0) {
      c1 = 0; c2 = 0;
      for (int r = 0; r < dp.world; r++) { c1 += c[r][0]; c2 += c[r][1]; }
      ntot = n * dp.world;
      st_release_sys(&ctl[3], step64);
    }
  }
  if (lane == 0) {
    w[0] = 1.0 / (double)ntot;
    w[1] = 1.0 / (double)c1;
    w[2] = 1.0 / (double)c2;
    *s.batch_counter += 1ull;
    *s.ticket = 0ull;
    if (s.reset_counts) { s.counts[0] = 0ull; s.counts[1] = 0ull; }
  }
}

// One thread per point.  Point i of batch b uses counter (i_lo, i_hi, b_lo, b_hi) and key = seed; its four words give
// x, y, z, R.  Clamp and sets follow the reference literally: both tests use the radii of the UN-clamped point
// (train.py:32-35), the clamp writes the VALUE `cutoff` into x, and the sets are taken after it (train.py:36-39).
__global__ void __launch_bounds__(256) sample_kernel(const SampleParams s) {
  pdl_wait();
  const unsigned long long batch = *s.batch_counter;
  unsigned c1 = 0, c2 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    const unsigned long long gi = (unsigned long long)(i + s.index_offset);  // index of the point in the global batch
    philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)batch, (uint32_t)(batch >> 32),
                  (uint32_t)s.seed, (uint32_t)(s.seed >> 32), r);
    float x = fmaf(s.xR - s.xL, u01(r[0]), s.xL);
    const float y = fmaf(s.yR - s.yL, u01(r[1]), s.yL);
    const float z = fmaf(s.zR - s.zL, u01(r[2]), s.zL);
    const float R = fmaf(s.RR - s.RL, u01(r[3]), s.RL);
    const float yz = fmaf(y, y, z * z);
    const float c2cut = s.cutoff * s.cutoff;
    const bool near1 = fmaf(x - R, x - R, yz) < c2cut, near2 = fmaf(x + R, x + R, yz) < c2cut;
    if (near1 || near2) x = s.cutoff;
    const float b2 = s.bcutoff * s.bcutoff;
    const unsigned m1 = fmaf(x - R, x - R, yz) >= b2, m2 = fmaf(x + R, x + R, yz) >= b2;
    s.x[i] = x; s.y[i] = y; s.z[i] = z; s.R[i] = R;
    s.mask[i] = (uint8_t)(m1 | (m2 << 1));
    c1 += m1; c2 += m2;
  }
  c1 = __reduce_add_sync(0xffffffffu, c1);
  c2 = __reduce_add_sync(0xffffffffu, c2);
  __shared__ unsigned sh1[8], sh2[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh1[w] = c1; sh2[w] = c2; }
  __syncthreads();
  __shared__ int last_block;
  if (threadIdx.x == 0) {
    unsigned t1 = 0, t2 = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) { t1 += sh1[k]; t2 += sh2[k]; }
    atomicAdd(&s.counts[0], (unsigned long long)t1);  // integer atomics: order-independent result
    atomicAdd(&s.counts[1], (unsigned long long)t2);
    __threadfence();
    last_block = atomicAdd(s.ticket, 1ull) == (unsigned long long)gridDim.x - 1;
  }
  __syncthreads();
  if (last_block && threadIdx.x < 32) sample_finish(s);  // every other block has added its counts and read the batch index
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

cudaError_t launch_enet_curve(const float* theta, const double* R, int n, double* E, double* dE, double* d2E, double* gate,
                              cudaStream_t st) {
  enet_curve_kernel<<<(n + 127) / 128, 128, 0, st>>>(theta, R, n, E, dE, d2E, gate);
  return cudaGetLastError();
}

}  // namespace pinn
