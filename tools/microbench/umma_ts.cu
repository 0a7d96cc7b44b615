// tcgen05 TS-mode TF32 validation + timing for the skinny mat-vecs of the PINN kernel (B200, sm_100a).
//   D[128 x N] (TMEM) = A[128 x K] (TMEM, written with tcgen05.st, lane = row) * B[N x K]^T (smem, K-major, no swizzle)
// Answers three questions before the product kernel relies on them:
//   1. are the shared-memory / instruction descriptors right (single-pass TF32 result vs CPU)?
//   2. does the tensor core truncate or round the fp32 operand to TF32, and is 3xTF32 (hi/lo split) fp32-accurate?
//   3. what does one round trip (tcgen05.st -> mma -> commit -> mbarrier -> tcgen05.ld) cost, and the MMA issue rate at N=16/32?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ts umma_ts.cu && ./umma_ts
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no-swizzle canonical layout: 8-row x 16-byte core matrices; SBO between 8-row groups, LBO between 16-byte K chunks
__host__ __device__ constexpr int b_off_floats(int n, int k, int N) { return (k / 4) * (32 * (N / 8)) + (n / 8) * 32 + (n % 8) * 4 + (k % 4); }

__device__ __forceinline__ uint64_t make_bdesc(uint32_t saddr, int N) {
  const uint32_t lbo = 128u * (N / 8), sbo = 128u;
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) /*D=f32*/ | (2u << 7) /*A=tf32*/ | (2u << 10) /*B=tf32*/ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
               "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
#define TC_WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")
#define TC_WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define TC_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define TC_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")

// mode 0: single pass with the raw fp32 operands; mode 1: 3xTF32
// A: [128][K], B: [N][K] row-major in global; D: [128][N].  Timing: `reps` round trips, cycles -> out_cycles[0..1]
template <int N, int K>
__global__ void __launch_bounds__(128, 1) umma_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                     float* __restrict__ D1, float* __restrict__ D3, int reps,
                                                     long long* out_cycles) {
  __shared__ __align__(128) float Bhi[N * K], Blo[N * K], Braw[N * K];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const float x = B[i];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    Braw[b_off_floats(n, k, N)] = x;
    Bhi[b_off_floats(n, k, N)] = hi;
    Blo[b_off_floats(n, k, N)] = x - hi;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  TC_FENCE_BEFORE();
  __syncthreads();
  TC_FENCE_AFTER();
  const uint32_t tb = tmem_base_s;
  const uint32_t lane_base = tb + ((uint32_t)(warp * 32) << 16);
  // columns: A raw/hi [0,K), A lo [K,2K), D [2K, 2K+N)
  const uint32_t cA = 0, cAlo = K, cD = 2 * K;
  constexpr uint32_t idesc = make_idesc(128, N);
  uint32_t parity = 0;

  float a[K];
  for (int k = 0; k < K; k++) a[k] = A[tid * K + k];

  for (int mode = 0; mode < 2; mode++) {
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t v[16], l[16];
      for (int i = 0; i < 16; i++) {
        const float x = a[k0 + i];
        if (mode == 0) { v[i] = __float_as_uint(x); l[i] = 0; }
        else {
          const uint32_t h = __float_as_uint(x) & 0xffffe000u;
          v[i] = h; l[i] = __float_as_uint(x - __uint_as_float(h));
        }
      }
      st16(lane_base + cA + k0, v);
      st16(lane_base + cAlo + k0, l);
    }
    TC_WAIT_ST();
    TC_FENCE_BEFORE();
    __syncthreads();
    if (tid == 0) {
      TC_FENCE_AFTER();
      uint32_t acc = 0;
      for (int ks = 0; ks < K / 8; ks++) {
        const uint32_t koff = ks * 2 * 128 * (N / 8);  // two 16-byte K chunks per step
        if (mode == 0) {
          mma_ts(tb + cD, tb + cA + ks * 8, make_bdesc(smem_u32(Braw) + koff, N), idesc, acc); acc = 1;
        } else {
          mma_ts(tb + cD, tb + cAlo + ks * 8, make_bdesc(smem_u32(Bhi) + koff, N), idesc, acc); acc = 1;
          mma_ts(tb + cD, tb + cA + ks * 8, make_bdesc(smem_u32(Blo) + koff, N), idesc, acc);
          mma_ts(tb + cD, tb + cA + ks * 8, make_bdesc(smem_u32(Bhi) + koff, N), idesc, acc);
        }
      }
      commit(smem_u32(&mbar));
    }
    mbar_wait(smem_u32(&mbar), parity); parity ^= 1;
    TC_FENCE_AFTER();
    float* Dout = mode == 0 ? D1 : D3;
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t v[16];
      ld16(lane_base + cD + n0, v);
      TC_WAIT_LD();
      if (blockIdx.x == 0)
        for (int i = 0; i < 16; i++) Dout[tid * N + n0 + i] = __uint_as_float(v[i]);
    }
    TC_FENCE_BEFORE();
    __syncthreads();
  }

  // ---- timing 1: full round trips (3xTF32) ----
  float sink = 0.0f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; r++) {
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t v[16], l[16];
      for (int i = 0; i < 16; i++) {
        const float x = a[k0 + i] + sink;
        const uint32_t h = __float_as_uint(x) & 0xffffe000u;
        v[i] = h; l[i] = __float_as_uint(x - __uint_as_float(h));
      }
      st16(lane_base + cA + k0, v);
      st16(lane_base + cAlo + k0, l);
    }
    TC_WAIT_ST();
    TC_FENCE_BEFORE();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid == 0) {
      TC_FENCE_AFTER();
      uint32_t acc = 0;
      for (int ks = 0; ks < K / 8; ks++) {
        const uint32_t koff = ks * 2 * 128 * (N / 8);
        mma_ts(tb + cD, tb + cAlo + ks * 8, make_bdesc(smem_u32(Bhi) + koff, N), idesc, acc); acc = 1;
        mma_ts(tb + cD, tb + cA + ks * 8, make_bdesc(smem_u32(Blo) + koff, N), idesc, acc);
        mma_ts(tb + cD, tb + cA + ks * 8, make_bdesc(smem_u32(Bhi) + koff, N), idesc, acc);
      }
      commit(smem_u32(&mbar));
    }
    mbar_wait(smem_u32(&mbar), parity); parity ^= 1;
    TC_FENCE_AFTER();
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t v[16];
      ld16(lane_base + cD + n0, v);
      TC_WAIT_LD();
      sink += __uint_as_float(v[0]) * 1e-30f;
    }
  }
  long long t1 = clock64();
  // ---- timing 2: MMA issue rate: 64 back-to-back MMAs per commit ----
  //   variant 0: all accumulate into one D tile; 1: round-robin over 4 D tiles (if they fit);
  //   2: two issuing threads (warps 0 and 1), 32 MMAs each, different D tiles
  long long tv[3];
  for (int variant = 0; variant < 3; variant++) {
    __syncthreads();
    long long t2 = clock64();
    for (int r = 0; r < reps; r++) {
      const uint64_t bd = make_bdesc(smem_u32(Bhi), N);
      if (variant < 2) {
        if (tid == 0) {
          for (int i = 0; i < 64; i++) mma_ts(tb + cD + ((variant == 1 && 4 * N + cD <= 512) ? (i & 3) * N : 0), tb + cA, bd, idesc, 1);
          commit(smem_u32(&mbar));
        }
      } else {
        if (tid == 32) {
          for (int i = 0; i < 32; i++) mma_ts(tb + cD + ((2 * N + cD <= 512) ? N : 0), tb + cA, bd, idesc, 1);
        }
        asm volatile("bar.sync 2, 64;" ::: "memory");  // warps 0,1 only
        if (tid == 0) {
          for (int i = 0; i < 32; i++) mma_ts(tb + cD, tb + cA, bd, idesc, 1);
          commit(smem_u32(&mbar));
        }
      }
      mbar_wait(smem_u32(&mbar), parity); parity ^= 1;
    }
    tv[variant] = clock64() - t2;
  }
  if (blockIdx.x == 0 && tid == 0) {
    out_cycles[0] = t1 - t0;
    out_cycles[1] = tv[0]; out_cycles[2] = tv[1]; out_cycles[3] = tv[2];
  }
  if (sink == 123.456f) D1[0] = sink;
  TC_FENCE_BEFORE();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }
static float tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

template <int N, int K>
static void run(int grid) {
  std::vector<float> A(128 * K), B(N * K), D1(128 * N), D3(128 * N);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.0f - 1.0f;
  float *dA, *dB, *dD1, *dD3; long long* dcyc;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD1, D1.size() * 4)); CK(cudaMalloc(&dD3, D3.size() * 4)); CK(cudaMalloc(&dcyc, 32));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const int reps = 2000;
  umma_kernel<N, K><<<grid, 128>>>(dA, dB, dD1, dD3, reps, dcyc);
  CK(cudaDeviceSynchronize());
  long long cyc[4];
  CK(cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dcyc, 32, cudaMemcpyDeviceToHost));
  double e_exact1 = 0, e_trunc = 0, e_rna = 0, e3 = 0, ref_max = 0;
  for (int m = 0; m < 128; m++)
    for (int n = 0; n < N; n++) {
      double ex = 0, tr = 0, rn = 0;
      for (int k = 0; k < K; k++) {
        ex += (double)A[m * K + k] * B[n * K + k];
        tr += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(B[n * K + k]);
        rn += (double)tf32_rna(A[m * K + k]) * tf32_rna(B[n * K + k]);
      }
      ref_max = fmax(ref_max, fabs(ex));
      e_exact1 = fmax(e_exact1, fabs(D1[m * N + n] - ex));
      e_trunc = fmax(e_trunc, fabs(D1[m * N + n] - tr));
      e_rna = fmax(e_rna, fabs(D1[m * N + n] - rn));
      e3 = fmax(e3, fabs(D3[m * N + n] - ex));
    }
  printf("{\"bench\":\"umma_ts_tf32\",\"M\":128,\"N\":%d,\"K\":%d,\"grid\":%d,\"ref_max\":%.4f,"
         "\"err_1pass_vs_exact\":%.3e,\"err_1pass_vs_trunc_model\":%.3e,\"err_1pass_vs_rna_model\":%.3e,"
         "\"err_3xtf32_vs_exact\":%.3e,\"roundtrip_cycles\":%.1f,\"mma_cycles_each_sameD\":%.2f,\"mma_cycles_each_4D\":%.2f,\"mma_cycles_each_2threads\":%.2f}\n",
         N, K, grid, ref_max, e_exact1, e_trunc, e_rna, e3, (double)cyc[0] / reps, (double)cyc[1] / reps / 64.0, (double)cyc[2] / reps / 64.0, (double)cyc[3] / reps / 64.0);
  cudaFree(dA); cudaFree(dB); cudaFree(dD1); cudaFree(dD3); cudaFree(dcyc);
}

int main() {
  run<16, 16>(1);
  run<32, 32>(1);
  run<64, 16>(1);
  run<128, 16>(1);
  run<16, 16>(148);
  return 0;
}
