// Pipe-rate microbenchmarks for B200 (sm_100a): which instruction forms reach the
// FP32 peak, and what SHFL / MUFU / legacy-mma rates are. Informs the kernel design
// in DESIGN.md; not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
__constant__ float cw[64];

template <int NACC>
__global__ void k_ffma_rrr(float* out, const float* in) {
  float a[NACC]; float b = in[0], c = in[1];
#pragma unroll
  for (int i = 0; i < NACC; i++) a[i] = in[2 + i] + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = fmaf(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < NACC; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// matvec-like: acc[i] += w_k * h[k]; w from constant bank, h in registers
template <int NACC>
__global__ void k_ffma_const(float* out, const float* in) {
  float a[NACC]; float h[4];
#pragma unroll
  for (int i = 0; i < NACC; i++) a[i] = in[2 + i] + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = in[20 + i] * threadIdx.x;
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int i = 0; i < NACC; i++) a[i] = fmaf(cw[(k * NACC + i) & 63], h[k], a[i]);
  }
  float s = 0; for (int i = 0; i < NACC; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// matvec-like with weights from shared memory via broadcast LDS.128 (4 weights / load, each used 4x)
__global__ void k_ffma_lds(float* out, const float* in) {
  __shared__ float4 sw[64];
  if (threadIdx.x < 64) sw[threadIdx.x] = make_float4(in[threadIdx.x & 31], in[1], in[2], in[3]);
  __syncthreads();
  float a[16]; float h[4];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = in[2 + i] + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; i++) h[i] = in[20 + i] * threadIdx.x;
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float4 w = sw[(it * 4 + k) & 63];
      a[0 + k * 4 + 0 & 15] = fmaf(w.x, h[0], a[0 + k * 4 + 0 & 15]);
      a[1 + k * 4 + 0 & 15] = fmaf(w.x, h[1], a[1 + k * 4 + 0 & 15]);
      a[2 + k * 4 + 0 & 15] = fmaf(w.x, h[2], a[2 + k * 4 + 0 & 15]);
      a[3 + k * 4 + 0 & 15] = fmaf(w.x, h[3], a[3 + k * 4 + 0 & 15]);
      a[4 + k * 4 & 15] = fmaf(w.y, h[0], a[4 + k * 4 & 15]);
      a[5 + k * 4 & 15] = fmaf(w.y, h[1], a[5 + k * 4 & 15]);
      a[6 + k * 4 & 15] = fmaf(w.y, h[2], a[6 + k * 4 & 15]);
      a[7 + k * 4 & 15] = fmaf(w.y, h[3], a[7 + k * 4 & 15]);
      a[8 + k * 4 & 15] = fmaf(w.z, h[0], a[8 + k * 4 & 15]);
      a[9 + k * 4 & 15] = fmaf(w.z, h[1], a[9 + k * 4 & 15]);
      a[10 + k * 4 & 15] = fmaf(w.z, h[2], a[10 + k * 4 & 15]);
      a[11 + k * 4 & 15] = fmaf(w.z, h[3], a[11 + k * 4 & 15]);
      a[12 + k * 4 & 15] = fmaf(w.w, h[0], a[12 + k * 4 & 15]);
      a[13 + k * 4 & 15] = fmaf(w.w, h[1], a[13 + k * 4 & 15]);
      a[14 + k * 4 & 15] = fmaf(w.w, h[2], a[14 + k * 4 & 15]);
      a[15 + k * 4 & 15] = fmaf(w.w, h[3], a[15 + k * 4 & 15]);
    }
  }
  float s = 0; for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_ffma2(float* out, const float* in) {
  float2 a[NACC]; float2 b = make_float2(in[0], in[1]), c = make_float2(in[1], in[0]);
#pragma unroll
  for (int i = 0; i < NACC; i++) a[i] = make_float2(in[2 + i] + threadIdx.x, in[3 + i]);
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) a[i] = __ffma2_rn(a[i], b, c);
  }
  float s = 0; for (int i = 0; i < NACC; i++) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 with one operand pair loaded by broadcast LDS.128 (2 duplicated weights), 4 FFMA2 per load
__global__ void k_ffma2_lds(float* out, const float* in) {
  __shared__ float4 sw[64];
  if (threadIdx.x < 64) sw[threadIdx.x] = make_float4(in[threadIdx.x & 31], in[threadIdx.x & 31], in[2], in[2]);
  __syncthreads();
  float2 a[8]; float2 h[2];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = make_float2(in[2 + i] + threadIdx.x, in[3 + i]);
  h[0] = make_float2(in[20] * threadIdx.x, in[21]); h[1] = make_float2(in[22] * threadIdx.x, in[23]);
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float4 w = sw[(it * 4 + k) & 63];
      float2 w0 = make_float2(w.x, w.y), w1 = make_float2(w.z, w.w);
      a[(k * 2 + 0) & 7] = __ffma2_rn(w0, h[0], a[(k * 2 + 0) & 7]);
      a[(k * 2 + 1) & 7] = __ffma2_rn(w0, h[1], a[(k * 2 + 1) & 7]);
      a[(k * 2 + 4) & 7] = __ffma2_rn(w1, h[0], a[(k * 2 + 4) & 7]);
      a[(k * 2 + 5) & 7] = __ffma2_rn(w1, h[1], a[(k * 2 + 5) & 7]);
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mufu(float* out, const float* in) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = in[2 + i] + threadIdx.x * 1e-3f;
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a[i]));
      asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(e));
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// sigmoid-shaped mix: per 16 FFMA, 2 MUFU (ratio in the real kernel is ~40:1)
__global__ void k_mix(float* out, const float* in) {
  float a[16]; float b = in[0], c = in[1];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = in[2 + i] + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], b, c);
    float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a[it & 15]));
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(e));
    a[(it + 1) & 15] += e;
  }
  float s = 0; for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(float* out, const float* in) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = in[2 + i] + threadIdx.x;
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] += __shfl_xor_sync(0xffffffffu, a[i], 1 << (it & 3));
  }
  float s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy mma.sync m16n8k8 tf32, 8 independent accumulators
__global__ void k_mma_tf32(float* out, const float* in) {
  float c[8][4];
  unsigned a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = __float_as_uint(in[i] + threadIdx.x);
  b[0] = __float_as_uint(in[5]); b[1] = __float_as_uint(in[6]);
#pragma unroll
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = in[i + j];
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same with NC independent accumulator chains per warp: the dependent-accumulate latency of the legacy path
// (1 warp per sub-core: cycles per HMMA = max(pipe time, latency / NC))
template <int NC>
__global__ void k_mma_tf32_chains(float* out, const float* in) {
  float c[NC][4];
  unsigned a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = __float_as_uint(in[i] + threadIdx.x);
  b[0] = __float_as_uint(in[5]); b[1] = __float_as_uint(in[6]);
#pragma unroll
  for (int i = 0; i < NC; i++) for (int j = 0; j < 4; j++) c[i][j] = in[i + j];
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int r = 0; r < 8 / NC; r++)
#pragma unroll
      for (int i = 0; i < NC; i++)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0; for (int i = 0; i < NC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// legacy mma.sync m16n8k16 bf16 / f16 (fp32 accumulate), 8 independent accumulators: is the half-precision legacy path
// faster per MAC than the TF32 one?  (It decides whether the cross terms of a 3xTF32 weight-gradient contraction are worth
// moving to bf16.)
template <int KIND>  // 0: bf16, 1: f16
__global__ void k_mma_h16(float* out, const float* in) {
  float c[8][4];
  unsigned a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; i++) a[i] = 0x3c003c00u + threadIdx.x + i;
  b[0] = 0x3c003c00u; b[1] = 0x3c003c01u;
#pragma unroll
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = in[i + j];
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0; for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
static void run(const char* name, F launch, double ops_per_thread, int threads, int blocks, const char* unit) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; i++) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); for (int i = 0; i < 10; i++) launch(); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10; if (ms < best) best = ms;
  }
  double total = ops_per_thread * threads * (double)blocks;
  printf("{\"bench\":\"%s\",\"ms\":%.4f,\"rate\":%.4e,\"unit\":\"%s\",\"per_clk_per_sm_at_1965MHz\":%.2f}\n",
         name, best, total / (best * 1e-3), unit, total / (best * 1e-3) / 148.0 / 1.965e9);
}

int main() {
  float *in, *out; CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&out, 148 * 64 * 1024 * 4));
  float hin[64]; for (int i = 0; i < 64; i++) hin[i] = 0.5f + 0.001f * i; hin[0] = 0.999f; hin[1] = 1e-3f;
  CK(cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice));
  CK(cudaMemcpyToSymbol(cw, hin, sizeof(hin)));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, p.multiProcessorCount, p.clockRate);
  const int T = 256, B = 148 * 8;
  run("ffma_rrr_acc8", [&] { k_ffma_rrr<8><<<B, T>>>(out, in); }, ITERS * 8.0, T, B, "FMA/s");
  run("ffma_rrr_acc16", [&] { k_ffma_rrr<16><<<B, T>>>(out, in); }, ITERS * 16.0, T, B, "FMA/s");
  run("ffma_const_acc16", [&] { k_ffma_const<16><<<B, T>>>(out, in); }, ITERS * 16.0, T, B, "FMA/s");
  run("ffma_lds_bcast", [&] { k_ffma_lds<<<B, T>>>(out, in); }, ITERS * 16.0, T, B, "FMA/s");
  run("ffma2_acc8", [&] { k_ffma2<8><<<B, T>>>(out, in); }, ITERS * 16.0, T, B, "FMA/s");
  run("ffma2_acc16", [&] { k_ffma2<16><<<B, T>>>(out, in); }, ITERS * 32.0, T, B, "FMA/s");
  run("ffma2_lds_bcast", [&] { k_ffma2_lds<<<B, T>>>(out, in); }, ITERS * 8.0, T, B, "FMA/s");
  run("mufu_ex2_rcp", [&] { k_mufu<<<B, T>>>(out, in); }, ITERS / 4 * 16.0, T, B, "MUFU/s");
  run("mix_16ffma_2mufu", [&] { k_mix<<<B, T>>>(out, in); }, ITERS * 16.0, T, B, "FMA/s");
  run("shfl_xor", [&] { k_shfl<<<B, T>>>(out, in); }, ITERS / 4 * 8.0, T, B, "SHFL-lane/s");
  run("mma_sync_tf32_m16n8k8", [&] { k_mma_tf32<<<B, T>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, T, B, "MAC/s");
  run("mma_sync_bf16_m16n8k16", [&] { k_mma_h16<0><<<B, T>>>(out, in); }, ITERS / 4 * 8.0 * 2048.0 / 32.0, T, B, "MAC/s");
  run("mma_sync_f16_m16n8k16", [&] { k_mma_h16<1><<<B, T>>>(out, in); }, ITERS / 4 * 8.0 * 2048.0 / 32.0, T, B, "MAC/s");
  // 4 warps per SM (one per sub-core): what a single warp's 8 independent chains get
  run("mma_sync_tf32_m16n8k8_1warp_per_subcore", [&] { k_mma_tf32<<<148, 128>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, 128, 148, "MAC/s");
  run("mma_sync_bf16_m16n8k16_1warp_per_subcore", [&] { k_mma_h16<0><<<148, 128>>>(out, in); }, ITERS / 4 * 8.0 * 2048.0 / 32.0, 128, 148, "MAC/s");
  // dependent-accumulate latency: 1 warp per sub-core, 1 / 2 / 4 / 8 chains; and 3 warps per sub-core with 2 and 4 chains each
  run("mma_tf32_1warp_1chain", [&] { k_mma_tf32_chains<1><<<148, 128>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, 128, 148, "MAC/s");
  run("mma_tf32_1warp_2chains", [&] { k_mma_tf32_chains<2><<<148, 128>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, 128, 148, "MAC/s");
  run("mma_tf32_1warp_4chains", [&] { k_mma_tf32_chains<4><<<148, 128>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, 128, 148, "MAC/s");
  run("mma_tf32_1warp_8chains", [&] { k_mma_tf32_chains<8><<<148, 128>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, 128, 148, "MAC/s");
  run("mma_tf32_3warps_2chains", [&] { k_mma_tf32_chains<2><<<148, 384>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, 384, 148, "MAC/s");
  run("mma_tf32_3warps_4chains", [&] { k_mma_tf32_chains<4><<<148, 384>>>(out, in); }, ITERS / 4 * 8.0 * 1024.0 / 32.0, 384, 148, "MAC/s");
  // long FP32 run to read the sustained clock with nvidia-smi
  for (int i = 0; i < 400; i++) k_ffma2<16><<<B, T>>>(out, in);
  CK(cudaDeviceSynchronize());
  return 0;
}
