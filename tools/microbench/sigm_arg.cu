// Error of the MUFU-based sigmoid for two ways of forming the MUFU.EX2 argument (v11 of the step kernel):
//   old: ex2(fl(fl(v + b) * -log2e))      new: ex2(fma(v, -log2e, fl(b * -log2e)))
// against float64, binned by the pre-activation u = v + b.   nvcc -O3 -arch=sm_100a -o sigm_arg sigm_arg.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ float ex2a(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float rcpa(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__global__ void k(const float* v, const float* b, float* o_old, float* o_new, float* e_old, float* e_new, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float c = -1.4426950408889634f;
  const float u = __fadd_rn(v[i], b[i]);
  const float eo = ex2a(__fmul_rn(u, c));
  const float en = ex2a(__fmaf_rn(v[i], c, __fmul_rn(b[i], c)));
  e_old[i] = eo; e_new[i] = en;
  o_old[i] = rcpa(1.0f + eo);
  o_new[i] = rcpa(1.0f + en);
}
int main() {
  const int n = 1 << 22;
  std::vector<float> v(n), b(n), oo(n), on(n), eo(n), en(n);
  srand(1);
  for (int i = 0; i < n; i++) {
    v[i] = ((float)rand() / RAND_MAX - 0.5f) * 40.0f;
    b[i] = ((float)rand() / RAND_MAX - 0.5f) * 4.0f;
  }
  float *dv, *db, *doo, *don, *deo, *den;
  cudaMalloc(&dv, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&doo, n * 4); cudaMalloc(&don, n * 4); cudaMalloc(&deo, n * 4); cudaMalloc(&den, n * 4);
  cudaMemcpy(dv, v.data(), n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 4, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(dv, db, doo, don, deo, den, n);
  cudaMemcpy(oo.data(), doo, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(on.data(), don, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(eo.data(), deo, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(en.data(), den, n * 4, cudaMemcpyDeviceToHost);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
  // bins of u: [-22,-14) [-14,-6) [-6,-2) [-2,2) [2,6) [6,14) [14,22)
  const double edges[8] = {-22, -14, -6, -2, 2, 6, 14, 22};
  printf("bin(u)        n     | exp: old mean/rms rel err, new mean/rms | sigmoid: old mean/rms rel err, new mean/rms | 1-s abs: old rms, new rms\n");
  for (int bin = 0; bin < 7; bin++) {
    double s[8] = {0}; long cnt = 0; double q[2] = {0, 0};
    for (int i = 0; i < n; i++) {
      const double u = (double)v[i] + (double)b[i];
      if (u < edges[bin] || u >= edges[bin + 1]) continue;
      const double ex = exp(-u), sg = 1.0 / (1.0 + ex);
      const double r[4] = {(eo[i] - ex) / ex, (en[i] - ex) / ex, (oo[i] - sg) / sg, (on[i] - sg) / sg};
      for (int j = 0; j < 4; j++) { s[2 * j] += r[j]; s[2 * j + 1] += r[j] * r[j]; }
      q[0] += (oo[i] - sg) * (oo[i] - sg); q[1] += (on[i] - sg) * (on[i] - sg);
      cnt++;
    }
    if (!cnt) continue;
    printf("[%4.0f,%4.0f) %8ld |", edges[bin], edges[bin + 1], cnt);
    for (int j = 0; j < 4; j++) printf(" %+.2e/%.2e%s", s[2 * j] / cnt, sqrt(s[2 * j + 1] / cnt), j == 1 ? " |" : "");
    printf(" | %.2e %.2e\n", sqrt(q[0] / cnt), sqrt(q[1] / cnt));
  }
  return 0;
}
