// Dependent-issue latency of scalar and packed fp32 arithmetic, shuffles and shared-memory loads on one warp:
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench lat_bench.cu && ./lat_bench
#include <cuda_runtime.h>
#include <cstdio>
__global__ void k(float* out, long long* cyc, float a, float b) {
  __shared__ float sm[64];
  sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
  __syncwarp();
  const int N = 4096;
  float x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __fmaf_rn(x, b, a);
  long long t1 = clock64();
  float2 y = make_float2(a + threadIdx.x, b);
  const float2 bb = make_float2(b, b), aa = make_float2(a, a);
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = __ffma2_rn(y, bb, aa);
  long long t2 = clock64();
  float z = a;
#pragma unroll 16
  for (int i = 0; i < N; ++i) z = __shfl_up_sync(0xffffffffu, z, 1) + 1.0f;
  long long t3 = clock64();
  int idx = threadIdx.x;
#pragma unroll 16
  for (int i = 0; i < N; ++i) idx = ((int)sm[idx & 63]) & 31;
  long long t4 = clock64();
  float w = a + 2.0f;
#pragma unroll 16
  for (int i = 0; i < N; ++i) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w)); w = r + 1.5f; }
  long long t5 = clock64();
  float2 q = make_float2(a, b);
#pragma unroll 16
  for (int i = 0; i < N; ++i) q = __fadd2_rn(q, aa);
  long long t6 = clock64();
  out[threadIdx.x] = x + y.x + y.y + z + idx + w + q.x + q.y;
  if (threadIdx.x == 0) { cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = (t4 - t3); cyc[4] = (t5 - t4); cyc[5] = t6 - t5; cyc[6] = N; }
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 128); cudaMalloc(&c, 64);
  k<<<1, 32>>>(o, c, 0.5f, 0.999f); k<<<1, 32>>>(o, c, 0.5f, 0.999f);
  long long h[7]; cudaMemcpy(h, c, 56, cudaMemcpyDeviceToHost);
  printf("cycles per dependent op: FFMA %.2f  FFMA2 %.2f  SHFL+FADD %.2f  LDS+cvt %.2f  MUFU.RCP+FADD %.2f  FADD2 %.2f\n", (double)h[0] / h[6], (double)h[1] / h[6],
         (double)h[2] / h[6], (double)h[3] / h[6], (double)h[4] / h[6], (double)h[5] / h[6]);
  return 0;
}
