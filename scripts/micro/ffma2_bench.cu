// Microbenchmark: issue cost of packed fp32x2 arithmetic (__ffma2_rn / __fadd2_rn / __fmul2_rn) vs scalar
// fp32 on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_bench.cu -o ffma2_bench
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x0 = make_float2(threadIdx.x * 1e-3f, 1.0f), x1 = make_float2(2.0f, 3.0f), x2 = make_float2(0.5f, 0.25f), x3 = make_float2(4.f, 5.f);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {  // scalar: 8 independent FMA chains (same math as the packed version)
      x0.x = __fmaf_rn(x0.x, a, b); x0.y = __fmaf_rn(x0.y, a, b);
      x1.x = __fmaf_rn(x1.x, a, b); x1.y = __fmaf_rn(x1.y, a, b);
      x2.x = __fmaf_rn(x2.x, a, b); x2.y = __fmaf_rn(x2.y, a, b);
      x3.x = __fmaf_rn(x3.x, a, b); x3.y = __fmaf_rn(x3.y, a, b);
    } else if (MODE == 1) {  // packed: 4 FFMA2
      x0 = __ffma2_rn(x0, A, B); x1 = __ffma2_rn(x1, A, B); x2 = __ffma2_rn(x2, A, B); x3 = __ffma2_rn(x3, A, B);
    } else if (MODE == 2) {  // scalar mul+add (explicitly rounded, like the loss kernels)
      x0.x = __fadd_rn(__fmul_rn(x0.x, a), b); x0.y = __fadd_rn(__fmul_rn(x0.y, a), b);
      x1.x = __fadd_rn(__fmul_rn(x1.x, a), b); x1.y = __fadd_rn(__fmul_rn(x1.y, a), b);
      x2.x = __fadd_rn(__fmul_rn(x2.x, a), b); x2.y = __fadd_rn(__fmul_rn(x2.y, a), b);
      x3.x = __fadd_rn(__fmul_rn(x3.x, a), b); x3.y = __fadd_rn(__fmul_rn(x3.y, a), b);
    } else {  // packed mul+add
      x0 = __fadd2_rn(__fmul2_rn(x0, A), B); x1 = __fadd2_rn(__fmul2_rn(x1, A), B);
      x2 = __fadd2_rn(__fmul2_rn(x2, A), B); x3 = __fadd2_rn(__fmul2_rn(x3, A), B);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x0.y + x1.x + x1.y + x2.x + x2.y + x3.x + x3.y;
}

template <int MODE>
float run(float* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, 16, 0.999f, 0.001f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  const int iters = 20000;
  const double flop = 148.0 * 8 * 256 * iters * 8;  // 8 fma-equivalents per iteration per thread
  float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
  printf("scalar FFMA      %.3f ms  %.1f GFMA/s\n", t0, flop / t0 / 1e6);
  printf("packed FFMA2     %.3f ms  %.1f GFMA/s\n", t1, flop / t1 / 1e6);
  printf("scalar FMUL+FADD %.3f ms  %.1f G(mul+add)/s\n", t2, flop / t2 / 1e6);
  printf("packed FMUL2+FADD2 %.3f ms  %.1f G(mul+add)/s\n", t3, flop / t3 / 1e6);
  return 0;
}
