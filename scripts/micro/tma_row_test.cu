// stand-alone check of the TMA row staging used by reproj_stream_tma_kernel (bbd_stream.cuh helpers)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define BBD_HD __host__ __device__ __forceinline__
#include "../../baseboostdepth_b200/csrc/bbd_stream.cuh"
using namespace bbd;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k2(const CUtensorMap* maps, int mode, int x, int y, int z, float* out) {
  extern __shared__ __align__(128) float smem[];
  float* bar = smem + 640;
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    tma_bar_init(bar, 1);
    const int bytes = mode == 3 ? 384 : 128;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
    if (mode == 3) tma_box3(&maps[0], smem, bar, x, y, z);
    else tma_box3(&maps[1], smem, bar, x, y, z);
  }
  __syncwarp();
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 16) && !done; ++spin)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(addr), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < 160; i += 32) out[i] = smem[i];
  if (threadIdx.x == 0) out[160] = (float)done;
}

__global__ void k(const __grid_constant__ CUtensorMap tt, const __grid_constant__ CUtensorMap td, const __grid_constant__ CUtensorMap ti,
                  bbd_reproj_args a, int x, int y, int s, int b, float* out) {
  extern __shared__ __align__(128) float smem[];
  float* bar = smem + 640;
  StreamTmaMaps m = {&tt, &td, &ti};
  if (threadIdx.x == 0) {
    tma_bar_init(bar, 4);
    tma_row_issue(m, a, smem, bar, x, y, s, b);
  }
  __syncwarp();
  unsigned done = 0;
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  for (int spin = 0; spin < (1 << 16) && !done; ++spin)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(addr), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < 160; i += 32) out[i] = smem[i];
  if (threadIdx.x == 0) { out[160] = (float)done; unsigned long long st = *reinterpret_cast<volatile unsigned long long*>(bar); out[161] = (float)(st >> 32); out[162] = (float)(st & 0xffffffffu); }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 96, H = 48, B = 3, S = 2;
  std::vector<float> ht((size_t)B * 3 * H * W), hd((size_t)S * B * H * W), hi((size_t)B * H * W);
  for (size_t i = 0; i < ht.size(); ++i) ht[i] = (float)i;
  for (size_t i = 0; i < hd.size(); ++i) hd[i] = 1e6f + i;
  for (size_t i = 0; i < hi.size(); ++i) hi[i] = 2e6f + i;
  float *t, *d, *im, *out;
  CK(cudaMalloc(&t, ht.size() * 4)); CK(cudaMalloc(&d, hd.size() * 4)); CK(cudaMalloc(&im, hi.size() * 4)); CK(cudaMalloc(&out, 164 * 4)); CK(cudaMemset(out, 0, 164*4));
  CK(cudaMemcpy(t, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d, hd.data(), hd.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(im, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice));
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)p;
  auto mk = [&](CUtensorMap* m, float* base, long planes, int bp) {
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {32, 1, (cuuint32_t)bp};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
  };
  CUtensorMap tt, td, ti;
  mk(&tt, t, 3L * B, 3); mk(&td, d, (long)S * B, 1); mk(&ti, im, B, 1);
  bbd_reproj_args a = {};
  a.batch = B; a.height = H; a.width = W; a.num_scales = S;
  const int x = -2, y = 5, s = 1, b = 2;
  if (mode == 0) k<<<1, 32, 4096>>>(tt, td, ti, a, x, y, s, b, out);
  else {
    CUtensorMap hm[2] = {tt, td};
    CUtensorMap* dm; CK(cudaMalloc(&dm, sizeof(hm))); CK(cudaMemcpy(dm, hm, sizeof(hm), cudaMemcpyHostToDevice));
    k2<<<1, 32, 4096>>>(dm, mode, mode == 4 ? 0 : (mode == 5 ? -4 : (mode == 6 ? 2 : (mode == 7 ? 68 : x))), y, mode == 3 ? 3 * b : s * B + b, out);
  }
  CK(cudaDeviceSynchronize());
  std::vector<float> ho(164);
  CK(cudaMemcpy(ho.data(), out, 164*4, cudaMemcpyDeviceToHost));
  printf("done=%g state hi=%g lo=%g\n", ho[160], ho[161], ho[162]);
  int bad = 0;
  for (int i = 0; i < 32; ++i) {
    const int u = x + i; const bool in = u >= 0 && u < W;
    for (int c = 0; c < 3; ++c) { float want = in ? ht[((size_t)(b * 3 + c) * H + y) * W + u] : 0; if (ho[c * 32 + i] != want) ++bad; }
    float wd = in ? hd[((size_t)(s * B + b) * H + y) * W + u] : 0; if (ho[96 + i] != wd) ++bad;
    float wi = in ? hi[((size_t)b * H + y) * W + u] : 0; if (ho[128 + i] != wi) ++bad;
  }
  printf("mismatches: %d  (first values %g %g %g | %g | %g)\n", bad, ho[0], ho[2], ho[3], ho[98], ho[130]);
  return bad != 0;
}
