// Micro-benchmark: what does the 4-tap bilinear gather of the warp cost on B200, per access path?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
// Coordinates follow the synthetic workload (random per-pixel disparity at scale s, upsampled; random
// small motion), lanes = columns, warps walk rows -- the access pattern of reproj_kernel.
//   A planar fp32, 12 LDG            B packed RGBA float4, 4 LDG.128
//   C tex2D<float4> point, pitch-linear   D tex2Dgather on a layered CUDA array (3 TLD4)
//   E tex2DLayered<float4> point on a CUDA array (block-linear)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int B = 12, H = 192, W = 640, HW = H * W;

__global__ void coords_kernel(const float* disp, int h, int w, const float* mot, float* cx, float* cy) {
  // disp (B,h,w) -> bilinear (align_corners=False) upsample -> depth -> parallax
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * HW) return;
  const int b = i / HW, y = (i % HW) / W, x = i % W;
  const float sy = (float)h / H, sx = (float)w / W;
  float fy = fmaxf((y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((x + 0.5f) * sx - 0.5f, 0.f);
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  float wy = fy - y0, wx = fx - x0;
  const float* d = disp + (size_t)b * h * w;
  float v = (1 - wy) * ((1 - wx) * d[y0 * w + x0] + wx * d[y0 * w + x1]) + wy * ((1 - wx) * d[y1 * w + x0] + wx * d[y1 * w + x1]);
  float depth = 1.0f / (0.01f + 9.99f * v);
  const float* m = mot + b * 6;  // tx, ty, tz, rx, ry, rz
  float X = (x - 320.f) / 371.f * depth, Y = (y - 96.f) / 368.f * depth, Z = depth;
  float Xr = X - m[5] * Y + m[4] * Z + m[0], Yr = m[5] * X + Y - m[3] * Z + m[1], Zr = -m[4] * X + m[3] * Y + Z + m[2];
  float u = 371.f * Xr / Zr + 320.f, vv = 368.f * Yr / Zr + 96.f;
  cx[i] = fminf(fmaxf(u, 0.f), W - 1.f);
  cy[i] = fminf(fmaxf(vv, 0.f), H - 1.f);
}

struct Tap { int x0, y0; float wx, wy; };
__device__ __forceinline__ Tap mk(float ix, float iy) {
  Tap t; t.x0 = (int)floorf(ix); t.y0 = (int)floorf(iy); t.wx = ix - t.x0; t.wy = iy - t.y0; return t;
}
__device__ __forceinline__ float bil(float nw, float ne, float sw, float se, const Tap& t) {
  return (1 - t.wy) * ((1 - t.wx) * nw + t.wx * ne) + t.wy * ((1 - t.wx) * sw + t.wx * se);
}

// grid: (ceil(W/32), H/ROWS, B); block 32 x NW (each warp walks ROWS/NW rows)
constexpr int ROWS = 48, NW = 4;
#define WALK for (int r = threadIdx.y; r < ROWS; r += NW)
#define PIX const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * ROWS + r, b = blockIdx.z; if (x >= W) continue; const int i = b * HW + y * W + x; const Tap t = mk(cx[i], cy[i]); const int x1 = min(t.x0 + 1, W - 1), y1 = min(t.y0 + 1, H - 1);

__global__ void kA(const float* img, const float* cx, const float* cy, float* out) {
  WALK { PIX
    const float* p = img + (size_t)b * 3 * HW;
    float acc = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* q = p + c * HW;
      acc += bil(q[t.y0 * W + t.x0], q[t.y0 * W + x1], q[y1 * W + t.x0], q[y1 * W + x1], t);
    }
    out[i] = acc;
  }
}
__global__ void kB(const float4* img, const float* cx, const float* cy, float* out) {
  WALK { PIX
    const float4* p = img + (size_t)b * HW;
    const float4 nw = p[t.y0 * W + t.x0], ne = p[t.y0 * W + x1], sw = p[y1 * W + t.x0], se = p[y1 * W + x1];
    out[i] = bil(nw.x, ne.x, sw.x, se.x, t) + bil(nw.y, ne.y, sw.y, se.y, t) + bil(nw.z, ne.z, sw.z, se.z, t);
  }
}
__global__ void kC(cudaTextureObject_t tex, const float* cx, const float* cy, float* out) {
  WALK { PIX
    const float fy0 = (float)(b * H + t.y0), fy1 = (float)(b * H + y1);
    const float4 nw = tex2D<float4>(tex, (float)t.x0, fy0), ne = tex2D<float4>(tex, (float)x1, fy0);
    const float4 sw = tex2D<float4>(tex, (float)t.x0, fy1), se = tex2D<float4>(tex, (float)x1, fy1);
    out[i] = bil(nw.x, ne.x, sw.x, se.x, t) + bil(nw.y, ne.y, sw.y, se.y, t) + bil(nw.z, ne.z, sw.z, se.z, t);
  }
}
__global__ void kD(cudaTextureObject_t tex, const float* cx, const float* cy, float* out) {
  WALK { PIX
    (void)x1; (void)y1;
    float acc = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // gather returns (x0,y1) (x1,y1) (x1,y0) (x0,y0) as .x .y .z .w for the footprint around (u,v)
      const float4 g = tex2Dgather<float4>(tex, t.x0 + 1.0f, (float)((b * 3 + c) * H + t.y0) + 1.0f, 0);
      acc += bil(g.w, g.z, g.x, g.y, t);
    }
    out[i] = acc;
  }
}
__global__ void kE(cudaTextureObject_t tex, const float* cx, const float* cy, float* out) {
  WALK { PIX
    const float4 nw = tex2DLayered<float4>(tex, (float)t.x0, (float)t.y0, b), ne = tex2DLayered<float4>(tex, (float)x1, (float)t.y0, b);
    const float4 sw = tex2DLayered<float4>(tex, (float)t.x0, (float)y1, b), se = tex2DLayered<float4>(tex, (float)x1, (float)y1, b);
    out[i] = bil(nw.x, ne.x, sw.x, se.x, t) + bil(nw.y, ne.y, sw.y, se.y, t) + bil(nw.z, ne.z, sw.z, se.z, t);
  }
}
// reference point: the same walk with no gather at all (coords in, one value out)
__global__ void kZ(const float* cx, const float* cy, float* out) {
  WALK { PIX
    (void)x1; (void)y1;
    out[i] = t.wx + t.wy;
  }
}

static float frand() { return rand() / (float)RAND_MAX; }

int main() {
  srand(7);
  std::vector<float> himg((size_t)B * 3 * HW);
  for (auto& v : himg) v = frand();
  std::vector<float4> hpk((size_t)B * HW);
  for (int b = 0; b < B; ++b)
    for (int p = 0; p < HW; ++p)
      hpk[(size_t)b * HW + p] = make_float4(himg[((size_t)b * 3 + 0) * HW + p], himg[((size_t)b * 3 + 1) * HW + p], himg[((size_t)b * 3 + 2) * HW + p], 0.f);
  float *img, *cx, *cy, *out, *ref;
  float4* pk;
  CK(cudaMalloc(&img, himg.size() * 4)); CK(cudaMemcpy(img, himg.data(), himg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&pk, hpk.size() * 16)); CK(cudaMemcpy(pk, hpk.data(), hpk.size() * 16, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cx, (size_t)B * HW * 4)); CK(cudaMalloc(&cy, (size_t)B * HW * 4));
  CK(cudaMalloc(&out, (size_t)B * HW * 4)); CK(cudaMalloc(&ref, (size_t)B * HW * 4));

  // C: pitch-linear float4 texture over the packed stack (B*H rows)
  cudaTextureObject_t texC, texD, texE;
  {
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = pk; rd.res.pitch2D.desc = cudaCreateChannelDesc<float4>();
    rd.res.pitch2D.width = W; rd.res.pitch2D.height = B * H; rd.res.pitch2D.pitchInBytes = W * 16;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&texC, &rd, &td, nullptr));
  }
  cudaArray_t arrD, arrE;
  {
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    CK(cudaMallocArray(&arrD, &cd, W, H * B * 3, cudaArrayTextureGather));
    CK(cudaMemcpy2DToArray(arrD, 0, 0, img, W * 4, W * 4, H * B * 3, cudaMemcpyDeviceToDevice));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arrD;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&texD, &rd, &td, nullptr));
  }
  {
    cudaChannelFormatDesc cd = cudaCreateChannelDesc<float4>();
    CK(cudaMalloc3DArray(&arrE, &cd, make_cudaExtent(W, H, B), cudaArrayLayered));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(pk, W * 16, W, H); p.dstArray = arrE; p.extent = make_cudaExtent(W, H, B); p.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arrE;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&texE, &rd, &td, nullptr));
  }

  std::vector<float> hmot(B * 6);
  float* mot; CK(cudaMalloc(&mot, B * 6 * 4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dim3 grid((W + 31) / 32, H / ROWS, B), block(32, NW);
  printf("px per launch: %d\n", B * HW);
  for (int s = 0; s < 4; ++s) {
    const int h = H >> s, w = W >> s;
    std::vector<float> hd((size_t)B * h * w);
    for (auto& v : hd) v = 0.01f + 0.29f * frand();
    float* disp; CK(cudaMalloc(&disp, hd.size() * 4)); CK(cudaMemcpy(disp, hd.data(), hd.size() * 4, cudaMemcpyHostToDevice));
    for (int b = 0; b < B; ++b)
      for (int k = 0; k < 6; ++k) {
        float n = 0; for (int j = 0; j < 12; ++j) n += frand(); n -= 6.f;  // ~N(0,1)
        hmot[b * 6 + k] = (k < 3 ? 0.02f : 0.01f) * n;
      }
    CK(cudaMemcpy(mot, hmot.data(), B * 6 * 4, cudaMemcpyHostToDevice));
    coords_kernel<<<(B * HW + 255) / 256, 256>>>(disp, h, w, mot, cx, cy);
    CK(cudaDeviceSynchronize());
    auto time = [&](const char* name, auto launch, bool check) {
      for (int i = 0; i < 3; ++i) launch();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      const int reps = 20;
      for (int i = 0; i < reps; ++i) launch();
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double err = 0;
      if (check) {
        std::vector<float> a((size_t)B * HW), r((size_t)B * HW);
        cudaMemcpy(a.data(), out, a.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(r.data(), ref, r.size() * 4, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < a.size(); ++i) err = fmax(err, fabs(a[i] - r[i]));
      }
      printf("scale %d  %-28s %8.2f us   %6.2f G px/s   maxerr %.2e\n", s, name, ms / reps * 1e3, B * HW / (ms / reps * 1e-3) * 1e-9, err);
    };
    time("Z no gather", [&] { kZ<<<grid, block>>>(cx, cy, ref); }, false);
    time("A planar 12xLDG", [&] { kA<<<grid, block>>>(img, cx, cy, ref); }, false);
    time("B packed 4xLDG.128", [&] { kB<<<grid, block>>>(pk, cx, cy, out); }, true);
    time("C tex2D float4 pitch2D", [&] { kC<<<grid, block>>>(texC, cx, cy, out); }, true);
    time("D tld4 tall array x3", [&] { kD<<<grid, block>>>(texD, cx, cy, out); }, true);
    time("E tex2DLayered float4 array", [&] { kE<<<grid, block>>>(texE, cx, cy, out); }, true);
    cudaFree(disp);
  }
  return 0;
}
