// Microbenchmark (measurement script, not product code): cycles per 32 KB weight chunk of the GEMV core of the rows
// schedule when the chunk is already in shared memory - the compute bound that a column-split CTA pair would run into.
//   variant 0: 4 rows, 32 k x 256 columns, 256 threads  (the shipped mapping: cg = tid & 63, ks = tid >> 6)
//   variant 1: 8 rows, 64 k x 128 columns, 256 threads  (cg = tid & 31, 8 k-slices of 8)
//   variant 2: 8 rows, 64 k x 128 columns, 512 threads  (16 k-slices of 4)
//   variant 3: raw FFMA2 issue rate, 16 independent accumulators per thread
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int NR>
__device__ __forceinline__ void fma_row(float2 (&acc)[NR / 2][4], const float* __restrict__ xs, float4 w) {
  const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y), w2 = make_float2(w.z, w.z), w3 = make_float2(w.w, w.w);
#pragma unroll
  for (int q = 0; q < NR / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(xs + 4 * q);
    const float2 xa = make_float2(v.x, v.y), xb = make_float2(v.z, v.w);
    acc[2 * q][0] = __ffma2_rn(xa, w0, acc[2 * q][0]);
    acc[2 * q][1] = __ffma2_rn(xa, w1, acc[2 * q][1]);
    acc[2 * q][2] = __ffma2_rn(xa, w2, acc[2 * q][2]);
    acc[2 * q][3] = __ffma2_rn(xa, w3, acc[2 * q][3]);
    acc[2 * q + 1][0] = __ffma2_rn(xb, w0, acc[2 * q + 1][0]);
    acc[2 * q + 1][1] = __ffma2_rn(xb, w1, acc[2 * q + 1][1]);
    acc[2 * q + 1][2] = __ffma2_rn(xb, w2, acc[2 * q + 1][2]);
    acc[2 * q + 1][3] = __ffma2_rn(xb, w3, acc[2 * q + 1][3]);
  }
}

// NR rows, chunk = CK k rows x NC columns (CK * NC = 8192 floats), KT k rows per thread and chunk
template <int NR, int CK, int NC, int KT>
__global__ void gemv_kernel(float* out, long long* cyc, int nchunks) {
  extern __shared__ __align__(16) float sm[];
  float* ring = sm;                       // 4 slots x 8192 floats
  float* xT = sm + 4 * 8192;              // [256][NR]
  for (int i = threadIdx.x; i < 4 * 8192; i += blockDim.x) ring[i] = 1e-3f * (float)((i * 7 + blockIdx.x) % 13);
  for (int i = threadIdx.x; i < 256 * NR; i += blockDim.x) xT[i] = 1e-2f * (float)(i % 11);
  __syncthreads();
  constexpr int CGS = NC / 4;             // threads across the columns
  const int cg = threadIdx.x % CGS, ks = threadIdx.x / CGS;
  float2 acc[NR / 2][4];
#pragma unroll
  for (int p = 0; p < NR / 2; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = make_float2(0.f, 0.f);
  const long long t0 = clock64();
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    const int slot = c & 3, k0 = (c * CK) & 255;
    const float* ws = ring + slot * 8192 + 4 * cg;
    const float* xs = xT + ((k0 + ks * KT) & 255) * NR;
    float4 w[KT];
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) w[kk] = *reinterpret_cast<const float4*>(ws + (ks * KT + kk) * NC);
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) fma_row<NR>(acc, xs + kk * NR, w[kk]);
    __syncwarp();
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < NR / 2; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) s += acc[p][c].x + acc[p][c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void ffma2_kernel(float* out, long long* cyc, int iters) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2((float)threadIdx.x, (float)i);
  const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(1e-4f, -1e-4f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const size_t smem = (4 * 8192 + 256 * 8) * 4;
  cudaFuncSetAttribute(gemv_kernel<4, 32, 256, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(gemv_kernel<8, 64, 128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(gemv_kernel<8, 64, 128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(gemv_kernel<8, 32, 256, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int n = 400;
  long long h[148];
  auto report = [&](const char* name, double per) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60;
    for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    printf("%-58s %8.1f .. %8.1f cycles per unit  (%s)\n", name, mn / per, mx / per, cudaGetErrorString(e));
  };
  for (int rep = 0; rep < 2; ++rep) {
    gemv_kernel<4, 32, 256, 8><<<148, 256, smem>>>(out, cyc, n); report("v0 4 rows, 32k x 256 cols, 256 thr (shipped), per chunk", n);
    gemv_kernel<8, 32, 256, 8><<<148, 256, smem>>>(out, cyc, n); report("v0b 8 rows, 32k x 256 cols, 256 thr, per chunk", n);
    gemv_kernel<8, 64, 128, 8><<<148, 256, smem>>>(out, cyc, n); report("v1 8 rows, 64k x 128 cols, 256 thr, per chunk", n);
    gemv_kernel<8, 64, 128, 4><<<148, 512, smem>>>(out, cyc, n); report("v2 8 rows, 64k x 128 cols, 512 thr, per chunk", n);
    ffma2_kernel<<<148, 128, 0>>>(out, cyc, 1000); report("v3 FFMA2 x16 per iteration, 1 warp / SMSP, per iteration", 1000);
    ffma2_kernel<<<148, 256, 0>>>(out, cyc, 1000); report("v3 FFMA2 x16 per iteration, 2 warps / SMSP, per iteration", 1000);
    ffma2_kernel<<<148, 512, 0>>>(out, cyc, 1000); report("v3 FFMA2 x16 per iteration, 4 warps / SMSP, per iteration", 1000);
  }
  return 0;
}
