// gather_ingredients.cu -- which ingredient of the SpMM inner loop costs L2->SM gather throughput?
// Starts from the pure gather of l2_gather_peak.cu (16 lanes x 16 B per 256-byte row, hashed row ids) and adds, one
// at a time: the two SHFL broadcasts per gather, a streamed (col, val) index array instead of hashed ids, the
// predicated form, FMA accumulation, and a per-row epilogue (load + store of an output row every 32 gathers).
// Tuning aid only.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_ingredients gather_ingredients.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16; return h; }

// MODE bit 0: shuffles, bit 1: streamed indices, bit 2: predicated, bit 3: epilogue every 32 gathers
template <int MODE>
__global__ void __launch_bounds__(256, 4) k(const float4* __restrict__ X, int n_rows, const int2* __restrict__ cv, long long per_group,
                                            float4* __restrict__ Y, float4* __restrict__ sink) {
  const int lane = threadIdx.x & 31, gl = lane & 15;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 16;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t h = mix((uint32_t)group * 2654435761u + 12345u);
  const int2* mycv = cv + group * per_group;
  for (long long it = 0; it < per_group; it += 16) {
    int c; float v;
    if (MODE & 2) { const int2 e = __ldg(mycv + it + gl); c = e.x; v = __int_as_float(e.y); }
    else { h = h * 1664525u + 1013904223u; const float u = (float)(mix(h) >> 8) * (1.0f / 16777216.0f); c = min((int)((float)n_rows * u * u), n_rows - 1); v = 1.0f; }
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      int ct; float vt;
      if (MODE & 1) { ct = __shfl_sync(0xffffffffu, c, t, 16); vt = __shfl_sync(0xffffffffu, v, t, 16); }
      else { ct = (int)(((uint32_t)c + (uint32_t)t * 2654435761u) % (uint32_t)n_rows); vt = v; }
      if (!(MODE & 4) || ct >= 0) {
        const float4 x = __ldg(X + (size_t)ct * 16 + gl);
        acc.x = fmaf(vt, x.x, acc.x); acc.y = fmaf(vt, x.y, acc.y); acc.z = fmaf(vt, x.z, acc.z); acc.w = fmaf(vt, x.w, acc.w);
      }
    }
    if ((MODE & 8) && ((it >> 4) & 1)) {
      const size_t r = (size_t)((group * 977 + (it >> 5)) % n_rows) * 16 + gl;
      const float4 a = Y[r];
      Y[r] = make_float4(a.x + acc.x, a.y + acc.y, a.z + acc.z, a.w + acc.w);
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (acc.x == 123.456f) sink[0] = acc;
}

template <int MODE>
static void run(const char* what, const float4* X, int n_rows, const int2* cv, long long per, float4* Y, float4* sink, int blocks) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) k<MODE><<<blocks, 256>>>(X, n_rows, cv, per, Y, sink);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < 20; ++i) k<MODE><<<blocks, 256>>>(X, n_rows, cv, per, Y, sink);
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 20;
  const double bytes = (double)per * blocks * 16 * 256;
  printf("%-62s %8.1f GB/s  (%.3f ms)\n", what, bytes / ms / 1e6, ms);
}

int main() {
  const int n_rows = 70839, blocks = 148 * 4;
  const long long groups = (long long)blocks * 16, per = 1792;        // 17 M gathers per launch
  float4 *X, *Y, *sink; int2* cv;
  CK(cudaMalloc(&X, (size_t)n_rows * 256)); CK(cudaMalloc(&Y, (size_t)n_rows * 256)); CK(cudaMalloc(&sink, 64));
  CK(cudaMemset(X, 0, (size_t)n_rows * 256)); CK(cudaMemset(Y, 0, (size_t)n_rows * 256));
  std::vector<int2> h((size_t)groups * per);
  uint32_t s = 12345u;
  for (auto& e : h) { s = s * 1664525u + 1013904223u; const float u = (float)(s >> 8) * (1.0f / 16777216.0f); e.x = (int)((float)n_rows * u * u) % n_rows; e.y = 0x3f800000; }
  CK(cudaMalloc(&cv, h.size() * 8)); CK(cudaMemcpy(cv, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  run<0>("pure gather (arithmetic row ids)", X, n_rows, cv, per, Y, sink, blocks);
  run<1>("+ 2 SHFL per gather", X, n_rows, cv, per, Y, sink, blocks);
  run<2>("+ streamed (col,val) indices, no SHFL", X, n_rows, cv, per, Y, sink, blocks);
  run<3>("+ streamed indices + SHFL", X, n_rows, cv, per, Y, sink, blocks);
  run<7>("+ streamed indices + SHFL + predicated", X, n_rows, cv, per, Y, sink, blocks);
  run<11>("+ streamed indices + SHFL + epilogue every 32", X, n_rows, cv, per, Y, sink, blocks);
  run<15>("+ streamed indices + SHFL + predicated + epilogue every 32", X, n_rows, cv, per, Y, sink, blocks);
  return 0;
}
