// l2_gather_peak.cu -- measures what the L2 -> SM path of this GPU can deliver for the access
// pattern of the propagation SpMM: random rows of an L2-resident fp32 table, 16 B per lane,
// LPR lanes per row, no index loads (row ids come from a hash), UNROLL independent gathers in
// flight per lane.  Also a plain streaming read of the same L2-resident table.
// This is the denominator of the "L2 roofline" quoted in DESIGN.md / bench.py; it is NOT part
// of the product.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_peak l2_gather_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

// SKEW: row r is drawn with probability ~ r^-1/2 (r = N u^2), the column popularity of the synthetic
// power-law graphs: the hottest rows are requested thousands of times per launch from all SMs
template <int V4, int UNROLL, bool SKEW = false>
__global__ void __launch_bounds__(256) gather_kernel(const float4* __restrict__ X, int n_rows, long long gathers_per_group,
                                                     float4* __restrict__ sink) {
  constexpr int LPR = V4;                       // lanes per row, one float4 per lane
  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t h = mix((uint32_t)group * 2654435761u + 12345u);
  for (long long it = 0; it < gathers_per_group; it += UNROLL) {
    float4 x[UNROLL];
#pragma unroll
    for (int q = 0; q < UNROLL; ++q) {
      h = h * 1664525u + 1013904223u;
      uint32_t r = __umulhi(mix(h), (uint32_t)n_rows);
      if (SKEW) {
        const float u = (float)(mix(h) >> 8) * (1.0f / 16777216.0f);
        r = min((uint32_t)((float)n_rows * u * u), (uint32_t)n_rows - 1);
      }
      x[q] = __ldg(X + (size_t)r * V4 + gl);
    }
#pragma unroll
    for (int q = 0; q < UNROLL; ++q) { acc.x += x[q].x; acc.y += x[q].y; acc.z += x[q].z; acc.w += x[q].w; }
  }
  if (acc.x == 123.456f) sink[0] = acc;         // never true; keeps the loads alive
}

template <int UNROLL>
__global__ void __launch_bounds__(256) stream_kernel(const float4* __restrict__ X, long long n4, int reps, float4* __restrict__ sink) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += stride * UNROLL) {
      float4 x[UNROLL];
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const long long kk = k + q * stride;
        x[q] = kk < n4 ? __ldcg(X + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) { acc.x += x[q].x; acc.y += x[q].y; acc.z += x[q].z; acc.w += x[q].w; }
    }
  }
  if (acc.x == 123.456f) sink[0] = acc;
}

template <typename F>
static float time_ms(F f, int reps = 20) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int k = 0; k < 3; ++k) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int k = 0; k < reps; ++k) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

template <int V4, int UNROLL, bool SKEW = false>
static void run_gather(const float4* X, int n_rows, float4* sink, int ctas_per_sm) {
  const long long total_gathers = 1ll << 24;                   // 16 M row gathers per launch
  const int blocks = 148 * ctas_per_sm;
  const long long groups = (long long)blocks * 256 / V4;
  long long per = (total_gathers / groups + UNROLL - 1) / UNROLL * UNROLL;
  const float ms = time_ms([&] { gather_kernel<V4, UNROLL, SKEW><<<blocks, 256>>>(X, n_rows, per, sink); });
  CK(cudaGetLastError());
  const double bytes = (double)per * groups * V4 * 16;
  printf("gather %s row=%4d B  unroll=%2d  ctas/sm=%d  table=%.1f MB : %8.1f GB/s  (%.3f ms)\n", SKEW ? "power-law" : "uniform  ", V4 * 16, UNROLL, ctas_per_sm,
         (double)n_rows * V4 * 16 / 1e6, bytes / ms / 1e6, ms);
}

int main(int argc, char** argv) {
  const int n_rows = argc > 1 ? atoi(argv[1]) : 70839;         // Gowalla shape: N = 29858 + 40981
  float4 *X, *sink;
  const size_t bytes = (size_t)n_rows * 256;
  CK(cudaMalloc(&X, bytes));
  CK(cudaMalloc(&sink, 64));
  CK(cudaMemset(X, 0, bytes));
  int clk = 0;
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("table rows %d, SM clock attr %d kHz\n", n_rows, clk);
  for (int c : {4, 8}) {
    run_gather<16, 4>(X, n_rows, sink, c);
    run_gather<16, 8>(X, n_rows, sink, c);
    run_gather<16, 16>(X, n_rows, sink, c);
  }
  run_gather<16, 8, true>(X, n_rows, sink, 4);
  run_gather<16, 8, true>(X, n_rows, sink, 8);
  run_gather<8, 8, true>(X, n_rows * 2, sink, 8);
  run_gather<8, 16>(X, n_rows * 2, sink, 8);
  run_gather<4, 16>(X, n_rows * 4, sink, 8);
  run_gather<2, 16>(X, n_rows * 8, sink, 8);
  run_gather<32, 8>(X, n_rows / 2, sink, 8);
  {
    const long long n4 = (long long)bytes / 16;
    for (int c : {4, 8}) {
      const int blocks = 148 * c, reps = 8;
      const float ms = time_ms([&] { stream_kernel<8><<<blocks, 256>>>(X, n4, reps, sink); });
      printf("stream (L2-resident %.1f MB, ld.cg) ctas/sm=%d : %8.1f GB/s\n", bytes / 1e6, c, (double)bytes * reps / ms / 1e6);
    }
  }
  return 0;
}
