// tma_gather_peak.cu -- can the Blackwell TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4: four rows of a 2-D
// tensor per instruction, straight into shared memory) feed the propagation SpMM faster than LDG.128 gathers do?
// Same access pattern as tools/l2_gather_peak.cu (random rows of an L2-resident fp32 table, hashed row ids, uniform or with
// the graph's power-law column popularity), but every warp runs a ring of STAGES x 32 rows in shared memory: lanes 0..7
// issue one gather4 each for the stage that was just consumed, all lanes read the landed rows back with LDS.128 (16 lanes
// per row, two rows per instruction) and accumulate them -- the inner loop a TMA-fed SpMM would have, without index loads.
// A checksum is compared with an LDG kernel that walks the same ids.  Tuning aid, NOT part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_peak tma_gather_peak.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// row id of entry e (0..31) of iteration it of global warp gw
__device__ __forceinline__ uint32_t row_of(uint32_t gw, uint32_t it, uint32_t e, int n_rows, bool skew) {
  const uint32_t h = mix(gw * 2654435761u + it * 40503u + e * 97u + 12345u);
  if (skew) {
    const float u = (float)(mix(h + 1u) >> 8) * (1.0f / 16777216.0f);
    return min((uint32_t)((float)n_rows * u * u), (uint32_t)n_rows - 1);
  }
  return __umulhi(h, (uint32_t)n_rows);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spins = 0;; ++spins) {
    uint32_t done;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (spins > (1u << 22)) __trap();           // a protocol bug traps instead of hanging the box
  }
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
               : "memory");
}

// D floats per row; ROWS rows per stage (multiple of 4, <= 32)
template <int D, int STAGES, int WARPS, int ROWS>
__global__ void __launch_bounds__(WARPS * 32) tma_gather_kernel(const __grid_constant__ CUtensorMap tmap, int n_rows, int iters,
                                                                 int skew, unsigned long long* __restrict__ sums) {
  constexpr int V4 = D / 4;
  constexpr int LPR = V4 < 32 ? V4 : 32;
  constexpr int RPI = 32 / LPR;                          // rows per LDS instruction
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float4* ring = reinterpret_cast<float4*>(smem_raw);    // [WARPS][STAGES][ROWS][V4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * STAGES * ROWS * D * 4);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * WARPS + warp;
  float4* my = ring + (size_t)warp * STAGES * ROWS * V4;
  uint64_t* mybar = bars + warp * STAGES;
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(mybar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int it) {
    const int s = it % STAGES;
    if (lane == 0) mbar_arrive_expect_tx(mybar + s, ROWS * D * 4);
    __syncwarp();
    if (lane < ROWS / 4) {
      const int e = lane * 4;
      tma_gather4(my + ((size_t)s * ROWS + e) * V4, &tmap, mybar + s, 0,
                  (int)row_of(gw, it, e, n_rows, skew), (int)row_of(gw, it, e + 1, n_rows, skew),
                  (int)row_of(gw, it, e + 2, n_rows, skew), (int)row_of(gw, it, e + 3, n_rows, skew));
    }
  };
  for (int it = 0; it < STAGES - 1 && it < iters; ++it) issue(it);
  const int gl = lane % LPR, grp = lane / LPR;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
    if (it + STAGES - 1 < iters) issue(it + STAGES - 1);     // the stage consumed in the previous iteration
    const int s = it % STAGES;
    mbar_wait(mybar + s, (it / STAGES) & 1);
    const float4* st = my + (size_t)s * ROWS * V4;
#pragma unroll
    for (int e = 0; e < ROWS; e += RPI) {
      const float4 x = st[(e + grp) * V4 + gl];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    __syncwarp();                                             // every lane is done with stage s before it is refilled
  }
  const unsigned long long t = (unsigned long long)(acc.x + acc.y + acc.z + acc.w);
  atomicAdd(sums + (gw & 1023), t);
}

template <int D, int UNROLL, int ROWS>
__global__ void __launch_bounds__(256) ldg_gather_kernel(const float4* __restrict__ X, int n_rows, int iters, int skew,
                                                         unsigned long long* __restrict__ sums) {
  constexpr int V4 = D / 4;
  constexpr int LPR = V4 < 32 ? V4 : 32;
  constexpr int RPI = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  const int gl = lane % LPR, grp = lane / LPR;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int e0 = 0; e0 < ROWS; e0 += RPI * UNROLL) {
      float4 x[UNROLL];
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) {
        const int e = e0 + q * RPI + grp;
        x[q] = e < ROWS ? __ldg(X + (size_t)row_of(gw, it, e, n_rows, skew) * V4 + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q = 0; q < UNROLL; ++q) { acc.x += x[q].x; acc.y += x[q].y; acc.z += x[q].z; acc.w += x[q].w; }
    }
  }
  const unsigned long long t = (unsigned long long)(acc.x + acc.y + acc.z + acc.w);
  atomicAdd(sums + (gw & 1023), t);
}

__global__ void fill_kernel(float* X, long long n, int d) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
    X[k] = (float)(((k / d) * 5 + (k % d)) & 7);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static unsigned long long checksum(unsigned long long* d_sums) {
  unsigned long long h[1024], t = 0;
  CK(cudaMemcpy(h, d_sums, sizeof(h), cudaMemcpyDeviceToHost));
  for (int k = 0; k < 1024; ++k) t += h[k];
  return t;
}

template <int D, int STAGES, int WARPS, int ROWS>
static void run(const float* X, int n_rows, EncodeTiledFn enc, int ctas_per_sm, int skew, unsigned long long* d_sums) {
  CUtensorMap tmap;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)n_rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)D, 1u};                    // gather4: a box is ONE row; the instruction names four
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return; }
  const size_t smem = (size_t)WARPS * STAGES * ROWS * D * 4 + WARPS * STAGES * 8;
  auto kern = tma_gather_kernel<D, STAGES, WARPS, ROWS>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = 148 * ctas_per_sm;
  const long long total_rows = 1ll << 24;
  const int iters = (int)(total_rows / ((long long)blocks * WARPS * ROWS));
  CK(cudaMemset(d_sums, 0, 1024 * 8));
  kern<<<blocks, WARPS * 32, smem>>>(tmap, n_rows, iters, skew, d_sums);
  CK(cudaDeviceSynchronize());
  const unsigned long long got = checksum(d_sums);
  CK(cudaMemset(d_sums, 0, 1024 * 8));
  ldg_gather_kernel<D, 8, ROWS><<<blocks * WARPS / 8, 256>>>((const float4*)X, n_rows, iters, skew, d_sums);
  CK(cudaDeviceSynchronize());
  const unsigned long long want = checksum(d_sums);
  const float ms = time_ms([&] { kern<<<blocks, WARPS * 32, smem>>>(tmap, n_rows, iters, skew, d_sums); });
  const float ms_ldg = time_ms([&] { ldg_gather_kernel<D, 8, ROWS><<<blocks * WARPS / 8, 256>>>((const float4*)X, n_rows, iters, skew, d_sums); });
  CK(cudaGetLastError());
  const double bytes = (double)iters * blocks * WARPS * ROWS * D * 4;
  printf("%s row=%4d B stages=%d warps=%d rows/stage=%2d ctas/sm=%d smem=%3zu KB : TMA gather4 %8.1f GB/s | LDG.128 (same ids, %d warps/SM) %8.1f GB/s | checksum %s (%llu vs %llu)\n",
         skew ? "power-law" : "uniform  ", D * 4, STAGES, WARPS, ROWS, ctas_per_sm, smem >> 10, bytes / ms / 1e6, ctas_per_sm * WARPS,
         bytes / ms_ldg / 1e6, got == want ? "OK" : "MISMATCH", got, want);
}

int main(int argc, char** argv) {
  const int n_rows = argc > 1 ? atoi(argv[1]) : 70839;         // Gowalla shape: N = 29858 + 40981
  float* X;
  unsigned long long* d_sums;
  const size_t bytes = (size_t)n_rows * 256;
  CK(cudaMalloc(&X, bytes));
  CK(cudaMalloc(&d_sums, 1024 * 8));
  fill_kernel<<<1024, 256>>>(X, (long long)bytes / 4, 64);
  CK(cudaDeviceSynchronize());
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  if (enc == nullptr) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  for (int skew : {0, 1}) {
    run<64, 2, 8, 32>(X, n_rows, enc, 1, skew, d_sums);
    run<64, 3, 8, 32>(X, n_rows, enc, 1, skew, d_sums);
    run<64, 3, 4, 32>(X, n_rows, enc, 2, skew, d_sums);
    run<64, 4, 4, 16>(X, n_rows, enc, 3, skew, d_sums);
    run<64, 3, 8, 16>(X, n_rows, enc, 2, skew, d_sums);
    run<64, 6, 8, 16>(X, n_rows, enc, 1, skew, d_sums);
    run<64, 3, 16, 16>(X, n_rows, enc, 1, skew, d_sums);
  }
  // narrow column slices of the d-sharded layout: 32-byte rows, the same table seen as 8 x as many rows
  fill_kernel<<<1024, 256>>>(X, (long long)bytes / 4, 8);
  CK(cudaDeviceSynchronize());
  run<8, 4, 8, 32>(X, n_rows * 8, enc, 4, 0, d_sums);
  run<8, 8, 8, 32>(X, n_rows * 8, enc, 4, 0, d_sums);
  fill_kernel<<<1024, 256>>>(X, (long long)bytes / 4, 16);
  CK(cudaDeviceSynchronize());
  run<16, 4, 8, 32>(X, n_rows * 4, enc, 4, 0, d_sums);
  return 0;
}
