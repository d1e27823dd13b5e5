// ngcf.cu -- the dense half of an NGCF layer, forward and backward, fused around the SpMM.
//
// Reference: recommender/NGCF.py:197-212 -- per layer
//     T = E W1 ;  E' = leaky_relu( A T + T + ((A E) * E) W2 , 0.01 )
// (torch.mm x 2, torch.sparse.mm x 2, four element-wise kernels, and autograd's mirror of all of it).
// A is linear, so A (E W1) = (A E) W1: with P = A E the layer is
//     Z = (P + E) W1 + (P * E) W2 = [P + E | P * E] [W1 ; W2] ,   E' = leaky_relu(Z)
// -- ONE propagation per layer instead of two (agcf_spmm_csr_f32 computes P), and one [N, 2d] x [2d, d] product whose
// left operand is formed on the fly from the rows of P and E while they are staged in shared memory.  The product is
// N x 2d x d (1.16 GFLOP per layer at the Gowalla shape, d = 64) next to a 44 us SpMM: fp32 FMA on CUDA cores with a
// register-blocked 64-row tile, exact fp32 like the reference's torch.mm (a TF32 tensor-core product would put 1e-3 on
// every activation; a 3 x TF32 split is not worth a kernel this small).
//
//   agcf_ngcf_dense_forward    E' (and the running layer mean)               from P, E, W = [W1 ; W2]
//   agcf_ngcf_dense_backward   dP, dE_direct and per-CTA partials of dW       from dE', E' (sign of Z), P, E, W^T
//   agcf_ngcf_reduce_wgrad     dW = sum of the partials in CTA order (deterministic)
//
// Backward:  dZ = dE' * (E' > 0 ? 1 : 0.01)      (leaky_relu keeps the sign, so E' tells which branch Z took)
//            [dA | dB] = dZ [W1 ; W2]^T ;  dP = dA + dB * E ;  dE_direct = dA + dB * P ;  dW = [P + E | P * E]^T dZ
//            dE = A dP + dE_direct (the caller's next agcf_spmm_csr_f32, A = A^T)
#include "common.cuh"

namespace agcf {
namespace ngcf {

constexpr int kTM = 64;        // rows per tile
constexpr int kKC = 64;        // k chunk staged in shared memory
constexpr float kSlope = 0.01f;

template <int CN>
__device__ __forceinline__ void load_cols(const float* __restrict__ src, float (&w)[CN]) {
  if constexpr (CN == 2) {
    const float2 v = *reinterpret_cast<const float2*>(src);
    w[0] = v.x; w[1] = v.y;
  } else {
#pragma unroll
    for (int q = 0; q < CN / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(src + 4 * q);
      w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
    }
  }
}

// acc[4][CN] += As[ty*4 + r][0 .. KC) . Bs[0 .. KC)[tx*CN + c]      (As rows padded to KC + 4 floats)
template <int D, int CN>
__device__ __forceinline__ void tile_mac(const float (*As)[kKC + 4], const float* __restrict__ Bs, int ldb, int ty, int tx,
                                         float (&acc)[4][CN]) {
#pragma unroll 4
  for (int k = 0; k < kKC; k += 4) {
    float4 a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(&As[ty * 4 + r][k]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float w[CN];
      load_cols<CN>(Bs + (size_t)(k + kk) * ldb + tx * CN, w);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float av = kk == 0 ? a[r].x : (kk == 1 ? a[r].y : (kk == 2 ? a[r].z : a[r].w));
#pragma unroll
        for (int c = 0; c < CN; ++c) acc[r][c] = fmaf(av, w[c], acc[r][c]);
      }
    }
  }
}

// ---------------------------------------------------------------------------- forward
template <int D>
__global__ void __launch_bounds__(256) dense_fwd_kernel(const float* __restrict__ P, const float* __restrict__ E,
                                                        const float* __restrict__ W, float* __restrict__ Enext,
                                                        const float* __restrict__ acc_in, float* __restrict__ acc_out,
                                                        float acc_div, int N) {
  constexpr int CN = D / 16;                       // output columns per thread
  __shared__ __align__(16) float As[kTM][kKC + 4];
  __shared__ __align__(16) float Ws[kKC * D];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * kTM;
  float acc[4][CN];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < CN; ++c) acc[r][c] = 0.f;
  for (int kc = 0; kc < 2 * D; kc += kKC) {
    for (int f = tid; f < kTM * (kKC / 4); f += 256) {
      const int r = f / (kKC / 4), k4 = f - r * (kKC / 4);
      const int kk = kc + 4 * k4;
      const int part = kk / D, col = kk - part * D;                  // 4 consecutive k stay inside one part (D % 4 == 0)
      const int row = row0 + r;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f), e = p;
      if (row < N) {
        p = __ldg(reinterpret_cast<const float4*>(P + (size_t)row * D + col));
        e = __ldg(reinterpret_cast<const float4*>(E + (size_t)row * D + col));
      }
      const float4 v = part == 0 ? make_float4(p.x + e.x, p.y + e.y, p.z + e.z, p.w + e.w)
                                 : make_float4(p.x * e.x, p.y * e.y, p.z * e.z, p.w * e.w);
      *reinterpret_cast<float4*>(&As[r][4 * k4]) = v;
    }
    for (int f = tid; f < kKC * D / 4; f += 256)
      reinterpret_cast<float4*>(Ws)[f] = __ldg(reinterpret_cast<const float4*>(W + (size_t)kc * D) + f);
    __syncthreads();
    tile_mac<D, CN>(As, Ws, D, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = row0 + ty * 4 + r;
    if (row >= N) continue;
    const size_t at = (size_t)row * D + tx * CN;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
      const float z = acc[r][c];
      const float o = z > 0.f ? z : kSlope * z;                      // F.leaky_relu(z, 0.01)
      Enext[at + c] = o;
      if (acc_out != nullptr) {
        float m = (acc_in != nullptr ? acc_in[at + c] : 0.f) + o;
        if (acc_div != 1.0f) m = __fdiv_rn(m, acc_div);
        acc_out[at + c] = m;
      }
    }
  }
}

// --------------------------------------------------------------------------- backward
template <int D>
struct BwdSmem {
  static constexpr int LD = D + 4;
  static constexpr size_t bytes = (size_t)(3 * kTM * LD + kKC * D) * 4 + (size_t)kTM * (kKC + 4) * 4;
};

template <int D>
__global__ void __launch_bounds__(256, 2) dense_bwd_kernel(const float* __restrict__ dOut, const float* __restrict__ Enext,
                                                           const float* __restrict__ P, const float* __restrict__ E,
                                                           const float* __restrict__ WT, float* __restrict__ dP,
                                                           float* __restrict__ dEdir, float* __restrict__ dWpart, int N) {
  constexpr int CN = D / 16, LD = D + 4;
  constexpr int JW = 2 * D / 16, CW = D / 16;       // weight-gradient block of a thread: JW rows of dW x CW columns
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  float* Ps = reinterpret_cast<float*>(bwd_smem);                    // [TM][LD]
  float* Es = Ps + kTM * LD;
  float* Zs = Es + kTM * LD;
  float* Bs = Zs + kTM * LD;                                         // [KC][D] chunk of W^T columns
  float (*As)[kKC + 4] = reinterpret_cast<float (*)[kKC + 4]>(Bs + kKC * D);   // [TM][KC + 4]: dZ chunk for the products
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int jy = tid >> 4, cx = tid & 15;
  float wacc[JW][CW];
#pragma unroll
  for (int a = 0; a < JW; ++a)
#pragma unroll
    for (int c = 0; c < CW; ++c) wacc[a][c] = 0.f;
  const int n_tiles = (N + kTM - 1) / kTM;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int row0 = tile * kTM;
    for (int f = tid; f < kTM * (D / 4); f += 256) {
      const int r = f / (D / 4), c4 = f - r * (D / 4);
      const int row = row0 + r;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f), e = p, z = p;
      if (row < N) {
        const size_t at = (size_t)row * D + 4 * c4;
        p = __ldg(reinterpret_cast<const float4*>(P + at));
        e = __ldg(reinterpret_cast<const float4*>(E + at));
        const float4 g = __ldg(reinterpret_cast<const float4*>(dOut + at));
        const float4 y = __ldg(reinterpret_cast<const float4*>(Enext + at));
        z = make_float4(y.x > 0.f ? g.x : kSlope * g.x, y.y > 0.f ? g.y : kSlope * g.y,
                        y.z > 0.f ? g.z : kSlope * g.z, y.w > 0.f ? g.w : kSlope * g.w);
      }
      *reinterpret_cast<float4*>(Ps + r * LD + 4 * c4) = p;
      *reinterpret_cast<float4*>(Es + r * LD + 4 * c4) = e;
      *reinterpret_cast<float4*>(Zs + r * LD + 4 * c4) = z;
    }
    __syncthreads();
    // [dA | dB] = dZ W^T: two products with the same left operand; K = D in chunks of KC
    float accA[4][CN], accB[4][CN];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < CN; ++c) { accA[r][c] = 0.f; accB[r][c] = 0.f; }
    for (int half = 0; half < 2; ++half) {
      for (int kc = 0; kc < D; kc += kKC) {
        const int kn = D - kc < kKC ? D - kc : kKC;                  // D = 32: one chunk of 32
        for (int f = tid; f < kTM * (kKC / 4); f += 256) {
          const int r = f / (kKC / 4), k4 = f - r * (kKC / 4);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * k4 < kn) v = *reinterpret_cast<const float4*>(Zs + r * LD + kc + 4 * k4);
          *reinterpret_cast<float4*>(&As[r][4 * k4]) = v;
        }
        for (int f = tid; f < kKC * D / 4; f += 256) {
          const int k = f / (D / 4), j4 = f - k * (D / 4);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < kn) v = __ldg(reinterpret_cast<const float4*>(WT + (size_t)(kc + k) * 2 * D + half * D + 4 * j4));
          reinterpret_cast<float4*>(Bs)[f] = v;
        }
        __syncthreads();
        if (half == 0) tile_mac<D, CN>(As, Bs, D, ty, tx, accA);
        else tile_mac<D, CN>(As, Bs, D, ty, tx, accB);
        __syncthreads();
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int lr = ty * 4 + r, row = row0 + lr;
      if (row >= N) continue;
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        const int col = tx * CN + c;
        const float e = Es[lr * LD + col], p = Ps[lr * LD + col];
        dP[(size_t)row * D + col] = fmaf(accB[r][c], e, accA[r][c]);
        dEdir[(size_t)row * D + col] = fmaf(accB[r][c], p, accA[r][c]);
      }
    }
    // dW[j'][c] += sum_n A'[n][j'] dZ[n][c]   with A' = [P + E | P * E] formed from the staged rows
    const int j0 = jy * JW, part = j0 / D, col0 = j0 - part * D;     // JW consecutive rows of dW stay inside one part
#pragma unroll 2
    for (int n = 0; n < kTM; ++n) {
      float a[JW], z[CW];
      load_cols<JW>(Ps + n * LD + col0, a);
      {
        float e[JW];
        load_cols<JW>(Es + n * LD + col0, e);
#pragma unroll
        for (int q = 0; q < JW; ++q) a[q] = part == 0 ? a[q] + e[q] : a[q] * e[q];
      }
      load_cols<CW>(Zs + n * LD + cx * CW, z);
#pragma unroll
      for (int q = 0; q < JW; ++q)
#pragma unroll
        for (int c = 0; c < CW; ++c) wacc[q][c] = fmaf(a[q], z[c], wacc[q][c]);
    }
    __syncthreads();                                                 // the next tile overwrites the staged rows
  }
  float* out = dWpart + (size_t)blockIdx.x * 2 * D * D;
#pragma unroll
  for (int q = 0; q < JW; ++q)
#pragma unroll
    for (int c = 0; c < CW; ++c) out[(size_t)(jy * JW + q) * D + cx * CW + c] = wacc[q][c];
}

__global__ void __launch_bounds__(256) reduce_wgrad_kernel(const float* __restrict__ part, int n_part, int n, float* __restrict__ dW) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= n) return;
  float s = 0.f;
  for (int b = 0; b < n_part; ++b) s += part[(size_t)b * n + k];      // CTA order: deterministic
  dW[k] = s;
}

}  // namespace ngcf
}  // namespace agcf

using namespace agcf;

extern "C" int agcf_ngcf_dense_forward(const float* P, const float* E, const float* W, float* Enext,
                                       const float* acc_in, float* acc_out, float acc_div,
                                       int32_t n_rows, int32_t d, agcf_stream_t stream) {
  if (!P || !E || !W || !Enext || n_rows < 0 || acc_div == 0.f) return AGCF_EINVAL;
  if (d != 32 && d != 64) return AGCF_EUNSUPPORTED;
  if (!aligned16(P) || !aligned16(E) || !aligned16(W) || !aligned16(Enext) || !aligned16(acc_in) || !aligned16(acc_out)) return AGCF_EINVAL;
  if (Enext == P || Enext == E) return AGCF_EINVAL;
  if (n_rows == 0) return AGCF_OK;
  const unsigned blocks = (unsigned)((n_rows + ngcf::kTM - 1) / ngcf::kTM);
  cudaStream_t st = (cudaStream_t)stream;
  if (d == 64) ngcf::dense_fwd_kernel<64><<<blocks, 256, 0, st>>>(P, E, W, Enext, acc_in, acc_out, acc_div, n_rows);
  else ngcf::dense_fwd_kernel<32><<<blocks, 256, 0, st>>>(P, E, W, Enext, acc_in, acc_out, acc_div, n_rows);
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}

extern "C" int agcf_ngcf_dense_backward(const float* dOut, const float* Enext, const float* P, const float* E,
                                        const float* WT, float* dP, float* dEdir, float* dW_partial,
                                        int32_t n_partials, int32_t n_rows, int32_t d, agcf_stream_t stream) {
  if (!dOut || !Enext || !P || !E || !WT || !dP || !dEdir || !dW_partial || n_rows < 0 || n_partials < 1) return AGCF_EINVAL;
  if (d != 32 && d != 64) return AGCF_EUNSUPPORTED;
  if (!aligned16(dOut) || !aligned16(Enext) || !aligned16(P) || !aligned16(E) || !aligned16(WT) || !aligned16(dP) ||
      !aligned16(dEdir) || !aligned16(dW_partial))
    return AGCF_EINVAL;
  if (dP == dOut || dEdir == dOut || dP == dEdir) return AGCF_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define AGCF_NGCF_BWD(DD)                                                                                      \
  {                                                                                                            \
    AGCF_CUDA_OK(cudaFuncSetAttribute(ngcf::dense_bwd_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)ngcf::BwdSmem<DD>::bytes));                                         \
    ngcf::dense_bwd_kernel<DD><<<(unsigned)n_partials, 256, ngcf::BwdSmem<DD>::bytes, st>>>(                   \
        dOut, Enext, P, E, WT, dP, dEdir, dW_partial, n_rows);                                                 \
  }
  if (d == 64) AGCF_NGCF_BWD(64) else AGCF_NGCF_BWD(32)
#undef AGCF_NGCF_BWD
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}

extern "C" int agcf_ngcf_reduce_wgrad(const float* dW_partial, int32_t n_partials, float* dW, int32_t d,
                                      agcf_stream_t stream) {
  if (!dW_partial || !dW || n_partials < 1) return AGCF_EINVAL;
  if (d != 32 && d != 64) return AGCF_EUNSUPPORTED;
  const int n = 2 * d * d;
  ngcf::reduce_wgrad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dW_partial, n_partials, n, dW);
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}
