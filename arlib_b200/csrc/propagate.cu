// propagate.cu -- normalized-adjacency propagation for graph-CF encoders on B200.
//
//   agcf_norm_adj_csr      values of D^-1/2 A D^-1/2 (bit-exact association)
//   agcf_spmm_csr_f32      Y = A X with fused layer-sum / addend / noise epilogue
//   agcf_sddmm_csr_f32     pattern-masked dL/dA
//   agcf_concat_rows_f32   [user_emb; item_emb] packing
//
// Reference call sites: recommender/LightGCN.py:230-240 (forward), its autograd
// (backward), recommender/SimGCL.py:198-210, XSimGCL.py:205-223 (noise),
// attack/White/PGA.py:97-117 (adjacency gradient), util/DataLoader.py:73-87 and
// recommender/LightGCN.py:212-215 (normalization).
//
// SpMM design (HBM/L2-bound gather, no tensor cores -- see DESIGN.md):
//   * a row of X is d fp32 = d/4 float4; LPR = min(d/4, 32) lanes own one row and
//     each lane loads 16 B, so one warp-level LDG.128 fetches whole 128 B lines of
//     32/LPR different neighbour rows (fully coalesced sectors);
//   * rows are processed in `row_order` (degree descending): the two rows sharing a
//     warp have ~equal length, long rows start first, the tail is 1-nnz rows;
//   * col/val of a row are read 16/32 at a time, coalesced, streaming (no L1
//     allocation) and broadcast by shuffle; up to LPR independent row gathers are
//     in flight per lane before the FMA chain consumes them;
//   * the first n_long rows (degree > a threshold chosen by the host) get a whole
//     CTA: 256/LPR lane groups stride over the row, partials are reduced through
//     shared memory in a fixed order (no atomics, deterministic);
//   * epilogue fuses: + addend (backward's G/(L+1)), SimGCL noise, Y store, and
//     the running layer sum / mean (acc_out = (acc_in + t) / acc_div).
#include "common.cuh"
#include <stdlib.h>

namespace agcf {

struct SpmmParams {
  const int32_t* rowptr;
  const int32_t* col;
  const float* val;
  const float4* X;
  float4* Y;
  const float4* addend;
  const float4* acc_in;
  float4* acc_out;
  float acc_div;
  const float4* noise;
  float eps;
  const int32_t* row_order;
  int32_t n_long;
  int32_t n_rows;
  const uint32_t* row_mask;   // nullable bitmap over rows: only rows with their bit set are computed / written
  const uint32_t* col_mask;   // nullable bitmap over columns: rows of X outside it are known to be zero (skipped)
  // fused all-gather: the same rows are also stored into the peer GPUs' copies of Y / acc_out
  // (NVLink P2P stores straight from the epilogue; peers see them after the next barrier)
  int n_peers;
  float4* peer_Y[AGCF_MAX_PEERS];
  float4* peer_acc[AGCF_MAX_PEERS];
};

__device__ __forceinline__ bool bit_set(const uint32_t* __restrict__ m, int k) {
  return (__ldg(m + (k >> 5)) >> (k & 31)) & 1u;
}

constexpr int default_lpr(int d) { return d / 8 < 8 ? 8 : (d / 8 > 32 ? 32 : d / 8); }   // 8 (d<=64), 16 (128), 32 (256)

template <int D, int LPR_ = default_lpr(D)>
struct RowCfg {
  static constexpr int V4 = D / 4;                    // float4 per row
  static constexpr int LPR = LPR_;                    // lanes per row
  static constexpr int VPL = V4 / LPR;                // float4 per lane
  static constexpr int RPW = 32 / LPR;                // rows per warp
  static constexpr int THREADS = 256;
  static constexpr int RPB = (THREADS / 32) * RPW;    // rows per block (short path)
  static constexpr int GROUPS = THREADS / LPR;        // lane groups per block (long path)
  static_assert(V4 % LPR == 0 && LPR <= 32 && VPL >= 1, "bad lane mapping");
};

__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }

template <typename C, bool NOISE>
__device__ __forceinline__ void spmm_epilogue(const SpmmParams& p, int row, bool valid,
                                              float4 (&t)[C::VPL], int gl) {
  const size_t rbase = (size_t)row * C::V4;
  if (p.addend != nullptr && valid) {
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) t[v] = add4(t[v], ld_stream_f4(p.addend + rbase + v * C::LPR + gl));
  }
  if (NOISE) {
    float4 nz[C::VPL];
    float ss = 0.f;
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) {
      nz[v] = valid ? ld_stream_f4(p.noise + rbase + v * C::LPR + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss += dot4(nz[v], nz[v]);
    }
    ss = group_sum<C::LPR>(ss);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);   // F.normalize(dim=-1, eps=1e-12)
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) {
      t[v].x = t[v].x + __fmul_rn(__fmul_rn(sgnf(t[v].x), __fdiv_rn(nz[v].x, nrm)), p.eps);
      t[v].y = t[v].y + __fmul_rn(__fmul_rn(sgnf(t[v].y), __fdiv_rn(nz[v].y, nrm)), p.eps);
      t[v].z = t[v].z + __fmul_rn(__fmul_rn(sgnf(t[v].z), __fdiv_rn(nz[v].z, nrm)), p.eps);
      t[v].w = t[v].w + __fmul_rn(__fmul_rn(sgnf(t[v].w), __fdiv_rn(nz[v].w, nrm)), p.eps);
    }
  }
  if (!valid) return;
  if (p.Y != nullptr) {
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) p.Y[rbase + v * C::LPR + gl] = t[v];
    for (int q = 0; q < p.n_peers; ++q) {
      if (p.peer_Y[q] == nullptr) continue;
#pragma unroll
      for (int v = 0; v < C::VPL; ++v) p.peer_Y[q][rbase + v * C::LPR + gl] = t[v];
    }
  }
  if (p.acc_out != nullptr) {
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.acc_in != nullptr) a = ld_stream_f4(p.acc_in + rbase + v * C::LPR + gl);
      float4 o = add4(a, t[v]);
      if (p.acc_div != 1.0f) {
        o.x = __fdiv_rn(o.x, p.acc_div); o.y = __fdiv_rn(o.y, p.acc_div);
        o.z = __fdiv_rn(o.z, p.acc_div); o.w = __fdiv_rn(o.w, p.acc_div);
      }
      p.acc_out[rbase + v * C::LPR + gl] = o;
      for (int q = 0; q < p.n_peers; ++q)
        if (p.peer_acc[q] != nullptr) p.peer_acc[q][rbase + v * C::LPR + gl] = o;
    }
  }
}

// accumulate nnz [s + first, s + first + stride*k ...) chunk-wise; `len` = row length,
// `first`/`stride` = this lane group's starting offset and per-iteration advance,
// `iters` and `bound_len` are WARP-UNIFORM (shuffles use the full mask): bound_len is
// an upper bound of (len - first) over the warp's groups.
//   MODE 0: inner loop fully unrolled over the LPR slots, loads predicated on (t < n)
//   MODE 1: inner loop runs min(LPR, bound) times (warp-uniform), branch on (ct >= 0)
//   MODE 2: as 1 but in explicit sub-blocks of 4 slots: all gathers first, then the FMAs
template <typename C, int MODE>
__device__ __forceinline__ void spmm_accumulate(const SpmmParams& p, int s, int len, int first, int stride,
                                                int iters, int bound_len, int gl, float4 (&acc)[C::VPL]) {
  int c_next = -1;
  float v_next = 0.f;
  if (first + gl < len) {
    c_next = ld_stream_i32(p.col + s + first + gl);
    v_next = ld_stream_f32(p.val + s + first + gl);
    if (p.col_mask != nullptr && !bit_set(p.col_mask, c_next)) c_next = -1;
  }
  int off = first;
  int bound = bound_len;
  for (int it = 0; it < iters; ++it, off += stride, bound -= stride) {
    const int c = c_next;
    const float v = v_next;
    c_next = -1;
    v_next = 0.f;
    if (off + stride + gl < len) {                       // prefetch the next chunk's indices
      c_next = ld_stream_i32(p.col + s + off + stride + gl);
      v_next = ld_stream_f32(p.val + s + off + stride + gl);
      if (p.col_mask != nullptr && !bit_set(p.col_mask, c_next)) c_next = -1;
    }
    if (MODE == 0 || MODE == 4) {
#pragma unroll
      for (int t = 0; t < C::LPR; ++t) {
        const int ct = __shfl_sync(0xffffffffu, c, t, C::LPR);
        const float vt = __shfl_sync(0xffffffffu, v, t, C::LPR);
        if (ct >= 0) {
          const float4* xr = p.X + (size_t)ct * C::V4 + gl;
#pragma unroll
          for (int vv = 0; vv < C::VPL; ++vv) {
            if (MODE == 4) fma4_packed(acc[vv], vt, ld_gather_f4(xr + vv * C::LPR));
            else fma4(acc[vv], vt, ld_gather_f4(xr + vv * C::LPR));
          }
        }
      }
    } else if (MODE == 1) {
      const int nmax = bound < C::LPR ? bound : C::LPR;   // warp-uniform
#pragma unroll 4
      for (int t = 0; t < nmax; ++t) {
        const int ct = __shfl_sync(0xffffffffu, c, t, C::LPR);
        const float vt = __shfl_sync(0xffffffffu, v, t, C::LPR);
        if (ct >= 0) {                                     // lanes past their row's end hold c = -1
          const float4* xr = p.X + (size_t)ct * C::V4 + gl;
#pragma unroll
          for (int vv = 0; vv < C::VPL; ++vv) fma4(acc[vv], vt, ld_gather_f4(xr + vv * C::LPR));
        }
      }
    } else if (MODE == 3) {
      // quad-blocked: warp-uniform branch per 4 slots (padding <= 3 slots), slots predicated,
      // packed FFMA2 accumulation
      const int nmax = bound < C::LPR ? bound : C::LPR;   // warp-uniform
#pragma unroll
      for (int q = 0; q < C::LPR / 4; ++q) {
        if (4 * q < nmax) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int ct = __shfl_sync(0xffffffffu, c, 4 * q + r, C::LPR);
            const float vt = __shfl_sync(0xffffffffu, v, 4 * q + r, C::LPR);
            if (ct >= 0) {
              const float4* xr = p.X + (size_t)ct * C::V4 + gl;
#pragma unroll
              for (int vv = 0; vv < C::VPL; ++vv) fma4_packed(acc[vv], vt, ld_gather_f4(xr + vv * C::LPR));
            }
          }
        }
      }
    } else {
      const int nmax = bound < C::LPR ? bound : C::LPR;   // warp-uniform
      for (int t0 = 0; t0 < nmax; t0 += 4) {
        float4 x[4][C::VPL];
        float vt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ct = __shfl_sync(0xffffffffu, c, (t0 + q) & (C::LPR - 1), C::LPR);
          vt[q] = __shfl_sync(0xffffffffu, v, (t0 + q) & (C::LPR - 1), C::LPR);
          const bool live = ct >= 0 && (t0 + q) < C::LPR;
          if (!live) vt[q] = 0.f;
          const float4* xr = p.X + (size_t)(live ? ct : 0) * C::V4 + gl;
#pragma unroll
          for (int vv = 0; vv < C::VPL; ++vv)
            x[q][vv] = live ? ld_gather_f4(xr + vv * C::LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int vv = 0; vv < C::VPL; ++vv) fma4(acc[vv], vt[q], x[q][vv]);
      }
    }
  }
}

template <int D, int LPR, int MODE, int MINB, bool NOISE>
__global__ void __launch_bounds__(256, MINB) spmm_csr_kernel(const SpmmParams p) {
  using C = RowCfg<D, LPR>;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (C::LPR - 1);
  float4 acc[C::VPL];
#pragma unroll
  for (int v = 0; v < C::VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  if ((int)blockIdx.x < p.n_long) {
    // ---- long row: the whole CTA cooperates on one row --------------------------
    __shared__ float4 part[C::GROUPS][C::V4];
    const int row = p.row_order != nullptr ? p.row_order[blockIdx.x] : (int)blockIdx.x;
    if (p.row_mask != nullptr && !bit_set(p.row_mask, row)) return;      // block-uniform
    const int s = p.rowptr[blockIdx.x];
    const int len = p.rowptr[blockIdx.x + 1] - s;
    const int g = threadIdx.x / C::LPR;
    const int stride = C::GROUPS * C::LPR;
    const int iters = (len + stride - 1) / stride;       // block-uniform
    spmm_accumulate<C, MODE>(p, s, len, g * C::LPR, stride, iters, len, gl, acc);
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) part[g][v * C::LPR + gl] = acc[v];
    __syncthreads();
    if (warp == 0) {
      const bool valid = lane < C::LPR;
      float4 t[C::VPL];
#pragma unroll
      for (int v = 0; v < C::VPL; ++v) {
        t[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          for (int gg = 0; gg < C::GROUPS; ++gg) t[v] = add4(t[v], part[gg][v * C::LPR + gl]);
        }
      }
      spmm_epilogue<C, NOISE>(p, row, valid, t, gl);
    }
    return;
  }

  // ---- short rows: one lane group per row ----------------------------------------
  const int grp = lane / C::LPR;
  const long long slot = (long long)p.n_long + ((long long)(blockIdx.x - p.n_long) * (C::THREADS / 32) + warp) * C::RPW + grp;
  bool valid = slot < p.n_rows;
  int row = 0, s = 0, len = 0;
  if (valid) {
    row = p.row_order != nullptr ? p.row_order[slot] : (int)slot;
    if (p.row_mask != nullptr && !bit_set(p.row_mask, row)) {
      valid = false;
    } else {
      s = p.rowptr[slot];
      len = p.rowptr[slot + 1] - s;
    }
  }
  int maxlen = len;
#pragma unroll
  for (int o = C::LPR; o < 32; o <<= 1) {
    const int other = __shfl_xor_sync(0xffffffffu, maxlen, o);
    maxlen = other > maxlen ? other : maxlen;
  }
  const int iters = (maxlen + C::LPR - 1) / C::LPR;      // warp-uniform
  spmm_accumulate<C, MODE>(p, s, len, 0, C::LPR, iters, maxlen, gl, acc);
  spmm_epilogue<C, NOISE>(p, row, valid, acc, gl);
}

// ---------------------------------------------------------------- persistent kernel
// The row-per-group kernel above is latency-bound: per task a warp walks a chain of
// DEPENDENT global loads (row id -> rowptr -> col/val -> gathers -> epilogue operand)
// and only the gather phase moves real data.  This version keeps every warp resident
// and software-pipelines that chain ACROSS tasks: while task t is gathering, the
// metadata of task t+2 and the first col/val chunk of task t+1 are already in flight.
struct TaskMeta {
  int row, s, len;
  bool valid;
};

template <typename C>
__device__ __forceinline__ TaskMeta load_task_meta(const SpmmParams& p, long long task, long long n_tasks, int grp) {
  TaskMeta m;
  m.row = 0; m.s = 0; m.len = 0; m.valid = false;
  const long long slot = (long long)p.n_long + task * C::RPW + grp;
  if (task < n_tasks && slot < p.n_rows) {
    m.row = p.row_order != nullptr ? __ldg(p.row_order + slot) : (int)slot;
    m.s = __ldg(p.rowptr + slot);
    m.len = __ldg(p.rowptr + slot + 1) - m.s;
    m.valid = true;
  }
  return m;
}

template <typename C>
__device__ __forceinline__ void load_first_chunk(const SpmmParams& p, TaskMeta& m, int gl, int& c, float& v) {
  // the row mask is resolved here (one stage after the row id arrived)
  if (m.valid && p.row_mask != nullptr && !bit_set(p.row_mask, m.row)) { m.valid = false; m.len = 0; }
  c = -1;
  v = 0.f;
  if (gl < m.len) {
    c = ld_stream_i32(p.col + m.s + gl);
    v = ld_stream_f32(p.val + m.s + gl);
    if (p.col_mask != nullptr && !bit_set(p.col_mask, c)) c = -1;
  }
}

template <int D, int LPR, int MINB, bool NOISE>
__global__ void __launch_bounds__(256, MINB) spmm_pipe_kernel(const SpmmParams p) {
  using C = RowCfg<D, LPR>;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (C::LPR - 1);

  if ((int)blockIdx.x < p.n_long) {
    // ---- long row: the whole CTA cooperates on one row (same as the simple kernel) ----
    __shared__ float4 part[C::GROUPS][C::V4];
    float4 acc[C::VPL];
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int row = p.row_order != nullptr ? p.row_order[blockIdx.x] : (int)blockIdx.x;
    if (p.row_mask != nullptr && !bit_set(p.row_mask, row)) return;
    const int s = p.rowptr[blockIdx.x];
    const int len = p.rowptr[blockIdx.x + 1] - s;
    const int g = threadIdx.x / C::LPR;
    const int stride = C::GROUPS * C::LPR;
    const int iters = (len + stride - 1) / stride;
    spmm_accumulate<C, 0>(p, s, len, g * C::LPR, stride, iters, len, gl, acc);
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) part[g][v * C::LPR + gl] = acc[v];
    __syncthreads();
    if (warp == 0) {
      const bool valid = lane < C::LPR;
      float4 t[C::VPL];
#pragma unroll
      for (int v = 0; v < C::VPL; ++v) {
        t[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          for (int gg = 0; gg < C::GROUPS; ++gg) t[v] = add4(t[v], part[gg][v * C::LPR + gl]);
        }
      }
      spmm_epilogue<C, NOISE>(p, row, valid, t, gl);
    }
    return;
  }

  const int grp = lane / C::LPR;
  const long long n_short = (long long)p.n_rows - p.n_long;
  const long long n_tasks = (n_short + C::RPW - 1) / C::RPW;
  const long long W = (long long)(gridDim.x - p.n_long) * (C::THREADS / 32);
  long long task = (long long)(blockIdx.x - p.n_long) * (C::THREADS / 32) + warp;
  if (task >= n_tasks) return;                                  // warp-uniform
  TaskMeta m0 = load_task_meta<C>(p, task, n_tasks, grp);
  TaskMeta m1 = load_task_meta<C>(p, task + W, n_tasks, grp);
  int c0; float v0;
  load_first_chunk<C>(p, m0, gl, c0, v0);
  for (; task < n_tasks; task += W) {
    TaskMeta m2 = load_task_meta<C>(p, task + 2 * W, n_tasks, grp);      // two tasks ahead
    int c1; float v1;
    load_first_chunk<C>(p, m1, gl, c1, v1);                                // one task ahead
    // ---- current task -------------------------------------------------------------
    float4 acc[C::VPL];
#pragma unroll
    for (int v = 0; v < C::VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int maxlen = m0.len;
#pragma unroll
    for (int o = C::LPR; o < 32; o <<= 1) {
      const int other = __shfl_xor_sync(0xffffffffu, maxlen, o);
      maxlen = other > maxlen ? other : maxlen;
    }
    int c = c0;
    float v = v0;
    for (int off = 0; off < maxlen; off += C::LPR) {
      int cn = -1;
      float vn = 0.f;
      if (off + C::LPR + gl < m0.len) {                          // next chunk of this row
        cn = ld_stream_i32(p.col + m0.s + off + C::LPR + gl);
        vn = ld_stream_f32(p.val + m0.s + off + C::LPR + gl);
        if (p.col_mask != nullptr && !bit_set(p.col_mask, cn)) cn = -1;
      }
#pragma unroll
      for (int t = 0; t < C::LPR; ++t) {
        const int ct = __shfl_sync(0xffffffffu, c, t, C::LPR);
        const float vt = __shfl_sync(0xffffffffu, v, t, C::LPR);
        if (ct >= 0) {
          const float4* xr = p.X + (size_t)ct * C::V4 + gl;
#pragma unroll
          for (int vv = 0; vv < C::VPL; ++vv) fma4(acc[vv], vt, ld_gather_f4(xr + vv * C::LPR));
        }
      }
      c = cn;
      v = vn;
    }
    spmm_epilogue<C, NOISE>(p, m0.row, m0.valid, acc, gl);
    m0 = m1; c0 = c1; v0 = v1; m1 = m2;
  }
}

template <int D, int LPR, int MINB>
static int launch_spmm_pipe(const SpmmParams& p, cudaStream_t st) {
  using C = RowCfg<D, LPR>;
  static int ctas_per_sm[2] = {0, 0};
  const int which = p.noise != nullptr ? 1 : 0;
  if (ctas_per_sm[which] == 0) {
    int n = 0;
    cudaError_t e = which ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmm_pipe_kernel<D, LPR, MINB, true>, 256, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spmm_pipe_kernel<D, LPR, MINB, false>, 256, 0);
    if (e != cudaSuccess) return cuda_fail(e);
    ctas_per_sm[which] = n > 0 ? n : 1;
  }
  const long long n_short = (long long)p.n_rows - p.n_long;
  const long long n_tasks = (n_short + C::RPW - 1) / C::RPW;
  long long persistent = (long long)kSMs * ctas_per_sm[which];
  const long long need = (n_tasks + 7) / 8;
  if (persistent > need) persistent = need;
  const long long blocks = p.n_long + persistent;
  if (blocks <= 0) return AGCF_OK;
  if (p.noise != nullptr)
    spmm_pipe_kernel<D, LPR, MINB, true><<<(unsigned)blocks, 256, 0, st>>>(p);
  else
    spmm_pipe_kernel<D, LPR, MINB, false><<<(unsigned)blocks, 256, 0, st>>>(p);
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}

template <int D, int LPR, int MODE, int MINB>
static int launch_spmm_cfg(const SpmmParams& p, cudaStream_t st) {
  using C = RowCfg<D, LPR>;
  const long long short_rows = (long long)p.n_rows - p.n_long;
  const long long blocks = p.n_long + (short_rows + C::RPB - 1) / C::RPB;
  if (blocks <= 0) return AGCF_OK;
  if (blocks > 0x7fffffffLL) return AGCF_EUNSUPPORTED;
  if (p.noise != nullptr)
    spmm_csr_kernel<D, LPR, MODE, MINB, true><<<(unsigned)blocks, C::THREADS, 0, st>>>(p);
  else
    spmm_csr_kernel<D, LPR, MODE, MINB, false><<<(unsigned)blocks, C::THREADS, 0, st>>>(p);
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}

// tuning hook (profiling only): AGCF_SPMM_VARIANT selects the lane mapping / inner loop for d = 64
static int spmm_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AGCF_SPMM_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int D>
static int launch_spmm(const SpmmParams& p, cudaStream_t st) {
  return launch_spmm_cfg<D, (D / 4 < 32 ? D / 4 : 32), 4, 1>(p, st);
}

template <>
int launch_spmm<64>(const SpmmParams& p, cudaStream_t st) {
  switch (spmm_variant()) {
    case 1: return launch_spmm_cfg<64, 8, 1, 1>(p, st);
    case 2: return launch_spmm_cfg<64, 8, 2, 1>(p, st);
    case 3: return launch_spmm_cfg<64, 16, 1, 1>(p, st);
    case 4: return launch_spmm_cfg<64, 16, 2, 1>(p, st);
    case 5: return launch_spmm_cfg<64, 8, 0, 1>(p, st);
    case 6: return launch_spmm_cfg<64, 16, 0, 6>(p, st);
    case 7: return launch_spmm_cfg<64, 16, 2, 6>(p, st);
    case 8: return launch_spmm_cfg<64, 8, 2, 5>(p, st);
    case 9: return launch_spmm_cfg<64, 16, 0, 8>(p, st);
    case 20: return launch_spmm_pipe<64, 16, 1>(p, st);
    case 21: return launch_spmm_pipe<64, 16, 4>(p, st);
    case 22: return launch_spmm_pipe<64, 16, 5>(p, st);
    case 23: return launch_spmm_pipe<64, 8, 1>(p, st);
    case 24: return launch_spmm_pipe<64, 8, 4>(p, st);
    case 30: return launch_spmm_cfg<64, 16, 4, 1>(p, st);
    case 31: return launch_spmm_cfg<64, 16, 4, 4>(p, st);
    case 32: return launch_spmm_cfg<64, 16, 4, 5>(p, st);
    case 33: return launch_spmm_cfg<64, 16, 4, 6>(p, st);
    case 34: return launch_spmm_cfg<64, 16, 0, 5>(p, st);
    case 35: return launch_spmm_cfg<64, 16, 0, 4>(p, st);
    case 10: return launch_spmm_cfg<64, 16, 3, 1>(p, st);
    case 11: return launch_spmm_cfg<64, 8, 3, 1>(p, st);
    case 12: return launch_spmm_cfg<64, 16, 3, 6>(p, st);
    case 13: return launch_spmm_cfg<64, 8, 3, 4>(p, st);
    case 14: return launch_spmm_cfg<64, 8, 3, 5>(p, st);
    case 99: return launch_spmm_cfg<64, 16, 0, 1>(p, st);
    default: return launch_spmm_cfg<64, 16, 4, 1>(p, st);      // 16 lanes/row, full unroll, FFMA2: best measured (profiles/)
  }
}

// ------------------------------------------------------------------------ SDDMM
template <int D>
__global__ void __launch_bounds__(256) sddmm_csr_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                        const float4* __restrict__ H, const float4* __restrict__ E,
                                                        float* __restrict__ gval, int accumulate,
                                                        const int32_t* __restrict__ row_order, int n_rows) {
  using C = RowCfg<D>;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (C::LPR - 1);
  const int grp = lane / C::LPR;
  const long long slot = ((long long)blockIdx.x * (C::THREADS / 32) + warp) * C::RPW + grp;
  const bool valid = slot < n_rows;
  int row = 0, s = 0, len = 0;
  if (valid) {
    row = row_order != nullptr ? row_order[slot] : (int)slot;
    s = rowptr[row];
    len = rowptr[row + 1] - s;
  }
  int maxlen = len;
#pragma unroll
  for (int o = C::LPR; o < 32; o <<= 1) {
    const int other = __shfl_xor_sync(0xffffffffu, maxlen, o);
    maxlen = other > maxlen ? other : maxlen;
  }
  float4 h[C::VPL];
#pragma unroll
  for (int v = 0; v < C::VPL; ++v)
    h[v] = valid ? __ldg(H + (size_t)row * C::V4 + v * C::LPR + gl) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int off = 0; off < maxlen; off += C::LPR) {
    int c = 0;
    if (off + gl < len) c = ld_stream_i32(col + s + off + gl);
    int n = len - off;
    n = n < 0 ? 0 : (n > C::LPR ? C::LPR : n);
    float mine = 0.f;
#pragma unroll
    for (int t = 0; t < C::LPR; ++t) {
      const int ct = __shfl_sync(0xffffffffu, c, t, C::LPR);
      float partial = 0.f;
      if (t < n) {
#pragma unroll
        for (int vv = 0; vv < C::VPL; ++vv)
          partial += dot4(h[vv], ld_gather_f4(E + (size_t)ct * C::V4 + vv * C::LPR + gl));
      }
      partial = group_sum<C::LPR>(partial);
      if (gl == t) mine = partial;
    }
    if (off + gl < len) {
      float* dst = gval + s + off + gl;
      *dst = accumulate ? (*dst + mine) : mine;
    }
  }
}

// ------------------------------------------------------------------ normalization
__global__ void __launch_bounds__(256) norm_adj_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                       const float* __restrict__ w, const float* __restrict__ d_row,
                                                       const float* __restrict__ d_col, float* __restrict__ val,
                                                       int32_t* __restrict__ row_of, int n_rows, long long nnz) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  // row of nnz p: largest i with rowptr[i] <= p (empty rows are skipped by the search)
  int lo = 0, hi = n_rows;           // invariant: rowptr[lo] <= p < rowptr[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)__ldg(rowptr + mid) <= p) lo = mid; else hi = mid;
  }
  if (row_of != nullptr) row_of[p] = lo;
  if (val != nullptr) val[p] = __fmul_rn(__fmul_rn(d_row[lo], w[p]), d_col[col[p]]);
}

__global__ void __launch_bounds__(256) concat_rows_kernel(const float4* __restrict__ a, long long na4,
                                                          const float4* __restrict__ b, long long nb4,
                                                          float4* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < na4 + nb4; k += stride)
    out[k] = k < na4 ? ld_stream_f4(a + k) : ld_stream_f4(b + (k - na4));
}

}  // namespace agcf

using namespace agcf;

extern "C" int agcf_spmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* val,
                                 const float* X, float* Y, const float* addend,
                                 const float* acc_in, float* acc_out, float acc_div,
                                 const float* noise, float eps,
                                 const int32_t* row_order, int32_t n_long,
                                 const uint32_t* row_mask, const uint32_t* col_mask,
                                 void* const* peer_Y_host, void* const* peer_acc_host, int32_t n_peers,
                                 int32_t n_rows, int32_t d, agcf_stream_t stream) {
  if (n_peers < 0 || n_peers > AGCF_MAX_PEERS) return AGCF_EINVAL;
  if (!rowptr || !col || !val || !X || n_rows < 0 || (Y == nullptr && acc_out == nullptr)) return AGCF_EINVAL;
  if (!supported_d(d)) return AGCF_EUNSUPPORTED;
  if (n_long < 0 || n_long > n_rows || (n_long > 0 && row_order == nullptr)) return AGCF_EINVAL;
  if (!aligned16(X) || !aligned16(Y) || !aligned16(addend) || !aligned16(acc_in) || !aligned16(acc_out) || !aligned16(noise))
    return AGCF_EINVAL;
  if (X == Y || X == acc_out) return AGCF_EINVAL;        // rows of X are read by other CTAs
  if (acc_div == 0.f) return AGCF_EINVAL;
  SpmmParams p;
  p.rowptr = rowptr; p.col = col; p.val = val;
  p.X = reinterpret_cast<const float4*>(X);
  p.Y = reinterpret_cast<float4*>(Y);
  p.addend = reinterpret_cast<const float4*>(addend);
  p.acc_in = reinterpret_cast<const float4*>(acc_in);
  p.acc_out = reinterpret_cast<float4*>(acc_out);
  p.acc_div = acc_div;
  p.noise = reinterpret_cast<const float4*>(noise);
  p.eps = eps;
  p.row_order = row_order; p.n_long = n_long; p.n_rows = n_rows;
  p.row_mask = row_mask; p.col_mask = col_mask;
  p.n_peers = n_peers;
  for (int q = 0; q < AGCF_MAX_PEERS; ++q) {
    p.peer_Y[q] = (q < n_peers && peer_Y_host) ? reinterpret_cast<float4*>(peer_Y_host[q]) : nullptr;
    p.peer_acc[q] = (q < n_peers && peer_acc_host) ? reinterpret_cast<float4*>(peer_acc_host[q]) : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 32: return launch_spmm<32>(p, st);
    case 64: return launch_spmm<64>(p, st);
    case 128: return launch_spmm<128>(p, st);
    case 256: return launch_spmm<256>(p, st);
  }
  return AGCF_EUNSUPPORTED;
}

extern "C" int agcf_sddmm_csr_f32(const int32_t* rowptr, const int32_t* col, const float* H, const float* E,
                                  float* gval, int32_t accumulate, const int32_t* row_order,
                                  int32_t n_rows, int32_t d, agcf_stream_t stream) {
  if (!rowptr || !col || !H || !E || !gval || n_rows < 0) return AGCF_EINVAL;
  if (!supported_d(d)) return AGCF_EUNSUPPORTED;
  if (!aligned16(H) || !aligned16(E)) return AGCF_EINVAL;
  if (n_rows == 0) return AGCF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float4* H4 = reinterpret_cast<const float4*>(H);
  const float4* E4 = reinterpret_cast<const float4*>(E);
#define AGCF_SDDMM(DD)                                                                          \
  {                                                                                             \
    const unsigned blocks = (unsigned)((n_rows + RowCfg<DD>::RPB - 1) / RowCfg<DD>::RPB);        \
    sddmm_csr_kernel<DD><<<blocks, 256, 0, st>>>(rowptr, col, H4, E4, gval, accumulate, row_order, n_rows); \
  }
  switch (d) {
    case 32: AGCF_SDDMM(32) break;
    case 64: AGCF_SDDMM(64) break;
    case 128: AGCF_SDDMM(128) break;
    case 256: AGCF_SDDMM(256) break;
  }
#undef AGCF_SDDMM
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}

static int norm_or_expand(const int32_t* rowptr, const int32_t* col, const float* w, const float* d_row,
                          const float* d_col, float* val, int32_t* row_of, int32_t n_rows, int64_t nnz,
                          cudaStream_t st) {
  if (n_rows == 0 || nnz == 0) return AGCF_OK;
  if (nnz < 0 || nnz > 0x7fffffffLL) return AGCF_EINVAL;
  const unsigned blocks = (unsigned)(((long long)nnz + 255) / 256);
  norm_adj_kernel<<<blocks, 256, 0, st>>>(rowptr, col, w, d_row, d_col, val, row_of, n_rows, nnz);
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}

extern "C" int agcf_norm_adj_csr(const int32_t* rowptr, const int32_t* col, const float* w,
                                 const float* d_row, const float* d_col, float* val,
                                 int32_t n_rows, int64_t nnz, agcf_stream_t stream) {
  if (!rowptr || !col || !w || !d_row || !d_col || !val || n_rows < 0) return AGCF_EINVAL;
  return norm_or_expand(rowptr, col, w, d_row, d_col, val, nullptr, n_rows, nnz, (cudaStream_t)stream);
}

extern "C" int agcf_csr_expand_rows(const int32_t* rowptr, int32_t* row_of, int32_t n_rows, int64_t nnz,
                                    agcf_stream_t stream) {
  if (!rowptr || !row_of || n_rows < 0) return AGCF_EINVAL;
  return norm_or_expand(rowptr, nullptr, nullptr, nullptr, nullptr, nullptr, row_of, n_rows, nnz, (cudaStream_t)stream);
}

extern "C" int agcf_concat_rows_f32(const float* a, int64_t n_a, const float* b, int64_t n_b,
                                    float* out, int32_t d, agcf_stream_t stream) {
  if (!out || n_a < 0 || n_b < 0 || (n_a > 0 && !a) || (n_b > 0 && !b) || d <= 0 || (d & 3)) return AGCF_EINVAL;
  if (!aligned16(a) || !aligned16(b) || !aligned16(out)) return AGCF_EINVAL;
  const long long na4 = n_a * (d / 4), nb4 = n_b * (d / 4);
  if (na4 + nb4 == 0) return AGCF_OK;
  long long blocks = (na4 + nb4 + 255) / 256;
  if (blocks > kSMs * 16) blocks = kSMs * 16;
  concat_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(a), na4, reinterpret_cast<const float4*>(b), nb4, reinterpret_cast<float4*>(out));
  AGCF_LAUNCH_OK();
  return AGCF_OK;
}
