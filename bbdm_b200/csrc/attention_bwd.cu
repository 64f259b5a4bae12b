// Backward of the multi-head softmax attention core (training path), FlashAttention-style: the
// T x T probability matrix is recomputed tile by tile, never materialised (the reference's autograd
// keeps [B*heads, T, T] fp32 -- 17 GB at T = 4096, B = 16 -- and recomputes it under checkpoint(),
// openaimodel.py:318, util.py:119-148).
//
//   forward (QKVAttentionLegacy / QKVAttention, openaimodel.py:350-413):
//     S = (q s)(k s)^T, s = D^-1/4 ;  P = softmax_row(S) ;  O = P V
//   given dO:
//     delta_i = sum_d dO_id O_id            L_i = log2 sum_j 2^(S_ij log2e)          (kernel 1)
//     P_ij = 2^(S_ij log2e - L_i) ;  dP = dO V^T ;  dS = P o (dP - delta)
//     dQ = s^2 dS K   (kernel 1, one CTA per 64-query tile, loops over the key tiles twice: L, then dQ)
//     dK = s^2 dS^T Q ;  dV = P^T dO   (kernel 2, one CTA per 64-key tile, loops over the query tiles)
//
// Exact fp32 FMA arithmetic on CUDA cores (64x64 tiles, 4x4 register micro-tiles, operands staged in
// shared memory in both orientations); deterministic (no atomics).  The attention core is <= 1.7 % of
// the UNet's FLOPs, so this kernel is about correctness and memory, not the tensor pipe.
#include "common.cuh"

namespace bbdm {

constexpr int AB_T = 64;           // tile edge (queries and keys)
constexpr int AB_LD = AB_T + 4;    // row stride of the P / dS tiles (float4-aligned, conflict-free)

struct AttnBwdParams {
  const float* qkv; const float* o; const float* dout; float* dqkv;
  float* lse; float* delta;        // [B*heads, T]
  int T, C, heads, order;
  float scale2, scale_log2;
};

__device__ __forceinline__ void head_offsets(const AttnBwdParams& p, int head, int D, int& qoff, int& koff, int& voff) {
  if (p.order == 0) { qoff = head * 3 * D; koff = qoff + D; voff = qoff + 2 * D; }
  else { qoff = head * D; koff = p.C + head * D; voff = 2 * p.C + head * D; }
}

// 64 rows x D columns of a [*, ld] fp32 matrix -> transposed tile dst_t[D][64] (and row-major dst_r[64][D])
template <int D>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int64_t ld, int t0, int T, float* dst_t, float* dst_r) {
  for (int i = threadIdx.x; i < AB_T * (D / 4); i += 256) {
    const int row = i % AB_T, ch = i / AB_T;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + row < T) v = ld_f4(src + (int64_t)(t0 + row) * ld + ch * 4);
    dst_t[(ch * 4 + 0) * AB_T + row] = v.x;
    dst_t[(ch * 4 + 1) * AB_T + row] = v.y;
    dst_t[(ch * 4 + 2) * AB_T + row] = v.z;
    dst_t[(ch * 4 + 3) * AB_T + row] = v.w;
    if (dst_r) *reinterpret_cast<float4*>(dst_r + row * D + ch * 4) = v;
  }
}

// acc[i][j] = sum_d At[d][ty*4+i] * Bt[d][tx*4+j]
template <int D>
__device__ __forceinline__ void mm_tt(const float* At, const float* Bt, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int d = 0; d < D; ++d) {
    const float4 a = *reinterpret_cast<const float4*>(At + d * AB_T + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(Bt + d * AB_T + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// kernel 1: per 64-query tile -- delta, log-sum-exp, dQ
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
attn_bwd_dq_kernel(const AttnBwdParams p) {
  constexpr int DC = D / 16;        // dQ columns per thread
  extern __shared__ __align__(16) float sm[];
  float* Qt = sm;                   // [D][64]
  float* dOt = Qt + D * AB_T;       // [D][64]
  float* Kt = dOt + D * AB_T;       // [D][64]
  float* Vt = Kt + D * AB_T;        // [D][64]
  float* Ks = Vt + D * AB_T;        // [64][D]
  float* dSs = Ks + AB_T * D;       // [64][AB_LD]
  float* delta_s = dSs + AB_T * AB_LD;   // [64]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int bh = blockIdx.y, b = bh / p.heads, head = bh % p.heads;
  const int q0 = blockIdx.x * AB_T;
  int qoff, koff, voff;
  head_offsets(p, head, D, qoff, koff, voff);
  const int64_t ld3 = 3 * (int64_t)p.C;
  const float* qkv_b = p.qkv + (int64_t)b * p.T * ld3;
  const float* do_b = p.dout + (int64_t)b * p.T * p.C + head * D;
  const float* o_b = p.o + (int64_t)b * p.T * p.C + head * D;
  const int n_tiles = (p.T + AB_T - 1) / AB_T;

  load_tile<D>(qkv_b + qoff, ld3, q0, p.T, Qt, nullptr);
  load_tile<D>(do_b, p.C, q0, p.T, dOt, nullptr);
  __syncthreads();
  {
    // delta_i = <dO_i, O_i>: 4 threads per row
    const int row = tid >> 2, part = tid & 3;
    float s = 0.f;
    if (q0 + row < p.T) {
      const float* orow = o_b + (int64_t)(q0 + row) * p.C;
      for (int d = part * (D / 4); d < (part + 1) * (D / 4); ++d) s = fmaf(dOt[d * AB_T + row], orow[d], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (part == 0) {
      delta_s[row] = s;
      if (q0 + row < p.T) p.delta[(int64_t)bh * p.T + q0 + row] = s;
    }
  }

  // ---- pass 1: row-wise log-sum-exp (base 2) of the scaled scores --------------------------------
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  for (int j = 0; j < n_tiles; ++j) {
    const int k0 = j * AB_T;
    __syncthreads();
    load_tile<D>(qkv_b + koff, ld3, k0, p.T, Kt, nullptr);
    __syncthreads();
    float s[4][4];
    mm_tt<D>(Qt, Kt, ty, tx, s);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[i][c] = (k0 + tx * 4 + c < p.T) ? s[i][c] * p.scale_log2 : -INFINITY;
        mx = fmaxf(mx, s[i][c]);
      }
      mx = group16_max(mx);
      const float mn = fmaxf(m[i], mx);           // finite: every tile holds at least one valid key
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) rs += exp2f(s[i][c] - mn);
      rs = group16_sum(rs);
      l[i] = l[i] * exp2f(m[i] - mn) + rs;
      m[i] = mn;
    }
  }
  float lse[4], dl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lse[i] = m[i] + log2f(l[i]);
    dl[i] = delta_s[ty * 4 + i];
    if (tx == 0 && q0 + ty * 4 + i < p.T) p.lse[(int64_t)bh * p.T + q0 + ty * 4 + i] = lse[i];
  }

  // ---- pass 2: dQ = s^2 * sum_j dS_j K_j -------------------------------------------------------
  float dq[4][DC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < DC; ++c) dq[i][c] = 0.f;
  for (int j = 0; j < n_tiles; ++j) {
    const int k0 = j * AB_T;
    __syncthreads();
    load_tile<D>(qkv_b + koff, ld3, k0, p.T, Kt, Ks);
    load_tile<D>(qkv_b + voff, ld3, k0, p.T, Vt, nullptr);
    __syncthreads();
    float s[4][4], dp[4][4];
    mm_tt<D>(Qt, Kt, ty, tx, s);
    mm_tt<D>(dOt, Vt, ty, tx, dp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float ds[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float pr = (k0 + tx * 4 + c < p.T) ? exp2f(fmaf(s[i][c], p.scale_log2, -lse[i])) : 0.f;
        ds[c] = pr * (dp[i][c] - dl[i]);
      }
      *reinterpret_cast<float4*>(dSs + (ty * 4 + i) * AB_LD + tx * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < AB_T; ++k) {
      float kv[DC];
#pragma unroll
      for (int c = 0; c < DC; ++c) kv[c] = Ks[k * D + tx * DC + c];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a = dSs[(ty * 4 + i) * AB_LD + k];
#pragma unroll
        for (int c = 0; c < DC; ++c) dq[i][c] = fmaf(a, kv[c], dq[i][c]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= p.T) continue;
    float* dst = p.dqkv + ((int64_t)b * p.T + q) * ld3 + qoff + tx * DC;
#pragma unroll
    for (int c = 0; c < DC; ++c) dst[c] = dq[i][c] * p.scale2;
  }
}

// ---------------------------------------------------------------------------------------------
// kernel 2: per 64-key tile -- dK, dV
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
attn_bwd_dkv_kernel(const AttnBwdParams p) {
  constexpr int DC = D / 16;
  extern __shared__ __align__(16) float sm[];
  float* Kt = sm;                   // [D][64]
  float* Vt = Kt + D * AB_T;
  float* Qt = Vt + D * AB_T;
  float* dOt = Qt + D * AB_T;
  float* Qs = dOt + D * AB_T;       // [64][D]
  float* dOs = Qs + AB_T * D;       // [64][D]
  float* Ps = dOs + AB_T * D;       // [64 q][AB_LD]
  float* dSs = Ps + AB_T * AB_LD;   // [64 q][AB_LD]
  float* lse_s = dSs + AB_T * AB_LD;
  float* delta_s = lse_s + AB_T;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int bh = blockIdx.y, b = bh / p.heads, head = bh % p.heads;
  const int k0 = blockIdx.x * AB_T;
  int qoff, koff, voff;
  head_offsets(p, head, D, qoff, koff, voff);
  const int64_t ld3 = 3 * (int64_t)p.C;
  const float* qkv_b = p.qkv + (int64_t)b * p.T * ld3;
  const float* do_b = p.dout + (int64_t)b * p.T * p.C + head * D;
  const int n_tiles = (p.T + AB_T - 1) / AB_T;

  load_tile<D>(qkv_b + koff, ld3, k0, p.T, Kt, nullptr);
  load_tile<D>(qkv_b + voff, ld3, k0, p.T, Vt, nullptr);

  float dk[4][DC], dv[4][DC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < DC; ++c) { dk[i][c] = 0.f; dv[i][c] = 0.f; }

  for (int j = 0; j < n_tiles; ++j) {
    const int q0 = j * AB_T;
    __syncthreads();
    load_tile<D>(qkv_b + qoff, ld3, q0, p.T, Qt, Qs);
    load_tile<D>(do_b, p.C, q0, p.T, dOt, dOs);
    if (tid < AB_T) {
      const bool ok = q0 + tid < p.T;
      lse_s[tid] = ok ? p.lse[(int64_t)bh * p.T + q0 + tid] : 0.f;
      delta_s[tid] = ok ? p.delta[(int64_t)bh * p.T + q0 + tid] : 0.f;
    }
    __syncthreads();
    float s[4][4], dp[4][4];
    mm_tt<D>(Qt, Kt, ty, tx, s);        // rows = queries (ty), cols = keys (tx)
    mm_tt<D>(dOt, Vt, ty, tx, dp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const bool qok = q0 + r < p.T;
      float pr[4], ds[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        pr[c] = (qok && k0 + tx * 4 + c < p.T) ? exp2f(fmaf(s[i][c], p.scale_log2, -lse_s[r])) : 0.f;
        ds[c] = pr[c] * (dp[i][c] - delta_s[r]);
      }
      *reinterpret_cast<float4*>(Ps + r * AB_LD + tx * 4) = make_float4(pr[0], pr[1], pr[2], pr[3]);
      *reinterpret_cast<float4*>(dSs + r * AB_LD + tx * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncthreads();
    // dV[k][d] += sum_q P[q][k] dO[q][d] ;  dK[k][d] += sum_q dS[q][k] Q[q][d]   (keys ty*4.., d tx*DC..)
#pragma unroll 4
    for (int q = 0; q < AB_T; ++q) {
      const float4 pa = *reinterpret_cast<const float4*>(Ps + q * AB_LD + ty * 4);
      const float4 da = *reinterpret_cast<const float4*>(dSs + q * AB_LD + ty * 4);
      const float pv[4] = {pa.x, pa.y, pa.z, pa.w}, dsv[4] = {da.x, da.y, da.z, da.w};
      float qv[DC], dov[DC];
#pragma unroll
      for (int c = 0; c < DC; ++c) { qv[c] = Qs[q * D + tx * DC + c]; dov[c] = dOs[q * D + tx * DC + c]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < DC; ++c) {
          dv[i][c] = fmaf(pv[i], dov[c], dv[i][c]);
          dk[i][c] = fmaf(dsv[i], qv[c], dk[i][c]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= p.T) continue;
    float* row = p.dqkv + ((int64_t)b * p.T + k) * ld3;
#pragma unroll
    for (int c = 0; c < DC; ++c) {
      row[koff + tx * DC + c] = dk[i][c] * p.scale2;
      row[voff + tx * DC + c] = dv[i][c];
    }
  }
}

template <int D>
static int launch_bwd(const AttnBwdParams& p, int B, cudaStream_t s) {
  const size_t sm1 = (size_t)(5 * D * AB_T + AB_T * AB_LD + AB_T) * sizeof(float);
  const size_t sm2 = (size_t)(6 * D * AB_T + 2 * AB_T * AB_LD + 2 * AB_T) * sizeof(float);
  BBDM_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  BBDM_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  const dim3 grid((p.T + AB_T - 1) / AB_T, B * p.heads);
  attn_bwd_dq_kernel<D><<<grid, 256, sm1, s>>>(p);
  BBDM_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<D><<<grid, 256, sm2, s>>>(p);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // namespace bbdm

using namespace bbdm;

// dqkv [B,T,3C] = gradient of the attention core w.r.t. its qkv input, given out = attention(qkv)
// [B,T,C] (saved from the forward) and dout [B,T,C].  lse / delta: [B*heads*T] fp32 workspaces.
extern "C" int bbdm_attention_bwd(const float* qkv, const float* out, const float* dout, int B, int T, int C, int heads,
                                  int order, float* dqkv, float* lse, float* delta, void* stream) {
  BBDM_REQUIRE(qkv && out && dout && dqkv && lse && delta, "attention_bwd: null pointer");
  BBDM_REQUIRE(B > 0 && T > 0 && heads > 0 && C % heads == 0 && (order == 0 || order == 1), "attention_bwd: bad shape");
  BBDM_REQUIRE((int64_t)B * heads <= 65535, "attention_bwd: B*heads = %lld exceeds the grid limit", (long long)B * heads);
  const int D = C / heads;
  const double scale = 1.0 / sqrt(sqrt((double)D));
  AttnBwdParams p{qkv, out, dout, dqkv, lse, delta, T, C, heads, order, (float)(scale * scale),
                  (float)(scale * scale * 1.4426950408889634)};
  cudaStream_t s = (cudaStream_t)stream;
  switch (D) {
    case 16: return launch_bwd<16>(p, B, s);
    case 32: return launch_bwd<32>(p, B, s);
    case 64: return launch_bwd<64>(p, B, s);
    default:
      BBDM_REQUIRE(false, "attention_bwd: head_dim %d not supported (16, 32, 64)", D);
  }
  return BBDM_OK;
}
