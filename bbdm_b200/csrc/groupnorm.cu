// GroupNorm statistics + the fused operand-preparation pass (both HBM-bound).
//
// gn_stats   : 4 B/element read, fp64 accumulation, fixed reduction order (deterministic).
// prep       : reads cat(src1,src2) once, writes GN-affine(+FiLM)(+SiLU)(+up/down) and/or the raw
//              resampled tensor, as fp32 and/or split-bf16 (tensor-core operand planes).
//              Algorithmic bytes: 4 B read + 4 B written (hi+lo) per output element and result.
#include "common.cuh"

namespace bbdm {

// ------------------------------------------------------------------------------------------
// Statistics, stage 1: block (slice s, sample b) -> per-group (sum, sumsq) in fp64.
// Thread layout: L column-owners x R pixel rows; each thread keeps VEC channel accumulators.
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ s1, int c1, const float* __restrict__ s2, int c2,
                  int64_t HW, int groups, int L, int R, double* __restrict__ ws) {
  extern __shared__ double sm[];  // [C][2] per-channel sums of this block
  const int C = c1 + c2;
  const int CV = C / VEC;
  const int b = blockIdx.y, S = gridDim.x, s = blockIdx.x;
  const int64_t p0 = HW * s / S, p1 = HW * (s + 1) / S;
  const int col = threadIdx.x % L, row = threadIdx.x / L;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.0;
  __syncthreads();
  for (int cv0 = 0; cv0 < CV; cv0 += L) {   // uniform trip count: barriers inside
    const int cv = cv0 + col;
    const bool active = (row < R) && (cv < CV);
    const int c = cv * VEC;
    double a[VEC], q[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) a[v] = q[v] = 0.0;
    if (active) {
      const float* base;
      int cs, cc;
      if (c < c1) { base = s1; cs = c1; cc = c; } else { base = s2; cs = c2; cc = c - c1; }
      // 4 independent loads in flight per thread, accumulated in a fixed order (deterministic)
      const float* ptr0 = base + (int64_t)b * HW * cs + cc;
      int64_t p = p0 + row;
      for (; p + 3 * (int64_t)R < p1; p += 4 * (int64_t)R) {
        float x[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float* ptr = ptr0 + (p + (int64_t)u * R) * cs;
          if (VEC == 4) {
            const float4 t = ld_f4(ptr);
            x[u][0] = t.x; x[u][1 % VEC] = t.y; x[u][2 % VEC] = t.z; x[u][3 % VEC] = t.w;
          } else {
            x[u][0] = *ptr;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < VEC; ++v) { a[v] += (double)x[u][v]; q[v] += (double)x[u][v] * (double)x[u][v]; }
      }
      for (; p < p1; p += R) {
        const float* ptr = ptr0 + p * cs;
        float x[VEC];
        if (VEC == 4) {
          const float4 t = ld_f4(ptr);
          x[0] = t.x; x[1 % VEC] = t.y; x[2 % VEC] = t.z; x[3 % VEC] = t.w;
        } else {
          x[0] = *ptr;
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) { a[v] += (double)x[v]; q[v] += (double)x[v] * (double)x[v]; }
      }
    }
    // combine the R pixel-rows of each column in a fixed order (deterministic)
    for (int r = 0; r < R; ++r) {
      if (active && row == r) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { sm[2 * (c + v)] += a[v]; sm[2 * (c + v) + 1] += q[v]; }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    double a = 0.0, q = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += sm[2 * c]; q += sm[2 * c + 1]; }
    double* o = ws + (((int64_t)b * groups + g) * S + s) * 2;
    o[0] = a;
    o[1] = q;
  }
}

__global__ void gn_finalize_kernel(const double* __restrict__ ws, int S, int groups, double count,
                                   float eps, float* __restrict__ mean, float* __restrict__ rstd, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (b, g)
  if (i >= n) return;
  double a = 0.0, q = 0.0;
  for (int s = 0; s < S; ++s) { a += ws[((int64_t)i * S + s) * 2]; q += ws[((int64_t)i * S + s) * 2 + 1]; }
  const double m = a / count;
  double var = q / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// Statistics from the partial sums fused into the conv epilogues: one CTA per (b, group);
// thread i owns (row, channel-in-group) pairs i, i+256, ... in a fixed order, then a fixed
// shared-memory tree in fp64 => deterministic.
__global__ void __launch_bounds__(256)
gn_from_partials_kernel(const float* __restrict__ p1, int c1, int rows1, const float* __restrict__ p2, int c2,
                        int rows2, int groups, double count, float eps, float* __restrict__ mean,
                        float* __restrict__ rstd) {
  __shared__ double sa[256], sq[256];
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = (c1 + c2) / groups;
  double a = 0.0, q = 0.0;
  for (int cc = 0; cc < cpg; ++cc) {
    const int c = g * cpg + cc;
    const float* base;
    int cs, ci, rows;
    if (c < c1) { base = p1; cs = c1; ci = c; rows = rows1; } else { base = p2; cs = c2; ci = c - c1; rows = rows2; }
    for (int r = threadIdx.x; r < rows; r += 256) {
      const float2 v = *reinterpret_cast<const float2*>(base + (((int64_t)b * rows + r) * cs + ci) * 2);
      a += (double)v.x;
      q += (double)v.y;
    }
  }
  sa[threadIdx.x] = a;
  sq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double m = sa[0] / count;
    double var = sq[0] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[blockIdx.x] = (float)m;
    rstd[blockIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// ------------------------------------------------------------------------------------------
// Operand preparation.
// ------------------------------------------------------------------------------------------
struct PrepParams {
  const float* src1; int c1;
  const float* src2; int c2;
  int B, Hs, Ws, H, W, C, groups, cpg;
  const float* mean; const float* rstd; const float* gamma; const float* beta;
  const float* fscale; const float* fshift; int64_t fstride;
  int silu, resample;
  float* act_f32; __nv_bfloat16* act_hi; __nv_bfloat16* act_lo;
  float* raw_f32; __nv_bfloat16* raw_hi; __nv_bfloat16* raw_lo;
};

template <int VEC>
__device__ __forceinline__ void load_vec(const PrepParams& p, int b, int hs, int ws, int c, float* x) {
  const float* base;
  int cs, cc;
  if (c < p.c1) { base = p.src1; cs = p.c1; cc = c; } else { base = p.src2; cs = p.c2; cc = c - p.c1; }
  const float* ptr = base + (((int64_t)b * p.Hs + hs) * p.Ws + ws) * cs + cc;
  if (VEC == 4) {
    const float4 t = ld_f4(ptr);
    x[0] = t.x; x[1 % VEC] = t.y; x[2 % VEC] = t.z; x[3 % VEC] = t.w;
  } else {
    x[0] = *ptr;
  }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* f32, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t off,
                                          const float* v) {
  if (VEC == 4) {
    const float4 t = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    if (f32) st_f4(f32 + off, t);
    if (hi) {
      uint2 h, l;
      split4(t, h, l);
      *reinterpret_cast<uint2*>(hi + off) = h;
      *reinterpret_cast<uint2*>(lo + off) = l;
    }
  } else {
    if (f32) f32[off] = v[0];
    if (hi) {
      __nv_bfloat16 h, l;
      split_bf16(v[0], h, l);
      hi[off] = h;
      lo[off] = l;
    }
  }
}

// One CTA per output image row (b, h).  The per-channel affine of this sample
//   y = x * sc + sh,  sc = rstd*gamma*(1+film_scale),  sh = (beta - mean*rstd*gamma)*(1+film_scale) + film_shift
// is folded once per CTA into shared memory; the element loop is then 1 FMA + SiLU per value,
// free of integer division (thread -> (pixel lane, channel chunk) is fixed).
template <int VEC>
__global__ void __launch_bounds__(256)
prep_kernel(const PrepParams p, int TPP, int NCH, int R) {
  extern __shared__ float aff[];            // [C] scale, [C] shift
  float* sc = aff;
  float* sh = aff + p.C;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const bool want_act = p.mean != nullptr;
  const bool want_raw = p.raw_f32 || p.raw_hi;
  if (want_act) {
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      const int g = c / p.cpg;
      const float s0 = p.rstd[b * p.groups + g] * p.gamma[c];
      const float h0 = p.beta[c] - p.mean[b * p.groups + g] * s0;
      float f1 = 1.0f, f0 = 0.0f;
      if (p.fscale) { f1 = 1.0f + p.fscale[(int64_t)b * p.fstride + c]; f0 = p.fshift[(int64_t)b * p.fstride + c]; }
      sc[c] = s0 * f1;
      sh[c] = fmaf(h0, f1, f0);
    }
    __syncthreads();
  }
  const int lane_c = threadIdx.x % TPP, pl = threadIdx.x / TPP;
  if (pl >= R) return;
  int n_src = 1, hs0 = h;
  if (p.resample == BBDM_RESAMPLE_UP2) hs0 = h >> 1;
  else if (p.resample == BBDM_RESAMPLE_DOWN2) { hs0 = h * 2; n_src = 4; }
  // U pixels per iteration: U independent 16-byte loads in flight per thread before any math
  constexpr int U = 4;
  for (int w0 = pl; w0 < p.W; w0 += U * R) {
    for (int j = 0; j < NCH; ++j) {
      const int c = (lane_c + j * TPP) * VEC;
      float a[VEC], s[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) { a[v] = want_act ? sc[c + v] : 0.f; s[v] = want_act ? sh[c + v] : 0.f; }
      if (n_src == 1) {
        float x[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * R;
          if (w < p.W) load_vec<VEC>(p, b, hs0, (p.resample == BBDM_RESAMPLE_UP2) ? (w >> 1) : w, c, x[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * R;
          if (w >= p.W) continue;
          float act[VEC];
          if (want_act) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              float y = fmaf(x[u][v], a[v], s[v]);
              if (p.silu) y = __fdividef(y, 1.0f + __expf(-y));
              act[v] = y;
            }
          }
          const int64_t off = (((int64_t)b * p.H + h) * p.W + w) * p.C + c;
          if (want_act) store_vec<VEC>(p.act_f32, p.act_hi, p.act_lo, off, act);
          if (want_raw) store_vec<VEC>(p.raw_f32, p.raw_hi, p.raw_lo, off, x[u]);
        }
      } else {
        // 2x2 average pooling of the activated (and of the raw) tensor
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * R;
          if (w >= p.W) continue;
          float x[4][VEC], act[VEC], raw[VEC];
#pragma unroll
          for (int k = 0; k < 4; ++k) load_vec<VEC>(p, b, hs0 + (k >> 1), w * 2 + (k & 1), c, x[k]);
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            act[v] = raw[v] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              raw[v] += x[k][v];
              if (want_act) {
                float y = fmaf(x[k][v], a[v], s[v]);
                if (p.silu) y = __fdividef(y, 1.0f + __expf(-y));
                act[v] += y;
              }
            }
            act[v] *= 0.25f;
            raw[v] *= 0.25f;
          }
          const int64_t off = (((int64_t)b * p.H + h) * p.W + w) * p.C + c;
          if (want_act) store_vec<VEC>(p.act_f32, p.act_hi, p.act_lo, off, act);
          if (want_raw) store_vec<VEC>(p.raw_f32, p.raw_hi, p.raw_lo, off, raw);
        }
      }
    }
  }
}

}  // namespace bbdm

using namespace bbdm;

extern "C" {

int bbdm_gn_stats(const float* src1, int c1, const float* src2, int c2, int B, int H, int W,
                  int groups, float eps, float* mean, float* rstd, double* workspace, void* stream) {
  BBDM_REQUIRE(src1 && mean && rstd && workspace, "gn_stats: null pointer");
  if (!src2) c2 = 0;
  const int C = c1 + c2;
  BBDM_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && groups > 0 && C % groups == 0,
               "gn_stats: bad shape B=%d C=%d groups=%d", B, C, groups);
  const int64_t HW = (int64_t)H * W;
  const bool vec4 = (c1 % 4 == 0) && (c2 % 4 == 0);
  const int VEC = vec4 ? 4 : 1;
  const int CV = C / VEC;
  const int L = CV < 256 ? CV : 256;
  int R = 256 / L;
  if (R > 8) R = 8;
  if ((int64_t)R > HW) R = (int)HW;
  int S = (8 * num_sms() + B - 1) / B;
  if (S > BBDM_GN_MAX_SLICES) S = BBDM_GN_MAX_SLICES;
  if ((int64_t)S * R * 16 > HW) S = (int)(HW / ((int64_t)R * 16));
  if (S < 1) S = 1;
  const size_t smem = (size_t)C * 2 * sizeof(double);
  BBDM_REQUIRE(smem <= 160 * 1024, "gn_stats: C=%d too large", C);
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(S, B);
  if (vec4) {
    if (smem > 48 * 1024)
      BBDM_CUDA_CHECK(cudaFuncSetAttribute(gn_partial_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gn_partial_kernel<4><<<grid, 256, smem, s>>>(src1, c1, src2, c2, HW, groups, L, R, workspace);
  } else {
    if (smem > 48 * 1024)
      BBDM_CUDA_CHECK(cudaFuncSetAttribute(gn_partial_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gn_partial_kernel<1><<<grid, 256, smem, s>>>(src1, c1, src2, c2, HW, groups, L, R, workspace);
  }
  BBDM_LAUNCH_CHECK();
  const int n = B * groups;
  gn_finalize_kernel<<<(n + 127) / 128, 128, 0, s>>>(workspace, S, groups, (double)HW * (C / groups), eps, mean, rstd, n);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_gn_finalize_partials(const float* part1, int c1, int rows1, const float* part2, int c2, int rows2,
                              int B, int hw, int groups, float eps, float* mean, float* rstd, void* stream) {
  BBDM_REQUIRE(part1 && mean && rstd && c1 > 0 && rows1 > 0 && B > 0 && hw > 0 && groups > 0, "gn_finalize_partials: bad args");
  if (!part2) { c2 = 0; rows2 = 0; }
  BBDM_REQUIRE((c1 + c2) % groups == 0, "gn_finalize_partials: C %% groups != 0");
  gn_from_partials_kernel<<<B * groups, 256, 0, (cudaStream_t)stream>>>(
      part1, c1, rows1, part2, c2, rows2, groups, (double)hw * ((c1 + c2) / groups), eps, mean, rstd);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_prep_operand(const BbdmPrepArgs* a, void* stream) {
  BBDM_REQUIRE(a && a->src1, "prep_operand: null args");
  PrepParams p;
  p.src1 = a->src1; p.c1 = a->c1;
  p.src2 = a->src2; p.c2 = a->src2 ? a->c2 : 0;
  p.B = a->B; p.Hs = a->Hs; p.Ws = a->Ws;
  p.C = p.c1 + p.c2;
  p.groups = a->groups > 0 ? a->groups : 1;
  BBDM_REQUIRE(p.B > 0 && p.Hs > 0 && p.Ws > 0 && p.C > 0, "prep_operand: bad shape");
  p.resample = a->resample;
  if (a->resample == BBDM_RESAMPLE_UP2) { p.H = p.Hs * 2; p.W = p.Ws * 2; }
  else if (a->resample == BBDM_RESAMPLE_DOWN2) {
    BBDM_REQUIRE(p.Hs % 2 == 0 && p.Ws % 2 == 0, "prep_operand: odd size for 2x2 pooling");
    p.H = p.Hs / 2; p.W = p.Ws / 2;
  } else { BBDM_REQUIRE(a->resample == BBDM_RESAMPLE_NONE, "prep_operand: bad resample"); p.H = p.Hs; p.W = p.Ws; }
  p.mean = a->mean; p.rstd = a->rstd; p.gamma = a->gamma; p.beta = a->beta;
  if (p.mean) {
    BBDM_REQUIRE(p.rstd && p.gamma && p.beta && p.C % p.groups == 0, "prep_operand: incomplete GroupNorm args");
    BBDM_REQUIRE(a->act_f32 || (a->act_hi && a->act_lo), "prep_operand: no act output");
  }
  p.cpg = p.C / p.groups;
  p.fscale = a->film_scale; p.fshift = a->film_shift; p.fstride = a->film_stride;
  BBDM_REQUIRE((p.fscale == nullptr) == (p.fshift == nullptr), "prep_operand: film scale/shift mismatch");
  p.silu = a->silu;
  p.act_f32 = a->act_f32; p.act_hi = (__nv_bfloat16*)a->act_hi; p.act_lo = (__nv_bfloat16*)a->act_lo;
  p.raw_f32 = a->raw_f32; p.raw_hi = (__nv_bfloat16*)a->raw_hi; p.raw_lo = (__nv_bfloat16*)a->raw_lo;
  BBDM_REQUIRE((p.act_hi == nullptr) == (p.act_lo == nullptr) && (p.raw_hi == nullptr) == (p.raw_lo == nullptr),
               "prep_operand: hi/lo planes must come in pairs");
  BBDM_REQUIRE(p.mean || p.raw_f32 || p.raw_hi, "prep_operand: nothing to do");
  const bool vec4 = (p.c1 % 4 == 0) && (p.c2 % 4 == 0);
  const int CV = p.C / (vec4 ? 4 : 1);
  // threads per pixel: the largest divisor of CV that is <= 128; each thread owns NCH chunks
  int TPP = 1;
  for (int d = 1; d <= 128 && d <= CV; ++d) if (CV % d == 0) TPP = d;
  const int NCH = CV / TPP;
  int R = 256 / TPP;
  if (R > p.W) R = p.W;
  const int64_t rows = (int64_t)p.B * p.H;
  BBDM_REQUIRE(rows < (1ll << 31), "prep_operand: too many rows");
  const size_t smem = (size_t)p.C * 2 * sizeof(float);
  BBDM_REQUIRE(smem <= 48 * 1024, "prep_operand: C=%d too large", p.C);
  if (vec4) prep_kernel<4><<<(unsigned)rows, 256, smem, (cudaStream_t)stream>>>(p, TPP, NCH, R);
  else prep_kernel<1><<<(unsigned)rows, 256, smem, (cudaStream_t)stream>>>(p, TPP, NCH, R);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // extern "C"
