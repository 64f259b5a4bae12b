// Operand-producing passes of the SpatialTransformer (reference attention.py:36-50, 196-216): LayerNorm and GEGLU,
// each fused with the split into the bf16 (hi, lo) planes the following tcgen05 GEMM reads.  HBM-bound:
//   layernorm_split : 4 B read + 4 B written per element
//   geglu_split     : 8 B read + 4 B written per output element
#include "common.cuh"

namespace bbdm {

// One warp per token row: the row is held in registers (C <= 32 * LN_MAX_PER_LANE), two-pass mean / variance in
// fp32 like torch's LayerNorm (biased variance, eps inside the sqrt), then y = (x - mean) * rstd * gamma + beta.
constexpr int LN_MAX_PER_LANE = 64;       // C <= 2048

__global__ void __launch_bounds__(256)
layernorm_split_kernel(const float* __restrict__ x, int64_t rows, int C, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                       __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * C;
  const int n2 = C >> 1;                      // float2 elements per row
  float2 v[LN_MAX_PER_LANE / 2];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE / 2; ++i) {
    const int j = lane + 32 * i;
    v[i] = make_float2(0.f, 0.f);
    if (j < n2) { v[i] = *reinterpret_cast<const float2*>(xr + 2 * j); s += v[i].x + v[i].y; }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE / 2; ++i) {
    const int j = lane + 32 * i;
    if (j < n2) { const float a = v[i].x - mean, b = v[i].y - mean; q = fmaf(a, a, fmaf(b, b, q)); }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE / 2; ++i) {
    const int j = lane + 32 * i;
    if (j < n2) {
      const float2 g = *reinterpret_cast<const float2*>(gamma + 2 * j), b = *reinterpret_cast<const float2*>(beta + 2 * j);
      const float y0 = fmaf((v[i].x - mean) * rstd, g.x, b.x), y1 = fmaf((v[i].y - mean) * rstd, g.y, b.y);
      const int64_t off = row * C + 2 * j;
      if (out_f32) *reinterpret_cast<float2*>(out_f32 + off) = make_float2(y0, y1);
      if (out_hi) {
        uint32_t h, l;
        split2x(y0, y1, h, l);
        *reinterpret_cast<uint32_t*>(out_hi + off) = h;
        *reinterpret_cast<uint32_t*>(out_lo + off) = l;
      }
    }
  }
}

// out[r][n] = u[r][n] * gelu(u[r][N + n]), exact (erf) GELU like F.gelu's default; u = the GEGLU projection [rows][2N]
__global__ void __launch_bounds__(256)
geglu_split_kernel(const float* __restrict__ u, int64_t rows, int N, float* __restrict__ out_f32,
                   __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  const int64_t n4 = rows * (N / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (N / 4);
    const int c = (int)(i - r * (N / 4)) * 4;
    const float4 a = ld_f4(u + r * 2 * N + c), g = ld_f4(u + r * 2 * N + N + c);
    float4 o;
    o.x = a.x * (0.5f * g.x * (1.0f + erff(g.x * 0.70710678118654752440f)));
    o.y = a.y * (0.5f * g.y * (1.0f + erff(g.y * 0.70710678118654752440f)));
    o.z = a.z * (0.5f * g.z * (1.0f + erff(g.z * 0.70710678118654752440f)));
    o.w = a.w * (0.5f * g.w * (1.0f + erff(g.w * 0.70710678118654752440f)));
    const int64_t off = r * N + c;
    if (out_f32) st_f4(out_f32 + off, o);
    if (out_hi) {
      uint2 h, l;
      split4(o, h, l);
      *reinterpret_cast<uint2*>(out_hi + off) = h;
      *reinterpret_cast<uint2*>(out_lo + off) = l;
    }
  }
}

}  // namespace bbdm

using namespace bbdm;

extern "C" {

int bbdm_layernorm_split(const float* x, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                         float* out_f32, void* out_hi, void* out_lo, void* stream) {
  BBDM_REQUIRE(x && gamma && beta && rows > 0 && C > 0, "layernorm_split: bad args");
  BBDM_REQUIRE(C % 2 == 0 && C <= 32 * LN_MAX_PER_LANE, "layernorm_split: C must be even and <= %d (got %d)", 32 * LN_MAX_PER_LANE, C);
  BBDM_REQUIRE(out_f32 || (out_hi && out_lo), "layernorm_split: no output");
  BBDM_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "layernorm_split: hi/lo must come in pairs");
  const int64_t blocks = (rows + 7) / 8;
  BBDM_REQUIRE(blocks < (1ll << 31), "layernorm_split: too many rows");
  layernorm_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, C, gamma, beta, eps, out_f32,
                                                                           (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_geglu_split(const float* u, int64_t rows, int N, float* out_f32, void* out_hi, void* out_lo, void* stream) {
  BBDM_REQUIRE(u && rows > 0 && N > 0 && N % 4 == 0, "geglu_split: bad args (N %% 4 == 0 required)");
  BBDM_REQUIRE(out_f32 || (out_hi && out_lo), "geglu_split: no output");
  BBDM_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "geglu_split: hi/lo must come in pairs");
  const int64_t n4 = rows * (N / 4);
  int64_t g = (n4 + 255) / 256;
  if (g > (int64_t)num_sms() * 16) g = (int64_t)num_sms() * 16;
  geglu_split_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(u, rows, N, out_f32, (__nv_bfloat16*)out_hi,
                                                                   (__nv_bfloat16*)out_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // extern "C"
