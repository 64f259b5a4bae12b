// General fp32 convolution on CUDA cores (SIMT implicit GEMM, 64 pixels x TN couts per CTA).
// Serves the UNet edges (stem Cin=3..32, head Cout=3..16), conv-mode Down/Upsample and any
// channel count the tcgen05 kernel does not take.  Exact fp32 FMA accumulation.
#include "common.cuh"

namespace bbdm {

constexpr int CD_TM = 64;   // output pixels per CTA
constexpr int CD_TK = 16;   // cin chunk

template <int TN>           // couts per CTA: 64 or 16
__global__ void __launch_bounds__(256)
conv_direct_kernel(const float* __restrict__ src, const float* __restrict__ wp,
                   const float* __restrict__ bias, const float* __restrict__ res,
                   float* __restrict__ out, int B, int H, int W, int Cin, int Cout, int k,
                   int stride, int Ho, int Wo) {
  constexpr int TPN = TN / 4;          // threads along N (each 4 couts)
  constexpr int TPM = 256 / TPN;       // threads along M
  constexpr int PM = CD_TM / TPM;      // pixels per thread (TN=64: 4 ; TN=16: 1)
  __shared__ float As[CD_TK][CD_TM + 4];
  __shared__ float Ws[CD_TK][TN + 4];
  const int tid = threadIdx.x;
  const int tn = tid % TPN, tm = tid / TPN;
  const int64_t M = (int64_t)B * Ho * Wo;
  const int64_t m0 = (int64_t)blockIdx.x * CD_TM;
  const int n0 = blockIdx.y * TN;
  const int pad = k / 2;

  // A-load assignment: 64 pixels x 16 cin = 1024 elements, 4 per thread
  const int a_p = tid / 4;             // pixel within tile
  const int a_k = (tid % 4) * 4;       // first cin within chunk
  int ab = 0, aho = 0, awo = 0;
  const bool a_valid = (m0 + a_p) < M;
  if (a_valid) {
    int64_t m = m0 + a_p;
    awo = (int)(m % Wo); m /= Wo;
    aho = (int)(m % Ho);
    ab = (int)(m / Ho);
  }
  float acc[PM][4];
#pragma unroll
  for (int i = 0; i < PM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < k * k; ++tap) {
    const int dy = tap / k - pad, dx = tap % k - pad;
    const int hi = aho * stride + dy, wi = awo * stride + dx;
    const bool in_img = a_valid && hi >= 0 && hi < H && wi >= 0 && wi < W;
    const float* arow = src + (((int64_t)ab * H + hi) * W + wi) * Cin;
    for (int c0 = 0; c0 < Cin; c0 += CD_TK) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + a_k + j;
        As[a_k + j][a_p] = (in_img && c < Cin) ? arow[c] : 0.f;
      }
      // W tile: 16 x TN
      for (int i = tid; i < CD_TK * TN; i += 256) {
        const int kk = i / TN, n = i % TN;
        const int c = c0 + kk, co = n0 + n;
        Ws[kk][n] = (c < Cin && co < Cout) ? wp[((int64_t)tap * Cin + c) * Cout + co] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < CD_TK; ++kk) {
        float a[PM], w[4];
#pragma unroll
        for (int i = 0; i < PM; ++i) a[i] = As[kk][tm * PM + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tn * 4 + j];
#pragma unroll
        for (int i = 0; i < PM; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < PM; ++i) {
    const int64_t m = m0 + tm * PM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tn * 4 + j;
      if (co >= Cout) continue;
      float v = acc[i][j] + (bias ? bias[co] : 0.f);
      if (res) v += res[m * Cout + co];
      out[m * Cout + co] = v;
    }
  }
}

}  // namespace bbdm

using namespace bbdm;

extern "C" int bbdm_conv_direct(const float* src, const float* w_packed, const float* bias,
                                const float* residual, float* out, int B, int H, int W, int Cin,
                                int Cout, int k, int stride, void* stream) {
  BBDM_REQUIRE(src && w_packed && out, "conv_direct: null pointer");
  BBDM_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (k == 1 || k == 3) && (stride == 1 || stride == 2),
               "conv_direct: bad shape");
  const int pad = k / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t M = (int64_t)B * Ho * Wo;
  const int64_t gm = (M + CD_TM - 1) / CD_TM;
  BBDM_REQUIRE(gm < (1ll << 31), "conv_direct: too many pixels");
  cudaStream_t s = (cudaStream_t)stream;
  if (Cout <= 16) {
    dim3 grid((unsigned)gm, (Cout + 15) / 16);
    conv_direct_kernel<16><<<grid, 256, 0, s>>>(src, w_packed, bias, residual, out, B, H, W, Cin, Cout, k, stride, Ho, Wo);
  } else {
    dim3 grid((unsigned)gm, (Cout + 63) / 64);
    conv_direct_kernel<64><<<grid, 256, 0, s>>>(src, w_packed, bias, residual, out, B, H, W, Cin, Cout, k, stride, Ho, Wo);
  }
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}
