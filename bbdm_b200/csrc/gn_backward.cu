// Backward of the fused GroupNorm-affine (+FiLM) (+SiLU) operand preparation (training path).
//
//   forward (prep_kernel):  xh = (x - mean) * rstd ;  y = gamma * xh + beta ;  z = y * f1 + f0 ;
//                           a = silu(z)            (f1 = 1 + film_scale, f0 = film_shift, per (b, c))
//   given dA (gradient w.r.t. a, NHWC fp32):
//     dz   = dA * silu'(z)
//     A1[b,c] = sum_p dz            A2[b,c] = sum_p dz * xh          <- gn_bwd_reduce  (pass 1)
//     dshift = A1 ; dscale = gamma*A2 + beta*A1 ; dbeta = sum_b f1*A1 ; dgamma = sum_b f1*A2
//     s1[b,g] = sum_{c in g} gamma*f1*A1 ; s2[b,g] = sum_{c in g} gamma*f1*A2       (host, tiny)
//     dx = rstd * ( dz*gamma*f1 - (s1 + xh*s2) / n )                  <- gn_bwd_apply   (pass 2)
// Both passes are HBM-bound: they read x and dA (8 B/element); pass 2 writes dx (4 B/element).
// Replaces the autograd of GroupNorm32 + SiLU + scale-shift (openaimodel.py:205-206,229-230,270-274).
#include "common.cuh"

namespace bbdm {

struct GnBwdParams {
  const float* x; const float* da;
  int B, C, groups, cpg;
  int64_t HW;
  const float* mean; const float* rstd; const float* gamma; const float* beta;
  const float* fscale; const float* fshift; int64_t fstride;
  int silu;
};

__device__ __forceinline__ float dz_of(const GnBwdParams& p, float xh, float ga, float be, float f1, float f0, float da) {
  if (!p.silu) return da;
  const float z = fmaf(fmaf(ga, xh, be), f1, f0);
  const float sg = __fdividef(1.0f, 1.0f + __expf(-z));
  return da * sg * fmaf(z, 1.0f - sg, 1.0f);
}

// pass 1: grid (S, B); thread = (float4 channel chunk, pixel lane); partial [B][S][C][2]
__global__ void __launch_bounds__(256)
gn_bwd_reduce_kernel(const GnBwdParams p, int L, int R, float* __restrict__ part) {
  extern __shared__ float sm[];   // [C][2]
  const int C4 = p.C / 4;
  const int b = blockIdx.y, S = gridDim.x, s = blockIdx.x;
  const int64_t p0 = p.HW * s / S, p1 = p.HW * (s + 1) / S;
  const int col = threadIdx.x % L, row = threadIdx.x / L;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  for (int cv0 = 0; cv0 < C4; cv0 += L) {
    const int cv = cv0 + col;
    const bool active = row < R && cv < C4;
    const int c = cv * 4;
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    if (active) {
      float mu[4], rs[4], ga[4], be[4], f1[4], f0[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int g = (c + v) / p.cpg;
        mu[v] = p.mean[b * p.groups + g]; rs[v] = p.rstd[b * p.groups + g];
        ga[v] = p.gamma[c + v]; be[v] = p.beta[c + v];
        f1[v] = p.fscale ? 1.0f + p.fscale[(int64_t)b * p.fstride + c + v] : 1.0f;
        f0[v] = p.fscale ? p.fshift[(int64_t)b * p.fstride + c + v] : 0.0f;
      }
      for (int64_t px = p0 + row; px < p1; px += R) {
        const int64_t off = ((int64_t)b * p.HW + px) * p.C + c;
        const float4 xv = ld_f4(p.x + off), dv = ld_f4(p.da + off);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float xh = (xs[v] - mu[v]) * rs[v];
          const float dz = dz_of(p, xh, ga[v], be[v], f1[v], f0[v], ds[v]);
          a1[v] += dz;
          a2[v] = fmaf(dz, xh, a2[v]);
        }
      }
    }
    for (int r = 0; r < R; ++r) {       // fixed combination order over the pixel lanes
      if (active && row == r) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { sm[2 * (c + v)] += a1[v]; sm[2 * (c + v) + 1] += a2[v]; }
      }
      __syncthreads();
    }
  }
  float* o = part + ((int64_t)b * S + s) * p.C * 2;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) o[i] = sm[i];
}

__global__ void gn_bwd_reduce_final_kernel(const float* __restrict__ part, int S, int n, float* __restrict__ out, int BC2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over B*C*2
  if (i >= BC2) return;
  const int b = i / n, r = i % n;
  double s = 0.0;
  for (int k = 0; k < S; ++k) s += (double)part[((int64_t)b * S + k) * n + r];
  out[i] = (float)s;
}

// pass 2: one CTA per (b, pixel-row block); per-channel constants in smem
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const GnBwdParams p, const float* __restrict__ s1, const float* __restrict__ s2, float inv_n,
                    float* __restrict__ dx, int64_t px_per_block) {
  extern __shared__ float sm[];   // per channel: mu, rs, ga, be, f1, f0, k1 (= rs*ga*f1), t1 (= rs*s1*inv_n), t2 (= rs*s2*inv_n)
  float* cst = sm;
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int g = c / p.cpg;
    const float mu = p.mean[b * p.groups + g], rs = p.rstd[b * p.groups + g];
    const float ga = p.gamma[c], be = p.beta[c];
    const float f1 = p.fscale ? 1.0f + p.fscale[(int64_t)b * p.fstride + c] : 1.0f;
    const float f0 = p.fscale ? p.fshift[(int64_t)b * p.fstride + c] : 0.0f;
    float* q = cst + c * 9;
    q[0] = mu; q[1] = rs; q[2] = ga; q[3] = be; q[4] = f1; q[5] = f0;
    q[6] = rs * ga * f1;
    q[7] = rs * s1[b * p.groups + g] * inv_n;
    q[8] = rs * s2[b * p.groups + g] * inv_n;
  }
  __syncthreads();
  const int C4 = p.C / 4;
  const int64_t px0 = (int64_t)blockIdx.x * px_per_block;
  const int64_t n4 = px_per_block * C4;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    const int64_t px = px0 + i / C4;
    if (px >= p.HW) break;
    const int c = (int)(i % C4) * 4;
    const int64_t off = ((int64_t)b * p.HW + px) * p.C + c;
    const float4 xv = ld_f4(p.x + off), dv = ld_f4(p.da + off);
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float* q = cst + (c + v) * 9;
      const float xh = (xs[v] - q[0]) * q[1];
      const float dz = dz_of(p, xh, q[2], q[3], q[4], q[5], ds[v]);
      o[v] = fmaf(dz, q[6], -fmaf(xh, q[8], q[7]));
    }
    st_f4(dx + off, make_float4(o[0], o[1], o[2], o[3]));
  }
}

}  // namespace bbdm

using namespace bbdm;

extern "C" {

// a12[b][c][2] = (sum_p dz, sum_p dz*xh).  workspace: B*64*C*2 floats.
int bbdm_gn_bwd_reduce(const float* x, const float* da, int B, int H, int W, int C, int groups, const float* mean,
                       const float* rstd, const float* gamma, const float* beta, const float* film_scale,
                       const float* film_shift, int64_t film_stride, int silu, float* a12, float* workspace,
                       void* stream) {
  BBDM_REQUIRE(x && da && mean && rstd && gamma && beta && a12 && workspace, "gn_bwd_reduce: null pointer");
  BBDM_REQUIRE(B > 0 && B <= 65535 && C % 4 == 0 && C % groups == 0, "gn_bwd_reduce: bad shape");
  GnBwdParams p{x, da, B, C, groups, C / groups, (int64_t)H * W, mean, rstd, gamma, beta, film_scale, film_shift,
                film_stride, silu};
  const int C4 = C / 4;
  const int L = C4 < 256 ? C4 : 256;
  int R = 256 / L;
  if (R > 8) R = 8;
  if ((int64_t)R > p.HW) R = (int)p.HW;
  int S = (8 * num_sms() + B - 1) / B;
  if (S > 64) S = 64;
  if ((int64_t)S * R * 8 > p.HW) S = (int)(p.HW / ((int64_t)R * 8));
  if (S < 1) S = 1;
  const size_t smem = (size_t)C * 2 * sizeof(float);
  BBDM_REQUIRE(smem <= 48 * 1024, "gn_bwd_reduce: C too large");
  cudaStream_t s = (cudaStream_t)stream;
  gn_bwd_reduce_kernel<<<dim3(S, B), 256, smem, s>>>(p, L, R, workspace);
  BBDM_LAUNCH_CHECK();
  const int n = B * C * 2;
  gn_bwd_reduce_final_kernel<<<(n + 255) / 256, 256, 0, s>>>(workspace, S, C * 2, a12, n);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

// dx = rstd * (dz*gamma*f1 - (s1 + xh*s2)/n),  s1/s2: [B,groups],  n = H*W*C/groups
int bbdm_gn_bwd_apply(const float* x, const float* da, int B, int H, int W, int C, int groups, const float* mean,
                      const float* rstd, const float* gamma, const float* beta, const float* film_scale,
                      const float* film_shift, int64_t film_stride, int silu, const float* s1, const float* s2,
                      float* dx, void* stream) {
  BBDM_REQUIRE(x && da && mean && rstd && gamma && beta && s1 && s2 && dx, "gn_bwd_apply: null pointer");
  BBDM_REQUIRE(B > 0 && B <= 65535 && C % 4 == 0 && C % groups == 0, "gn_bwd_apply: bad shape");
  GnBwdParams p{x, da, B, C, groups, C / groups, (int64_t)H * W, mean, rstd, gamma, beta, film_scale, film_shift,
                film_stride, silu};
  const size_t smem = (size_t)C * 9 * sizeof(float);
  if (smem > 48 * 1024) BBDM_CUDA_CHECK(cudaFuncSetAttribute(gn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BBDM_REQUIRE(smem <= 200 * 1024, "gn_bwd_apply: C too large");
  // ~32K elements per CTA
  int64_t ppb = 32768 / C;
  if (ppb < 1) ppb = 1;
  const int64_t nblk = (p.HW + ppb - 1) / ppb;
  const float inv_n = 1.0f / (float)((double)p.HW * (C / groups));
  gn_bwd_apply_kernel<<<dim3((unsigned)nblk, B), 256, smem, (cudaStream_t)stream>>>(p, s1, s2, inv_n, dx, ppb);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // extern "C"
