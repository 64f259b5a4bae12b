// The two VQGAN-specific steps either side of the latent bridge loop:
//   * row softmax of the single-head AttnBlock's T x T score matrix, written as split-bf16 planes
//     (the A operand of the P.V tensor-core GEMM)                      model/VQGAN/model.py:140-192
//   * VectorQuantizer2 nearest-codebook lookup                          model/VQGAN/quantize.py:271-312
// Everything else of the autoencoder (ResnetBlocks, 1x1/3x3 convs, GroupNorm, resampling) runs on the
// same kernels as the UNet.
#include "common.cuh"

namespace bbdm {

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();                      // red may still be read from a previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];   // fixed order, every thread
  return r;
}

// exp(scale*s - m) with the product rounded first, as the reference scales the scores before its softmax
__device__ __forceinline__ float sexp(float s, float scale, float m) { return expf(__fsub_rn(__fmul_rn(s, scale), m)); }

// one CTA per row: p = softmax(scale * s) -> hi/lo planes.  N % 4 == 0.
__global__ void __launch_bounds__(256)
softmax_rows_split_kernel(const float* __restrict__ src, int64_t N, float scale, __nv_bfloat16* __restrict__ hi,
                          __nv_bfloat16* __restrict__ lo) {
  __shared__ float red[8];
  const float* row = src + (int64_t)blockIdx.x * N;
  float mx = -INFINITY;
  for (int64_t i = threadIdx.x * 4; i < N; i += 1024) {
    const float4 v = ld_f4(row + i);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  // scale > 0: max(scale * s) = scale * max(s) exactly (monotone rounding)
  const float m = block_reduce(mx, red, true) * scale;
  float sum = 0.f;
  for (int64_t i = threadIdx.x * 4; i < N; i += 1024) {
    const float4 v = ld_f4(row + i);
    sum += sexp(v.x, scale, m) + sexp(v.y, scale, m) + sexp(v.z, scale, m) + sexp(v.w, scale, m);
  }
  const float inv = 1.0f / block_reduce(sum, red, false);
  __nv_bfloat16* h = hi + (int64_t)blockIdx.x * N;
  __nv_bfloat16* l = lo + (int64_t)blockIdx.x * N;
  for (int64_t i = threadIdx.x * 4; i < N; i += 1024) {
    const float4 v = ld_f4(row + i);
    const float4 p = make_float4(sexp(v.x, scale, m) * inv, sexp(v.y, scale, m) * inv, sexp(v.z, scale, m) * inv,
                                 sexp(v.w, scale, m) * inv);
    uint2 ph, pl;
    split4(p, ph, pl);
    *reinterpret_cast<uint2*>(h + i) = ph;
    *reinterpret_cast<uint2*>(l + i) = pl;
  }
}

// nearest codebook entry per latent vector; d = (|z|^2 + |e|^2) - 2 z.e exactly as the reference writes it
// (fp32, first minimum wins); z_q = z + (e - z) (the straight-through expression's forward value).
constexpr int VQ_TILE = 512;
constexpr int VQ_MAXD = 16;

__global__ void __launch_bounds__(256)
vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ cb, int64_t N, int n_e, int D,
                  float* __restrict__ zq, long long* __restrict__ idx) {
  __shared__ float es[VQ_TILE * VQ_MAXD];
  __shared__ float e2[VQ_TILE];
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float zv[VQ_MAXD];
  float z2 = 0.f;
#pragma unroll
  for (int d = 0; d < VQ_MAXD; ++d) {
    zv[d] = (n < N && d < D) ? z[n * D + d] : 0.f;
    if (d < D) z2 = __fadd_rn(z2, __fmul_rn(zv[d], zv[d]));
  }
  float best = INFINITY;
  int best_i = 0;
  for (int t0 = 0; t0 < n_e; t0 += VQ_TILE) {
    const int cnt = min(VQ_TILE, n_e - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * D; i += blockDim.x) es[i] = cb[(int64_t)t0 * D + i];
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      float s = 0.f;
      for (int d = 0; d < D; ++d) s = __fadd_rn(s, __fmul_rn(es[i * D + d], es[i * D + d]));
      e2[i] = s;
    }
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      float dot = 0.f;
#pragma unroll
      for (int d = 0; d < VQ_MAXD; ++d)
        if (d < D) dot = fmaf(zv[d], es[i * D + d], dot);
      const float dist = __fsub_rn(__fadd_rn(z2, e2[i]), __fmul_rn(2.0f, dot));
      if (dist < best) { best = dist; best_i = t0 + i; }
    }
  }
  if (n < N) {
    idx[n] = best_i;
#pragma unroll
    for (int d = 0; d < VQ_MAXD; ++d)
      if (d < D) zq[n * D + d] = __fadd_rn(zv[d], __fsub_rn(cb[(int64_t)best_i * D + d], zv[d]));
  }
}

// space-to-depth by 2 + bf16 split: dst[b, i, j, (a*2+b2)*C + c] = src[b, 2i+a, 2j+b2, c]
__global__ void __launch_bounds__(256)
s2d_split_kernel(const float* __restrict__ src, int64_t n4, int H, int W, int C, __nv_bfloat16* __restrict__ hi,
                 __nv_bfloat16* __restrict__ lo) {
  const int C4 = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c = (int)(r % C4) * 4; r /= C4;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int64_t b = r / H;
    const float4 v = ld_f4(src + i * 4);
    uint2 ph, pl;
    split4(v, ph, pl);
    const int64_t o = (((b * (H / 2) + h / 2) * (W / 2) + w / 2) * 4 + (h & 1) * 2 + (w & 1)) * C + c;
    *reinterpret_cast<uint2*>(hi + o) = ph;
    *reinterpret_cast<uint2*>(lo + o) = pl;
  }
}

}  // namespace bbdm

using namespace bbdm;

// [B,H,W,C] fp32 -> split-bf16 planes [B,H/2,W/2,4C] (channel = (row parity*2 + col parity)*C + c): the A operand
// of a stride-2 3x3 convolution run as a 2x2-tap tensor-core conv (BbdmConvArgs.taps = 4).
extern "C" int bbdm_s2d_split(const float* src, int B, int H, int W, int C, void* out_hi, void* out_lo, void* stream) {
  BBDM_REQUIRE(src && out_hi && out_lo, "s2d_split: null pointer");
  BBDM_REQUIRE(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0, "s2d_split: bad shape");
  const int64_t n4 = (int64_t)B * H * W * (C / 4);
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  s2d_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, n4, H, W, C, (__nv_bfloat16*)out_hi,
                                                                       (__nv_bfloat16*)out_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

extern "C" int bbdm_softmax_rows_split(const float* src, int64_t rows, int64_t cols, float scale, void* out_hi,
                                       void* out_lo, void* stream) {
  BBDM_REQUIRE(src && out_hi && out_lo, "softmax_rows_split: null pointer");
  BBDM_REQUIRE(rows > 0 && rows < (1ll << 31) && cols > 0 && cols % 4 == 0 && scale > 0.f,
               "softmax_rows_split: bad shape (cols must be a multiple of 4, scale > 0)");
  softmax_rows_split_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(
      src, cols, scale, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

extern "C" int bbdm_vq_nearest(const float* z, const float* codebook, int64_t n_vectors, int n_embed, int dim,
                               float* z_q, long long* indices, void* stream) {
  BBDM_REQUIRE(z && codebook && z_q && indices, "vq_nearest: null pointer");
  BBDM_REQUIRE(n_vectors > 0 && n_embed > 0 && dim > 0 && dim <= VQ_MAXD, "vq_nearest: bad shape (dim <= %d)", VQ_MAXD);
  const int64_t blocks = (n_vectors + 255) / 256;
  BBDM_REQUIRE(blocks < (1ll << 31), "vq_nearest: too many vectors");
  vq_nearest_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, codebook, n_vectors, n_embed, dim, z_q, indices);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}
