// HBM-bound elementwise / layout / small dense kernels of the BBDM hot path.
//   bridge q_sample / p_sample  (BrownianBridgeModel.py:128-160,171-201)  20-24 B/element
//   NCHW<->NHWC edge transforms, table gather, small fp32 linear, weight repacking
#include "common.cuh"

namespace bbdm {

// ------------------------------------------------------------------------------------------
// q_sample.  One fp32 rounding per reference torch op (no FMA contraction) => bit-exact.
//   x_t = ((1-m) * x0 + m * y) + s * noise ;  obj(grad) = m * (y - x0) + s * noise
// ------------------------------------------------------------------------------------------
template <int OBJ>
__global__ void __launch_bounds__(256)
q_sample_kernel(const float4* __restrict__ x0, const float4* __restrict__ y,
                const float4* __restrict__ nz, const int64_t* __restrict__ t,
                const float* __restrict__ m_tab, const float* __restrict__ v_tab,
                float4* __restrict__ xt, float4* __restrict__ obj, int64_t n4_per_sample, int T,
                unsigned long long* __restrict__ fault) {
  const int b = blockIdx.y;
  int64_t tb = t[b];
  if (tb < 0 || tb >= T) {
    // the reference's gather raises here (model/utils.py:6); kernels cannot raise: set the device fault word
    // (bbdm_check_device_fault reports it) and read a valid row instead of out-of-bounds memory
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(fault, 0xB0000000ull | (unsigned)(tb & 0xFFFFFF));
    tb = tb < 0 ? 0 : T - 1;
  }
  const float m = m_tab[tb];
  const float s = __fsqrt_rn(v_tab[tb]);
  const float om = __fsub_rn(1.0f, m);
  const int64_t base = (int64_t)b * n4_per_sample;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4_per_sample;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = x0[base + i], c = y[base + i], e = nz[base + i];
    float4 o, q;
#define BBDM_Q1(f)                                                                      \
  {                                                                                     \
    const float sn = __fmul_rn(s, e.f);                                                 \
    o.f = __fadd_rn(__fadd_rn(__fmul_rn(om, a.f), __fmul_rn(m, c.f)), sn);              \
    if (OBJ == BBDM_OBJ_GRAD) q.f = __fadd_rn(__fmul_rn(m, __fsub_rn(c.f, a.f)), sn);   \
    else if (OBJ == BBDM_OBJ_NOISE) q.f = e.f;                                          \
    else q.f = __fsub_rn(c.f, a.f);                                                     \
  }
    BBDM_Q1(x) BBDM_Q1(y) BBDM_Q1(z) BBDM_Q1(w)
#undef BBDM_Q1
    xt[base + i] = o;
    obj[base + i] = q;
  }
}

// ------------------------------------------------------------------------------------------
// p_sample update (scalars identical across the batch in sampling).
// ------------------------------------------------------------------------------------------
template <int OBJ, bool CLIP, bool LAST>
__global__ void __launch_bounds__(256)
p_sample_kernel(const float4* __restrict__ xt, const float4* __restrict__ y,
                const float4* __restrict__ eps, const float4* __restrict__ nz,
                BbdmPSampleCoef c, const BbdmPSampleCoef* __restrict__ c_dev, float4* __restrict__ out,
                float4* __restrict__ x0o, int64_t n4) {
  if (c_dev) c = *c_dev;   // per-step scalars from device memory (CUDA-graph replay of the loop)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 X = xt[i], Y = y[i], E = eps[i];
    float4 N = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!LAST) N = nz[i];
    float4 O, R;
#define BBDM_P1(f)                                                                             \
  {                                                                                            \
    float x0;                                                                                  \
    if (OBJ == BBDM_OBJ_GRAD) x0 = __fsub_rn(X.f, E.f);                                        \
    else if (OBJ == BBDM_OBJ_NOISE)                                                            \
      x0 = __fdiv_rn(__fsub_rn(__fsub_rn(X.f, __fmul_rn(c.m_t, Y.f)),                          \
                               __fmul_rn(c.sqrt_var_t, E.f)), c.one_minus_m_t);                \
    else x0 = __fsub_rn(Y.f, E.f);                                                             \
    if (CLIP) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);                                              \
    R.f = x0;                                                                                  \
    if (LAST) O.f = x0;                                                                        \
    else {                                                                                     \
      const float a = __fadd_rn(__fmul_rn(c.one_minus_m_nt, x0), __fmul_rn(c.m_nt, Y.f));      \
      const float d = __fsub_rn(__fsub_rn(X.f, __fmul_rn(c.one_minus_m_t, x0)),                \
                                __fmul_rn(c.m_t, Y.f));                                        \
      const float mean = __fadd_rn(a, __fmul_rn(c.c_xt, d));                                   \
      O.f = __fadd_rn(mean, __fmul_rn(c.sigma_t, N.f));                                        \
    }                                                                                          \
  }
    BBDM_P1(x) BBDM_P1(y) BBDM_P1(z) BBDM_P1(w)
#undef BBDM_P1
    out[i] = O;
    if (x0o) x0o[i] = R;
  }
}

// ------------------------------------------------------------------------------------------
// NCHW (+NCHW ctx) -> NHWC concat.  C is tiny (3..32): one thread per pixel, coalesced reads
// per channel plane, contiguous C-vector write per pixel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_nhwc_cat_kernel(const float* __restrict__ x, int c1, const float* __restrict__ ctx, int c2,
                        int64_t HW, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int C = c1 + c2;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW;
       p += (int64_t)gridDim.x * blockDim.x) {
    float* o = out + ((int64_t)b * HW + p) * C;
    for (int c = 0; c < c1; ++c) o[c] = x[((int64_t)b * c1 + c) * HW + p];
    for (int c = 0; c < c2; ++c) o[c1 + c] = ctx[((int64_t)b * c2 + c) * HW + p];
  }
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float* __restrict__ src, int C, int64_t HW, float* __restrict__ out) {
  const int b = blockIdx.y;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW;
       p += (int64_t)gridDim.x * blockDim.x) {
    const float* s = src + ((int64_t)b * HW + p) * C;
    for (int c = 0; c < C; ++c) out[((int64_t)b * C + c) * HW + p] = s[c];
  }
}

__global__ void gather_rows_kernel(const float* __restrict__ table, int rows, int width,
                                   const int64_t* __restrict__ idx, float* __restrict__ out,
                                   unsigned long long* __restrict__ fault) {
  const int b = blockIdx.x;
  int64_t r = idx[b];
  if (r < 0 || r >= rows) {   // an index error in the reference; here: device fault word + a valid row
    if (threadIdx.x == 0) atomicExch(fault, 0xB1000000ull | (unsigned)(r & 0xFFFFFF));
    r = r < 0 ? 0 : rows - 1;
  }
  for (int i = threadIdx.x; i < width; i += blockDim.x) out[(int64_t)b * width + i] = table[r * width + i];
}

// ------------------------------------------------------------------------------------------
// Small fp32 linear: one warp per output column n, all B rows (B <= 64 per pass) at once so
// the weight row is read exactly once (weight-bandwidth bound, 4 B/param).
// ------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ bias, float* __restrict__ out, int B, int K, int N,
                  int act_in, int act_out, int b0) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* wr = w + (int64_t)warp * K;
  float acc[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) acc[b] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = wr[k];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      if (b0 + b < B) {
        float xv = x[(int64_t)(b0 + b) * K + k];
        if (act_in) xv = silu_f(xv);
        acc[b] = fmaf(xv, wv, acc[b]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    const float s = warp_sum(acc[b]);
    if (lane == 0 && b0 + b < B) {
      float v = s + (bias ? bias[warp] : 0.f);
      if (act_out) v = silu_f(v);
      out[(int64_t)(b0 + b) * N + warp] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Weight repacking: OIHW fp32 -> [tap][Cout][Cin] split bf16 / [tap][Cin][Cout] fp32
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weight_split_kernel(const float* __restrict__ w, int Cout, int Cin, int kk, int Cout_pad,
                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int64_t n = (int64_t)kk * Cout * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int tap = (int)(i / ((int64_t)Cin * Cout));
    const float v = w[((int64_t)co * Cin + ci) * kk + tap];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    const int64_t o = ((int64_t)tap * Cout_pad + co) * Cin + ci;
    hi[o] = h;
    lo[o] = l;
  }
}

// data-gradient layout straight from OIHW: dst[t'][ci][co] = w[co][ci][kk-1-t']  (kernel flipped,
// Cin/Cout swapped) -- the conv that maps dY to dX is bbdm_conv_umma with these planes
__global__ void __launch_bounds__(256)
pack_weight_split_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int kk,
                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int64_t n = (int64_t)kk * Cout * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int tap = (int)(i / ((int64_t)Cin * Cout));
    const float v = w[((int64_t)co * Cin + ci) * kk + (kk - 1 - tap)];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// forward AND data-gradient layouts in one pass over OIHW (training re-packs every weight each step):
//   fwd[tap][co][ci] = w[co][ci][tap]          dgrad[t'][ci][co] = w[co][ci][kk-1-t']
// 32x32 (co, ci) tile staged in shared memory: coalesced reads of the contiguous [ci][tap] runs, coalesced
// bf16x2 writes along ci (fwd) and along co (dgrad).
template <int KK>
__global__ void __launch_bounds__(256)
pack_weight_split_both_kernel(const float* __restrict__ w, int Cout, int Cin,
                              __nv_bfloat16* __restrict__ f_hi, __nv_bfloat16* __restrict__ f_lo,
                              __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo) {
  constexpr int LD = 32 * KK + 1;
  __shared__ float t[32 * LD];
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int nci = min(32, Cin - ci0), nco = min(32, Cout - co0);
  for (int e = threadIdx.x; e < 32 * 32 * KK; e += 256) {
    const int r = e / (32 * KK), c = e % (32 * KK);
    if (r < nco && c < nci * KK) t[r * LD + c] = w[((int64_t)(co0 + r) * Cin + ci0) * KK + c];
  }
  __syncthreads();
  if (f_hi) {
    for (int e = threadIdx.x; e < KK * 32 * 16; e += 256) {          // pairs along ci
      const int cp = e % 16, r = (e / 16) % 32, tap = e / (16 * 32);
      const int ci = cp * 2;
      if (r >= nco || ci >= nci) continue;
      const float a = t[r * LD + ci * KK + tap], b = (ci + 1 < nci) ? t[r * LD + (ci + 1) * KK + tap] : 0.f;
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(a, ah, al);
      split_bf16(b, bh, bl);
      const int64_t o = ((int64_t)tap * Cout + co0 + r) * Cin + ci0 + ci;
      if (ci + 1 < nci && (o & 1) == 0) {
        *reinterpret_cast<__nv_bfloat162*>(f_hi + o) = __nv_bfloat162(ah, bh);
        *reinterpret_cast<__nv_bfloat162*>(f_lo + o) = __nv_bfloat162(al, bl);
      } else {
        f_hi[o] = ah; f_lo[o] = al;
        if (ci + 1 < nci) { f_hi[o + 1] = bh; f_lo[o + 1] = bl; }
      }
    }
  }
  if (d_hi) {
    for (int e = threadIdx.x; e < KK * 32 * 16; e += 256) {          // pairs along co
      const int rp = e % 16, c = (e / 16) % 32, tap = e / (16 * 32);
      const int r = rp * 2;
      if (c >= nci || r >= nco) continue;
      const float a = t[r * LD + c * KK + tap], b = (r + 1 < nco) ? t[(r + 1) * LD + c * KK + tap] : 0.f;
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(a, ah, al);
      split_bf16(b, bh, bl);
      const int64_t o = ((int64_t)(KK - 1 - tap) * Cin + ci0 + c) * Cout + co0 + r;
      if (r + 1 < nco && (o & 1) == 0) {
        *reinterpret_cast<__nv_bfloat162*>(d_hi + o) = __nv_bfloat162(ah, bh);
        *reinterpret_cast<__nv_bfloat162*>(d_lo + o) = __nv_bfloat162(al, bl);
      } else {
        d_hi[o] = ah; d_lo[o] = al;
        if (r + 1 < nco) { d_hi[o + 1] = bh; d_lo[o + 1] = bl; }
      }
    }
  }
}

__global__ void __launch_bounds__(256)
pack_weight_f32_kernel(const float* __restrict__ w, int Cout, int Cin, int kk, float* __restrict__ out) {
  const int64_t n = (int64_t)kk * Cout * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int tap = (int)(i / ((int64_t)Cin * Cout));
    out[i] = w[((int64_t)co * Cin + ci) * kk + tap];
  }
}

static inline int grid_for(int64_t n, int block, int max_blocks) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}


// ------------------------------------------------------------------------------------------
// Output path of sample_to_eval (runners/utils.py:67-74): NCHW fp32 in [-1,1] (or [0,1]) -> NHWC uint8, the
// reference's op order with one fp32 rounding per torch op, truncating cast (torch .to(uint8)) => byte-exact.
//   x = clamp(x*0.5 + 0.5, 0, 1)   (to_normal)      y = clamp(x*255 + 0.5, 0, 255)      u8 = (uint8) y
// One thread per pixel: coalesced reads per channel plane, C contiguous bytes written per pixel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
denorm_to_uint8_kernel(const float* __restrict__ x, int C, int64_t HW, int to_normal, uint8_t* __restrict__ out) {
  const int b = blockIdx.y;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    uint8_t* o = out + ((int64_t)b * HW + p) * C;
    for (int c = 0; c < C; ++c) {
      float v = x[((int64_t)b * C + c) * HW + p];
      if (to_normal) v = fminf(fmaxf(__fadd_rn(__fmul_rn(v, 0.5f), 0.5f), 0.0f), 1.0f);
      v = fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.0f), 0.5f), 0.0f), 255.0f);
      o[c] = (uint8_t)v;          // truncation toward zero, like tensor.to(torch.uint8)
    }
  }
}

// ------------------------------------------------------------------------------------------
// SpatialRescaler (encoders/modules.py:106-134): n x bilinear interpolate(scale 0.5) + optional 1x1 channel map.
// At scale 0.5 (align_corners=False) the source coordinate of output pixel d is 2d + 0.5: both interpolation weights
// are exactly 0.5, so one stage is h0*(w0*a + w1*b) + h1*(w0*c + w1*d) over the 2x2 block -- the products by 0.5 are
// exact, the three additions round.  n stages = the same expression applied recursively (sizes floor-halved per
// stage, so a level-L pixel always finds its 2x2 block inside the level-(L-1) image).
template <int L>
__device__ __forceinline__ float half_block(const float* __restrict__ p, int W0, int y, int x) {
  if constexpr (L == 0) {
    return p[(int64_t)y * W0 + x];
  } else {
    const float a = half_block<L - 1>(p, W0, 2 * y, 2 * x), b = half_block<L - 1>(p, W0, 2 * y, 2 * x + 1);
    const float c = half_block<L - 1>(p, W0, 2 * y + 1, 2 * x), d = half_block<L - 1>(p, W0, 2 * y + 1, 2 * x + 1);
    return __fadd_rn(__fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, a), __fmul_rn(0.5f, b))),
                     __fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, c), __fmul_rn(0.5f, d))));
  }
}

constexpr int RESCALE_MAX_C = 16;

template <int L>
__global__ void __launch_bounds__(128)
spatial_rescale_kernel(const float* __restrict__ src, int C, int H0, int W0, int Ho, int Wo,
                       const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                       float* __restrict__ out) {
  const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= Wo) return;
  float v[RESCALE_MAX_C];
#pragma unroll
  for (int c = 0; c < RESCALE_MAX_C; ++c)
    if (c < C) v[c] = half_block<L>(src + ((int64_t)b * C + c) * H0 * W0, W0, y, x);
  const int64_t plane = (int64_t)Ho * Wo, o = (int64_t)y * Wo + x;
  if (w == nullptr) {
#pragma unroll
    for (int c = 0; c < RESCALE_MAX_C; ++c)
      if (c < C) out[((int64_t)b * C + c) * plane + o] = v[c];
    return;
  }
  for (int co = 0; co < Cout; ++co) {
    float acc = bias ? bias[co] : 0.0f;
#pragma unroll
    for (int c = 0; c < RESCALE_MAX_C; ++c)
      if (c < C) acc = fmaf(w[co * C + c], v[c], acc);
    out[((int64_t)b * Cout + co) * plane + o] = acc;
  }
}

}  // namespace bbdm

using namespace bbdm;

extern "C" {

int bbdm_bridge_q_sample(const float* x0, const float* y, const float* noise, const int64_t* t,
                         const float* m_t, const float* variance_t, int num_timesteps,
                         int objective, float* x_t_out, float* objective_out, int B,
                         int64_t n_per_sample, void* stream) {
  BBDM_REQUIRE(x0 && y && noise && t && m_t && variance_t && x_t_out && objective_out, "q_sample: null pointer");
  BBDM_REQUIRE(B > 0 && B <= 65535 && n_per_sample > 0 && n_per_sample % 4 == 0,
               "q_sample: need 0 < B <= 65535 and n_per_sample %% 4 == 0 (got %d, %lld)", B, (long long)n_per_sample);
  BBDM_REQUIRE(num_timesteps > 0, "q_sample: num_timesteps must be > 0");
  unsigned long long* fault = device_fault_ptr();
  BBDM_REQUIRE(fault != nullptr, "q_sample: device fault word unavailable");
  const int64_t n4 = n_per_sample / 4;
  dim3 grid(grid_for(n4, 256, num_sms() * 8), B);
  cudaStream_t s = (cudaStream_t)stream;
#define BBDM_QL(O)                                                                              \
  q_sample_kernel<O><<<grid, 256, 0, s>>>((const float4*)x0, (const float4*)y, (const float4*)noise, t, \
                                          m_t, variance_t, (float4*)x_t_out, (float4*)objective_out, n4, num_timesteps, fault)
  if (objective == BBDM_OBJ_GRAD) BBDM_QL(BBDM_OBJ_GRAD);
  else if (objective == BBDM_OBJ_NOISE) BBDM_QL(BBDM_OBJ_NOISE);
  else if (objective == BBDM_OBJ_YSUBX) BBDM_QL(BBDM_OBJ_YSUBX);
  else BBDM_REQUIRE(false, "q_sample: unknown objective %d", objective);
#undef BBDM_QL
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

static int p_sample_launch(const float* x_t, const float* y, const float* eps, const float* noise,
                           BbdmPSampleCoef coef, const BbdmPSampleCoef* coef_dev, int objective,
                           int clip_denoised, int is_last, float* x_out, float* x0_out, int64_t n,
                           void* stream) {
  BBDM_REQUIRE(x_t && y && eps && x_out, "p_sample: null pointer");
  BBDM_REQUIRE(is_last || noise, "p_sample: noise required unless is_last");
  BBDM_REQUIRE(n > 0 && n % 4 == 0, "p_sample: n %% 4 != 0");
  BBDM_REQUIRE(objective >= 0 && objective <= 2, "p_sample: unknown objective %d", objective);
  const int64_t n4 = n / 4;
  const int grid = grid_for(n4, 256, num_sms() * 8);
  cudaStream_t s = (cudaStream_t)stream;
#define BBDM_PL(O, C, L)                                                                     \
  p_sample_kernel<O, C, L><<<grid, 256, 0, s>>>((const float4*)x_t, (const float4*)y,       \
                                                (const float4*)eps, (const float4*)noise, coef, \
                                                coef_dev, (float4*)x_out, (float4*)x0_out, n4)
#define BBDM_PL2(O)                                      \
  if (clip_denoised) { if (is_last) BBDM_PL(O, true, true); else BBDM_PL(O, true, false); } \
  else { if (is_last) BBDM_PL(O, false, true); else BBDM_PL(O, false, false); }
  if (objective == BBDM_OBJ_GRAD) { BBDM_PL2(BBDM_OBJ_GRAD) }
  else if (objective == BBDM_OBJ_NOISE) { BBDM_PL2(BBDM_OBJ_NOISE) }
  else { BBDM_PL2(BBDM_OBJ_YSUBX) }
#undef BBDM_PL2
#undef BBDM_PL
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_bridge_p_sample(const float* x_t, const float* y, const float* eps, const float* noise,
                         BbdmPSampleCoef coef, int objective, int clip_denoised, int is_last,
                         float* x_out, float* x0_out, int64_t n, void* stream) {
  return p_sample_launch(x_t, y, eps, noise, coef, nullptr, objective, clip_denoised, is_last, x_out, x0_out, n, stream);
}

int bbdm_bridge_p_sample_dev(const float* x_t, const float* y, const float* eps, const float* noise,
                             const BbdmPSampleCoef* coef_dev, int objective, int clip_denoised,
                             int is_last, float* x_out, float* x0_out, int64_t n, void* stream) {
  BBDM_REQUIRE(coef_dev != nullptr, "p_sample_dev: null coefficient pointer");
  BbdmPSampleCoef z = {0, 0, 0, 0, 0, 0, 0};
  return p_sample_launch(x_t, y, eps, noise, z, coef_dev, objective, clip_denoised, is_last, x_out, x0_out, n, stream);
}

int bbdm_nchw_to_nhwc_cat(const float* x, int c1, const float* ctx, int c2, int B, int H, int W,
                          float* out, void* stream) {
  BBDM_REQUIRE(x && out && c1 > 0 && (c2 == 0 || ctx) && B > 0 && B <= 65535, "nchw_to_nhwc_cat: bad args");
  const int64_t HW = (int64_t)H * W;
  dim3 grid(grid_for(HW, 256, 4096), B);
  nchw_to_nhwc_cat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, c1, ctx, ctx ? c2 : 0, HW, out);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_nhwc_to_nchw(const float* src, int B, int H, int W, int C, float* out, void* stream) {
  BBDM_REQUIRE(src && out && B > 0 && B <= 65535 && C > 0, "nhwc_to_nchw: bad args");
  const int64_t HW = (int64_t)H * W;
  dim3 grid(grid_for(HW, 256, 4096), B);
  nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, C, HW, out);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_gather_rows(const float* table, int rows, int width, const int64_t* idx, int B, float* out,
                     void* stream) {
  BBDM_REQUIRE(table && idx && out && rows > 0 && width > 0 && B > 0, "gather_rows: bad args");
  unsigned long long* fault = device_fault_ptr();
  BBDM_REQUIRE(fault != nullptr, "gather_rows: device fault word unavailable");
  gather_rows_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(table, rows, width, idx, out, fault);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_linear_f32(const float* x, const float* w, const float* bias, float* out, int B, int K,
                    int N, int act_in, int act_out, void* stream) {
  BBDM_REQUIRE(x && w && out && B > 0 && K > 0 && N > 0, "linear_f32: bad args");
  const int warps_per_block = 8;
  const int grid = (N + warps_per_block - 1) / warps_per_block;
  for (int b0 = 0; b0 < B; b0 += 8) {
    linear_f32_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, bias, out, B, K, N, act_in, act_out, b0);
    BBDM_LAUNCH_CHECK();
  }
  return BBDM_OK;
}

int bbdm_pack_weight_split_padded(const float* w, int Cout, int Cin, int k, int Cout_pad, void* w_hi, void* w_lo,
                                  void* stream) {
  BBDM_REQUIRE(w && w_hi && w_lo && Cout > 0 && Cin > 0 && (k == 1 || k == 3) && Cout_pad >= Cout,
               "pack_weight_split: bad args");
  const int64_t n = (int64_t)k * k * Cout * Cin;
  pack_weight_split_kernel<<<grid_for(n, 256, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      w, Cout, Cin, k * k, Cout_pad, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_pack_weight_split_taps(const float* w, int Cout, int Cin, int taps, void* w_hi, void* w_lo, void* stream) {
  BBDM_REQUIRE(w && w_hi && w_lo && Cout > 0 && Cin > 0 && taps > 0, "pack_weight_split_taps: bad args");
  const int64_t n = (int64_t)taps * Cout * Cin;
  pack_weight_split_kernel<<<grid_for(n, 256, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      w, Cout, Cin, taps, Cout, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_pack_weight_split_dgrad(const float* w, int Cout, int Cin, int k, void* w_hi, void* w_lo, void* stream) {
  BBDM_REQUIRE(w && w_hi && w_lo && Cout > 0 && Cin > 0 && (k == 1 || k == 3), "pack_weight_split_dgrad: bad args");
  const int64_t n = (int64_t)k * k * Cout * Cin;
  pack_weight_split_dgrad_kernel<<<grid_for(n, 256, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      w, Cout, Cin, k * k, (__nv_bfloat16*)w_hi, (__nv_bfloat16*)w_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_pack_weight_split_both(const float* w, int Cout, int Cin, int k, void* fwd_hi, void* fwd_lo, void* dgrad_hi,
                                void* dgrad_lo, void* stream) {
  BBDM_REQUIRE(w && Cout > 0 && Cin > 0 && (k == 1 || k == 3), "pack_weight_split_both: bad args");
  BBDM_REQUIRE((fwd_hi != nullptr) == (fwd_lo != nullptr) && (dgrad_hi != nullptr) == (dgrad_lo != nullptr) &&
               (fwd_hi || dgrad_hi), "pack_weight_split_both: give hi and lo of at least one layout");
  const dim3 grid((Cin + 31) / 32, (Cout + 31) / 32);
  BBDM_REQUIRE(grid.y <= 65535, "pack_weight_split_both: Cout too large");
  cudaStream_t s = (cudaStream_t)stream;
  if (k == 3)
    pack_weight_split_both_kernel<9><<<grid, 256, 0, s>>>(w, Cout, Cin, (__nv_bfloat16*)fwd_hi, (__nv_bfloat16*)fwd_lo,
                                                         (__nv_bfloat16*)dgrad_hi, (__nv_bfloat16*)dgrad_lo);
  else
    pack_weight_split_both_kernel<1><<<grid, 256, 0, s>>>(w, Cout, Cin, (__nv_bfloat16*)fwd_hi, (__nv_bfloat16*)fwd_lo,
                                                         (__nv_bfloat16*)dgrad_hi, (__nv_bfloat16*)dgrad_lo);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

int bbdm_pack_weight_split(const float* w, int Cout, int Cin, int k, void* w_hi, void* w_lo, void* stream) {
  return bbdm_pack_weight_split_padded(w, Cout, Cin, k, Cout, w_hi, w_lo, stream);
}

int bbdm_pack_weight_f32(const float* w, int Cout, int Cin, int k, float* out, void* stream) {
  BBDM_REQUIRE(w && out && Cout > 0 && Cin > 0 && (k == 1 || k == 3), "pack_weight_f32: bad args");
  const int64_t n = (int64_t)k * k * Cout * Cin;
  pack_weight_f32_kernel<<<grid_for(n, 256, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, k * k, out);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

/* images [B,C,H,W] fp32 -> uint8 [B,H,W,C] (the PNG byte layout PIL takes). */
int bbdm_denorm_to_uint8(const float* images, int B, int C, int H, int W, int to_normal, uint8_t* out, void* stream) {
  BBDM_REQUIRE(images && out && B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0, "denorm_to_uint8: bad args");
  const int64_t HW = (int64_t)H * W;
  dim3 grid(grid_for(HW, 256, num_sms() * 8), B);
  denorm_to_uint8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, C, HW, to_normal, out);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

/* SpatialRescaler: src [B,C,H,W] fp32 -> n_stages x bilinear(0.5) -> optional 1x1 map (w [Cout,C], bias [Cout] or NULL)
 * -> out [B, Cout or C, H >> n, W >> n] fp32 (NCHW, what the UNet stem concatenates). */
int bbdm_spatial_rescale(const float* src, int B, int C, int H, int W, int n_stages, const float* w, const float* bias,
                         int Cout, float* out, void* stream) {
  BBDM_REQUIRE(src && out && B > 0 && B <= 65535 && C > 0 && C <= RESCALE_MAX_C && H > 0 && W > 0,
               "spatial_rescale: bad args (need 0 < C <= %d)", RESCALE_MAX_C);
  BBDM_REQUIRE(n_stages >= 0 && n_stages <= 4, "spatial_rescale: n_stages %d not in [0, 4]", n_stages);
  BBDM_REQUIRE(w != nullptr || bias == nullptr, "spatial_rescale: bias without weights");
  BBDM_REQUIRE(w == nullptr || Cout > 0, "spatial_rescale: Cout must be > 0 with a channel map");
  int Ho = H, Wo = W;
  for (int i = 0; i < n_stages; ++i) { Ho /= 2; Wo /= 2; }
  BBDM_REQUIRE(Ho > 0 && Wo > 0 && Ho <= 65535, "spatial_rescale: %dx%d is too small for %d stages", H, W, n_stages);
  dim3 grid((Wo + 127) / 128, Ho, B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n_stages) {
    case 0: spatial_rescale_kernel<0><<<grid, 128, 0, st>>>(src, C, H, W, Ho, Wo, w, bias, Cout, out); break;
    case 1: spatial_rescale_kernel<1><<<grid, 128, 0, st>>>(src, C, H, W, Ho, Wo, w, bias, Cout, out); break;
    case 2: spatial_rescale_kernel<2><<<grid, 128, 0, st>>>(src, C, H, W, Ho, Wo, w, bias, Cout, out); break;
    case 3: spatial_rescale_kernel<3><<<grid, 128, 0, st>>>(src, C, H, W, Ho, Wo, w, bias, Cout, out); break;
    default: spatial_rescale_kernel<4><<<grid, 128, 0, st>>>(src, C, H, W, Ho, Wo, w, bias, Cout, out); break;
  }
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // extern "C"
