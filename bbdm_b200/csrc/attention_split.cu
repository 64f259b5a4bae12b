// Flash-style attention core on pre-split bf16 operand planes (the qkv 1x1 conv's epilogue
// writes qkv as hi/lo bf16 planes, so nothing is converted or re-split here).
//
//   S = (q . k) * D^-1/2        (= (q D^-1/4) . (k D^-1/4) of openaimodel.py:359-375)
//   O = softmax_fp32(S) v
//
// CTA = 8 warps x 16 query rows = 128 queries of one (batch, head); KV tiles of 64 keys,
// double-buffered with cp.async (16-byte chunks, zero-fill past T); fragments via ldmatrix
// (K plain, V transposed); split-bf16 x3 products on mma.sync.m16n8k16 with fp32 accumulate;
// every KV tile's P.V product starts from a zero accumulator and is added to O with a
// round-to-nearest fp32 add (the tensor core's own accumulate truncates).
#include "common.cuh"

namespace bbdm {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void split2p(float x, float y, uint32_t& hi, uint32_t& lo) { split2x(x, y, hi, lo); }

// Operand addressing: queries come from (q_hi, q_lo) [B][Tq][rs_q] at column qoff0 + head*hstride, keys / values from
// (kv_hi, kv_lo) [B][T][rs_kv] at koff0 / voff0 + head*hstride.  Self-attention passes the same planes for both
// (T == Tq); cross-attention (SpatialTransformer.attn2, reference attention.py:152-192) a separate K|V tensor with
// its own length T.
struct AttnOperands {
  const __nv_bfloat16* q_hi; const __nv_bfloat16* q_lo; int64_t rs_q; int qoff0;
  const __nv_bfloat16* kv_hi; const __nv_bfloat16* kv_lo; int64_t rs_kv; int koff0, voff0;
  int hstride, Tq;
};

template <int D>
__global__ void __launch_bounds__(256)
attention_split_kernel(const AttnOperands ops, int T, int C, int heads, float scale_log2,
                       float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
                       __nv_bfloat16* __restrict__ out_lo) {
  constexpr int KT = 64;                  // keys per tile
  constexpr int KS = D / 16;              // k-steps over head_dim
  constexpr int LD = D + 8;               // padded smem row (elements): 16 B skew, conflict-free ldmatrix
  constexpr int TILE = KT * LD;           // elements per plane tile
  extern __shared__ __align__(16) __nv_bfloat16 sm[];   // [2 stages][Kh, Kl, Vh, Vl][KT][LD]

  const int bh = blockIdx.y;
  const int b = bh / heads, head = bh % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t rs = ops.rs_kv, rs_q = ops.rs_q;      // row strides (elements)
  const int Tq = ops.Tq;
  const int qoff = ops.qoff0 + head * ops.hstride, koff = ops.koff0 + head * ops.hstride,
            voff = ops.voff0 + head * ops.hstride;
  const __nv_bfloat16* base_hi = ops.kv_hi + (int64_t)b * T * rs;
  const __nv_bfloat16* base_lo = ops.kv_lo + (int64_t)b * T * rs;
  const __nv_bfloat16* qbase_hi = ops.q_hi + (int64_t)b * Tq * rs_q;
  const __nv_bfloat16* qbase_lo = ops.q_lo + (int64_t)b * Tq * rs_q;

  // ---- stage loader: 4 plane tiles x KT rows x (D/8) 16-byte chunks -------------------------
  auto load_tile = [&](int stage, int k0) {
    constexpr int CPR = D / 8;            // chunks per row
    constexpr int N = 4 * KT * CPR;
    __nv_bfloat16* sbase = sm + stage * 4 * TILE;
    for (int i = threadIdx.x; i < N; i += 256) {
      const int plane = i / (KT * CPR), rem = i % (KT * CPR);
      const int key = rem / CPR, ch = rem % CPR;
      const int kk = k0 + key;
      const bool ok = kk < T;
      const __nv_bfloat16* src = ((plane & 1) ? base_lo : base_hi) + (int64_t)(ok ? kk : 0) * rs +
                                 ((plane < 2) ? koff : voff) + ch * 8;
      cp_async16(smem_addr(sbase + plane * TILE + key * LD + ch * 8), src, ok ? 16 : 0);
    }
  };

  // ---- Q fragments for this warp's 16 rows (registers, both planes) ----------------------------
  const int q0 = blockIdx.x * 128 + warp * 16;
  uint32_t qh[KS][4], ql[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
      for (int r2 = 0; r2 < 2; ++r2) {
        const int qr = q0 + g + r2 * 8;
        uint32_t vh = 0, vl = 0;
        if (qr < Tq) {
          const int64_t o = qr * rs_q + qoff + ks * 16 + h2 * 8 + 2 * t;
          vh = *reinterpret_cast<const uint32_t*>(qbase_hi + o);
          vl = *reinterpret_cast<const uint32_t*>(qbase_lo + o);
        }
        qh[ks][h2 * 2 + r2] = vh;
        ql[ks][h2 * 2 + r2] = vl;
      }

  float o[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  const int n_tiles = (T + KT - 1) / KT;
  load_tile(0, 0);
  cp_commit();
  for (int it = 0; it < n_tiles; ++it) {
    const int stage = it & 1;
    if (it + 1 < n_tiles) load_tile(stage ^ 1, (it + 1) * KT);
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    const __nv_bfloat16* Kh = sm + stage * 4 * TILE;
    const uint32_t kh_a = smem_addr(Kh), kl_a = kh_a + TILE * 2, vh_a = kh_a + 2 * TILE * 2, vl_a = kh_a + 3 * TILE * 2;
    const int k0 = it * KT;

    // ---- S = Q K^T : per 8-key n-tile j, ldmatrix.x4 covers two k-steps (b0,b1 | b0,b1) -----------
    float s[KT / 8][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      // lane -> row address: matrices (m = lane>>3): d offset m*8 within a 32-wide d span, row = key j*8 + (lane&7)
      const uint32_t roff = (uint32_t)(((j * 8 + (lane & 7)) * LD + (lane >> 3) * 8) * 2);
#pragma unroll
      for (int k2 = 0; k2 < (KS + 1) / 2; ++k2) {
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        if (D >= 32) {
          ldsm_x4(kh_a + roff + k2 * 64, h0, h1, h2, h3);
          ldsm_x4(kl_a + roff + k2 * 64, l0, l1, l2, l3);
        } else {   // D == 16: only matrices 0,1 are in range; re-read them for 2,3 (unused)
          const uint32_t r16 = (uint32_t)(((j * 8 + (lane & 7)) * LD + ((lane >> 3) & 1) * 8) * 2);
          ldsm_x4(kh_a + r16, h0, h1, h2, h3);
          ldsm_x4(kl_a + r16, l0, l1, l2, l3);
        }
        mma16816(s[j], ql[2 * k2], h0, h1);
        mma16816(s[j], qh[2 * k2], l0, l1);
        mma16816(s[j], qh[2 * k2], h0, h1);
        if (2 * k2 + 1 < KS) {
          mma16816(s[j], ql[2 * k2 + 1], h2, h3);
          mma16816(s[j], qh[2 * k2 + 1], l2, l3);
          mma16816(s[j], qh[2 * k2 + 1], h2, h3);
        }
      }
    }
    // ---- scale (log2 domain), mask keys >= T, online softmax ---------------------------------------
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      const int key = k0 + j * 8 + 2 * t;
      s[j][0] *= scale_log2; s[j][1] *= scale_log2; s[j][2] *= scale_log2; s[j][3] *= scale_log2;
      if (key >= T) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (key + 1 >= T) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : ex2f(m_run[r] - m_new);
      m_run[r] = m_new;
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int j = 0; j < D / 8; ++j) { o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1]; }
    uint32_t ph[KT / 16][4], pl[KT / 16][4];
#pragma unroll
    for (int j = 0; j < KT / 8; ++j) {
      s[j][0] = ex2f(s[j][0] - m_run[0]); s[j][1] = ex2f(s[j][1] - m_run[0]);
      s[j][2] = ex2f(s[j][2] - m_run[1]); s[j][3] = ex2f(s[j][3] - m_run[1]);
      l_run[0] += s[j][0] + s[j][1];
      l_run[1] += s[j][2] + s[j][3];
      // C-fragment of two adjacent n-tiles == A-fragment of one 16-key k-step
      split2p(s[j][0], s[j][1], ph[j >> 1][(j & 1) * 2 + 0], pl[j >> 1][(j & 1) * 2 + 0]);   // row g
      split2p(s[j][2], s[j][3], ph[j >> 1][(j & 1) * 2 + 1], pl[j >> 1][(j & 1) * 2 + 1]);   // row g+8
    }
    // ---- O += P V : V^T fragments by ldmatrix.trans; x4 = (keys 0-7 | 8-15) x (d-tile jd | jd+1) ------
#pragma unroll
    for (int jd2 = 0; jd2 < D / 16; ++jd2) {
      float ot0[4] = {0.f, 0.f, 0.f, 0.f}, ot1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < KT / 16; ++kk) {
        // matrix m = lane>>3: key block (m&1)*8, d block (m>>1)*8 ; row within = lane&7 (a key)
        const uint32_t voff2 = (uint32_t)(((kk * 16 + (lane >> 3 & 1) * 8 + (lane & 7)) * LD + jd2 * 16 + (lane >> 4) * 8) * 2);
        uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
        ldsm_x4_t(vh_a + voff2, h0, h1, h2, h3);
        ldsm_x4_t(vl_a + voff2, l0, l1, l2, l3);
        mma16816(ot0, pl[kk], h0, h1);
        mma16816(ot0, ph[kk], l0, l1);
        mma16816(ot0, ph[kk], h0, h1);
        mma16816(ot1, pl[kk], h2, h3);
        mma16816(ot1, ph[kk], l2, l3);
        mma16816(ot1, ph[kk], h2, h3);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) { o[2 * jd2][c] += ot0[c]; o[2 * jd2 + 1][c] += ot1[c]; }
    }
    __syncthreads();     // all warps done with this stage before it is refilled
  }
  cp_wait<0>();

  // ---- normalise and store -------------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r2 = 0; r2 < 2; ++r2) {
    const int qr = q0 + g + r2 * 8;
    if (qr >= Tq) continue;
    const float inv = 1.0f / l_run[r2];
    const int64_t off = ((int64_t)b * Tq + qr) * C + head * D + 2 * t;
#pragma unroll
    for (int jd = 0; jd < D / 8; ++jd) {
      const float x = o[jd][2 * r2] * inv, y = o[jd][2 * r2 + 1] * inv;
      if (out_f32) *reinterpret_cast<float2*>(out_f32 + off + jd * 8) = make_float2(x, y);
      if (out_hi) {
        uint32_t h, l;
        split2p(x, y, h, l);
        *reinterpret_cast<uint32_t*>(out_hi + off + jd * 8) = h;
        *reinterpret_cast<uint32_t*>(out_lo + off + jd * 8) = l;
      }
    }
  }
}

}  // namespace bbdm

using namespace bbdm;

static int launch_attention_split(const AttnOperands& ops, int B, int T, int C, int heads, float* out_f32, void* out_hi,
                                  void* out_lo, void* stream) {
  const int D = C / heads;
  BBDM_REQUIRE((int64_t)B * heads <= 65535, "attention_split: B*heads too large");
  dim3 grid((ops.Tq + 127) / 128, B * heads);
  cudaStream_t s = (cudaStream_t)stream;
  // softmax((q s)(k s)) with s = D^-1/4  ==  2^(log2(e) * D^-1/2 * (q.k) - max)
  const float scale_log2 = (float)(1.4426950408889634 / sqrt((double)D));
  __nv_bfloat16* oh = (__nv_bfloat16*)out_hi;
  __nv_bfloat16* ol = (__nv_bfloat16*)out_lo;
#define BBDM_AL(DD)                                                                                       \
  {                                                                                                       \
    const size_t smem = (size_t)2 * 4 * 64 * (DD + 8) * 2;                                                \
    static DeviceOnce cfgd;                                                                               \
    if (cfgd.need()) {                                                                                    \
      BBDM_CUDA_CHECK(cudaFuncSetAttribute(attention_split_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      cfgd.mark();                                                                                        \
    }                                                                                                     \
    attention_split_kernel<DD><<<grid, 256, smem, s>>>(ops, T, C, heads, scale_log2, out_f32, oh, ol);   \
  }
  if (D == 64) BBDM_AL(64)
  else if (D == 32) BBDM_AL(32)
  else if (D == 16) BBDM_AL(16)
  else {
    set_error("attention_split: head_dim %d not supported (16, 32, 64)", D);
    return BBDM_E_UNSUPPORTED;
  }
#undef BBDM_AL
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

extern "C" int bbdm_attention_split(const void* qkv_hi, const void* qkv_lo, int B, int T, int C, int heads,
                                    int order, float* out_f32, void* out_hi, void* out_lo, void* stream) {
  BBDM_REQUIRE(qkv_hi && qkv_lo && (out_f32 || (out_hi && out_lo)), "attention_split: null pointer");
  BBDM_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "attention_split: hi/lo must come in pairs");
  BBDM_REQUIRE(B > 0 && T > 0 && heads > 0 && C % heads == 0, "attention_split: bad shape");
  BBDM_REQUIRE(order == 0 || order == 1, "attention_split: order must be 0 (legacy) or 1");
  const int D = C / heads;
  AttnOperands ops;
  ops.q_hi = ops.kv_hi = (const __nv_bfloat16*)qkv_hi;
  ops.q_lo = ops.kv_lo = (const __nv_bfloat16*)qkv_lo;
  ops.rs_q = ops.rs_kv = 3 * (int64_t)C;
  ops.Tq = T;
  if (order == 0) { ops.hstride = 3 * D; ops.qoff0 = 0; ops.koff0 = D; ops.voff0 = 2 * D; }
  else { ops.hstride = D; ops.qoff0 = 0; ops.koff0 = C; ops.voff0 = 2 * C; }
  return launch_attention_split(ops, B, T, C, heads, out_f32, out_hi, out_lo, stream);
}

extern "C" int bbdm_attention_cross(const void* q_hi, const void* q_lo, const void* kv_hi, const void* kv_lo, int B,
                                    int Tq, int Tkv, int C, int heads, float* out_f32, void* out_hi, void* out_lo,
                                    void* stream) {
  BBDM_REQUIRE(q_hi && q_lo && kv_hi && kv_lo && (out_f32 || (out_hi && out_lo)), "attention_cross: null pointer");
  BBDM_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "attention_cross: hi/lo must come in pairs");
  BBDM_REQUIRE(B > 0 && Tq > 0 && Tkv > 0 && heads > 0 && C % heads == 0, "attention_cross: bad shape");
  AttnOperands ops;
  ops.q_hi = (const __nv_bfloat16*)q_hi; ops.q_lo = (const __nv_bfloat16*)q_lo;
  ops.kv_hi = (const __nv_bfloat16*)kv_hi; ops.kv_lo = (const __nv_bfloat16*)kv_lo;
  ops.rs_q = C; ops.rs_kv = 2 * (int64_t)C;
  ops.Tq = Tq;
  ops.hstride = C / heads; ops.qoff0 = 0; ops.koff0 = 0; ops.voff0 = C;
  return launch_attention_split(ops, B, Tkv, C, heads, out_f32, out_hi, out_lo, stream);
}
