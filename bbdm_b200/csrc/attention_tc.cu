// FlashAttention-style, warp-specialised attention core on tcgen05 tensor cores (sm_100a).
// AttentionBlock of the reference: openaimodel.py:281-413 (softmax((q s)(k s)^T) v, s = D^-1/4).
//
//   * one CTA = 128 queries of one (batch, head); KV tiles of 64 keys; head_dim D = 64.
//   * operands are the split-bf16 planes the qkv 1x1 conv wrote (qkv_hi/qkv_lo [B,T,3C]);
//     TMA (3-D tiled maps, SWIZZLE_128B) stages Q once and K/V tiles through a 4-stage ring.
//   * warp0 = TMA producer; warp1 = tcgen05.mma issuer (one lane) + TMEM owner;
//     warps 2-9 = TWO softmax / correction warpgroups, ONE QUERY ROW PER THREAD (no shuffles):
//     group g owns the KV tiles j = g (mod 2) and the S/P/O buffers g, so the softmax of tile j+1
//     runs concurrently with that of tile j (a single group left the tensor pipe 78 % idle); the
//     two partial (max, sum, O) states are merged once at the end.  Three S accumulators are in flight
//     (S_{j+3} is issued right behind P_j V_j), so a group never waits for its next scores: with two the
//     issue order P_{j-2}V_{j-2} -> S_j put two MMA groups between a group's tiles (34 % of all stall
//     samples sat in the s_full wait, tensor pipe 25 % active).
//   * S_j = Q K_j^T      : M=128 x N=64 x K=64, A=Q (K-major), B=K_j (K-major)      -> TMEM S[j%3]
//     O_j = P_j V_j      : M=128 x N=64 x K=64, A=P_j (K-major, written to smem by the softmax
//                          warps in the UMMA swizzle), B=V_j as an MN-major operand      -> TMEM O[j%2]
//     every product is split-bf16 x3 (lo.hi + hi.lo + hi.hi, fp32 accumulate).
//   * each O_j starts from a zero accumulator; the softmax warps fold it into fp32 REGISTER
//     accumulators with round-to-nearest adds and the online-softmax rescale (the tensor core's own
//     accumulate truncates) -- S_{j+1} and P_j V_j overlap the softmax of tile j.
//   * all mbarrier waits are watchdogged (device fault word, no GPU hang).
#include "tc_common.cuh"

namespace bbdm {

constexpr int AT_D = 64;
constexpr int AT_BQ = 128;                // queries per CTA
constexpr int AT_BK = 64;                 // keys per tile
constexpr int AT_STAGES = 4;
constexpr int AT_SBUF = 3;                // S accumulators in flight: S_{j+3} is issued right after P_j V_j
constexpr uint32_t AT_Q_BYTES = AT_BQ * AT_D * 2;     // 16 KiB per plane
constexpr uint32_t AT_KV_BYTES = AT_BK * AT_D * 2;    // 8 KiB per plane tile
constexpr uint32_t AT_P_BYTES = AT_BQ * AT_BK * 2;    // 16 KiB per plane
constexpr uint32_t AT_STAGE_BYTES = 4 * AT_KV_BYTES;  // K_hi K_lo V_hi V_lo
constexpr uint32_t AT_SMEM = 2 * AT_Q_BYTES + AT_STAGES * AT_STAGE_BYTES + 2 * 2 * AT_P_BYTES + 1024;

// MN-major SWIZZLE_128B descriptor (B operand = V tile [key][d], d contiguous): rows of 128 B are
// K (= key) indices, 8-row groups 1024 B apart (SBO); a single 64-wide MN atom so LBO is unused.
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;        // LBO: stride between 64-element MN atoms (not reached)
  d |= (uint64_t)(1024 >> 4) << 32;        // SBO: stride between 8-key groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct AttnParams {
  int T, C, heads, order;
  float scale_log2;
  float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  unsigned long long* fault;
};

__global__ void __launch_bounds__(320, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                    const __grid_constant__ CUtensorMap map_kv_hi, const __grid_constant__ CUtensorMap map_kv_lo,
                    const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[1 + 2 * AT_STAGES + 2 * AT_SBUF + 4];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_s;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_hi_a = base, q_lo_a = base + AT_Q_BYTES;
  const uint32_t kv_a = base + 2 * AT_Q_BYTES;                       // [stage][Kh Kl Vh Vl]
  const uint32_t p_a = kv_a + AT_STAGES * AT_STAGE_BYTES;            // [buf][Ph Pl]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_q = smem_u32(&bars[0]);
  const uint32_t bar_kvf = smem_u32(&bars[1]);                        // [STAGES]
  const uint32_t bar_kve = smem_u32(&bars[1 + AT_STAGES]);            // [STAGES]
  const uint32_t bar_sf = smem_u32(&bars[1 + 2 * AT_STAGES]);         // s_full[AT_SBUF]
  const uint32_t bar_se = bar_sf + 8 * AT_SBUF;                       // s_empty[AT_SBUF]
  const uint32_t bar_pf = bar_se + 8 * AT_SBUF;                       // p_full[2]
  const uint32_t bar_pe = bar_pf + 16;                                // p_empty[2]
  __shared__ __align__(8) uint64_t bars_o[4];
  const uint32_t bar_of = smem_u32(&bars_o[0]);                       // o_full[2]
  const uint32_t bar_oe = bar_of + 16;                                // o_empty[2]
  volatile int* abort_flag = &abort_s;

  if (threadIdx.x == 0) {
    abort_s = 0;
    mbar_init(bar_q, 1);
    for (int i = 0; i < AT_STAGES; ++i) { mbar_init(bar_kvf + 8 * i, 1); mbar_init(bar_kve + 8 * i, 1); }
    for (int i = 0; i < AT_SBUF; ++i) { mbar_init(bar_sf + 8 * i, 1); mbar_init(bar_se + 8 * i, 4); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_pf + 8 * i, 4); mbar_init(bar_pe + 8 * i, 1);
      mbar_init(bar_of + 8 * i, 1); mbar_init(bar_oe + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int bh = blockIdx.y, b = bh / p.heads, head = bh % p.heads;
  int qoff, koff, voff;
  if (p.order == 0) { qoff = head * 3 * AT_D; koff = qoff + AT_D; voff = qoff + 2 * AT_D; }
  else { qoff = head * AT_D; koff = p.C + head * AT_D; voff = 2 * p.C + head * AT_D; }
  const int q0 = blockIdx.x * AT_BQ;
  const int n_tiles = (p.T + AT_BK - 1) / AT_BK;

  if (warp == 0) {
    // ================================ TMA producer ============================================
    if (lane == 0) {
      mbar_expect_tx(bar_q, 2 * AT_Q_BYTES);
      tma_load_3d(q_hi_a, &map_q_hi, bar_q, qoff, q0, b);
      tma_load_3d(q_lo_a, &map_q_lo, bar_q, qoff, q0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % AT_STAGES, u = j / AT_STAGES;
        mbar_wait(bar_kve + 8 * st, (u & 1) ^ 1, abort_flag, p.fault, 0xB1000000ull | (unsigned)j);
        const uint32_t sb = kv_a + st * AT_STAGE_BYTES, full = bar_kvf + 8 * st;
        mbar_expect_tx(full, AT_STAGE_BYTES);
        tma_load_3d(sb, &map_kv_hi, full, koff, j * AT_BK, b);
        tma_load_3d(sb + AT_KV_BYTES, &map_kv_lo, full, koff, j * AT_BK, b);
        tma_load_3d(sb + 2 * AT_KV_BYTES, &map_kv_hi, full, voff, j * AT_BK, b);
        tma_load_3d(sb + 3 * AT_KV_BYTES, &map_kv_lo, full, voff, j * AT_BK, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==============================================
    // The whole warp runs this loop in uniform control flow (waits, descriptor arithmetic -> uniform datapath);
    // only the tcgen05.mma / commit instructions sit under the elected lane's predicate.  With the loop inside
    // `if (lane == 0)` every descriptor went through vector registers and 5 R2UR per MMA: 447 instructions per KV
    // tile on one lane, which -- not the tensor pipe (25 % active) -- set the tile period.
    const bool leader = elect_one_sync();
    {
      // D=f32, A=B=bf16, M=128, N=64; PV additionally: B is MN-major (bit 16)
      constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t IDESC_PV = IDESC_S | (1u << 16);
      const uint64_t dq_hi = make_sw128_desc(q_hi_a), dq_lo = make_sw128_desc(q_lo_a);
      auto issue_s = [&](int j) {
        const int st = j % AT_STAGES, u = j / AT_STAGES, sbuf = j % AT_SBUF, su = j / AT_SBUF;
        mbar_wait(bar_kvf + 8 * st, u & 1, abort_flag, p.fault, 0xB2000000ull | (unsigned)j);
        mbar_wait(bar_se + 8 * sbuf, (su & 1) ^ 1, abort_flag, p.fault, 0xB3000000ull | (unsigned)j);
        tc_fence_after();
        const uint32_t sb = kv_a + st * AT_STAGE_BYTES;
        const uint64_t dk_hi = make_sw128_desc(sb), dk_lo = make_sw128_desc(sb + AT_KV_BYTES);
        const uint32_t d_tmem = tmem_base + sbuf * 64;
        if (leader) {
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) {
            const uint64_t ko = (uint64_t)(k * 32 >> 4);
            tc_mma_bf16(d_tmem, dq_lo + ko, dk_hi + ko, IDESC_S, k ? 1u : 0u);
            tc_mma_bf16(d_tmem, dq_hi + ko, dk_lo + ko, IDESC_S, 1u);
            tc_mma_bf16(d_tmem, dq_hi + ko, dk_hi + ko, IDESC_S, 1u);
          }
          tc_commit(bar_sf + 8 * sbuf);
        }
        __syncwarp();
      };
      mbar_wait(bar_q, 0, abort_flag, p.fault, 0xB0000000ull);
      issue_s(0);
      for (int j = 1; j < AT_SBUF && j < n_tiles; ++j) issue_s(j);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % AT_STAGES, pbuf = j & 1, pu = j >> 1;
        mbar_wait(bar_pf + 8 * pbuf, pu & 1, abort_flag, p.fault, 0xB4000000ull | (unsigned)j);
        mbar_wait(bar_oe + 8 * pbuf, (pu & 1) ^ 1, abort_flag, p.fault, 0xB5000000ull | (unsigned)j);
        tc_fence_after();
        const uint32_t sb = kv_a + st * AT_STAGE_BYTES;
        const uint32_t pb = p_a + pbuf * 2 * AT_P_BYTES;
        const uint64_t dp_hi = make_sw128_desc(pb), dp_lo = make_sw128_desc(pb + AT_P_BYTES);
        const uint32_t d_tmem = tmem_base + AT_SBUF * 64 + pbuf * 64;
        const uint64_t dv0_hi = make_sw128_mn_desc(sb + 2 * AT_KV_BYTES), dv0_lo = make_sw128_mn_desc(sb + 3 * AT_KV_BYTES);
        if (leader) {
#pragma unroll
          for (int k = 0; k < AT_BK / 16; ++k) {
            const uint64_t ka = (uint64_t)(k * 32 >> 4);                // A = P: +32 B per 16 keys (K-major)
            const uint64_t kv = (uint64_t)(k * 16 * 128 >> 4);          // B = V: +16 key rows of 128 B (MN-major)
            tc_mma_bf16(d_tmem, dp_lo + ka, dv0_hi + kv, IDESC_PV, k ? 1u : 0u);
            tc_mma_bf16(d_tmem, dp_hi + ka, dv0_lo + kv, IDESC_PV, 1u);
            tc_mma_bf16(d_tmem, dp_hi + ka, dv0_hi + kv, IDESC_PV, 1u);
          }
          tc_commit(bar_of + 8 * pbuf);         // O_j ready
          tc_commit(bar_kve + 8 * st);          // K_j / V_j slot free
          tc_commit(bar_pe + 8 * pbuf);         // P buffer free
        }
        __syncwarp();
        if (j + AT_SBUF < n_tiles) issue_s(j + AT_SBUF);   // keeps S two-to-three tiles ahead of the softmax groups
      }
    }
    __syncwarp();
  } else {
    // ================================ softmax / correction / epilogue ===========================
    const int grp = (warp - 2) >> 2;          // softmax group 0 / 1 <-> KV tiles of parity grp, buffers grp
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float o_reg[AT_D];
#pragma unroll
    for (int i = 0; i < AT_D; ++i) o_reg[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    auto fold_o = [&](int j) {                // o_reg += O_j  (both relative to the same running max)
      const int obuf = j & 1, ou = j >> 1;
      mbar_wait(bar_of + 8 * obuf, ou & 1, abort_flag, p.fault, 0xB6000000ull | (unsigned)j);
      tc_fence_after();
      // both halves' TMEM loads overlap the adds of the other half (tcgen05.ld latency is several hundred cycles)
      uint32_t va[32], vb[32];
      tc_ld32(lane_addr + AT_SBUF * 64 + obuf * 64, va);
      tc_wait_ld();
      tc_ld32(lane_addr + AT_SBUF * 64 + obuf * 64 + 32, vb);
#pragma unroll
      for (int i = 0; i < 32; ++i) o_reg[i] += __uint_as_float(va[i]);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_reg[32 + i] += __uint_as_float(vb[i]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_oe + 8 * obuf);
    };

    for (int j = grp; j < n_tiles; j += 2) {
      const int sbuf = j % AT_SBUF, su = j / AT_SBUF;
      const int k0 = j * AT_BK;
      mbar_wait(bar_sf + 8 * sbuf, su & 1, abort_flag, p.fault, 0xB7000000ull | (unsigned)j);
      tc_fence_after();
      // S_j stays in TMEM and is read twice (row maximum, then exponentials) to keep the register footprint
      // small; every load is issued one step ahead so its latency hides behind the arithmetic of the other half
      uint32_t va[32], vb[32];
      const uint32_t s_addr = lane_addr + sbuf * 64;
      float mx = -INFINITY;
      tc_ld32(s_addr, va);
      tc_wait_ld();
      tc_ld32(s_addr + 32, vb);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (k0 + i < p.T) mx = fmaxf(mx, __uint_as_float(va[i]) * p.scale_log2);
      tc_wait_ld();
      tc_ld32(s_addr, va);                      // first half again, for the exponentials
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (k0 + 32 + i < p.T) mx = fmaxf(mx, __uint_as_float(vb[i]) * p.scale_log2);
      const float m_new = fmaxf(m_run, mx);
      const float corr = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m_new);
      m_run = m_new;

      // P_j = 2^(S_j - m) -> shared memory in the UMMA K-major SWIZZLE_128B layout (row = query)
      const int pbuf = j & 1, pu = j >> 1;
      mbar_wait(bar_pe + 8 * pbuf, (pu & 1) ^ 1, abort_flag, p.fault, 0xB8000000ull | (unsigned)j);
      const uint32_t ph = p_a + pbuf * 2 * AT_P_BYTES + row * 128, pl = ph + AT_P_BYTES;
      float rs = 0.f;
      auto emit_half = [&](uint32_t (&v)[32], int h) {
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[i] = (k0 + h * 32 + i < p.T) ? ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -m_new)) : 0.f;
          rs += e[i];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {           // 16-byte chunk (h*4 + c) = keys 8*(h*4+c) ..
          uint2 h0, l0, h1, l1;
          split4(make_float4(e[8 * c], e[8 * c + 1], e[8 * c + 2], e[8 * c + 3]), h0, l0);
          split4(make_float4(e[8 * c + 4], e[8 * c + 5], e[8 * c + 6], e[8 * c + 7]), h1, l1);
          const uint32_t off = (uint32_t)(((h * 4 + c) ^ (row & 7)) * 16);
          st_shared_v4(ph + off, h0.x, h0.y, h1.x, h1.y);
          st_shared_v4(pl + off, l0.x, l0.y, l1.x, l1.y);
        }
      };
      tc_wait_ld();
      tc_ld32(s_addr + 32, vb);
      emit_half(va, 0);
      tc_wait_ld();
      emit_half(vb, 1);
      l_run = l_run * corr + rs;
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_se + 8 * sbuf); mbar_arrive(bar_pf + 8 * pbuf); }

      // fold this group's previous tile, then rescale to the new running max
      if (j >= 2) fold_o(j - 2);
#pragma unroll
      for (int i = 0; i < AT_D; ++i) o_reg[i] *= corr;
    }
    {
      // last tile of this group (if it had any)
      const int last = ((n_tiles - 1 - grp) >= 0) ? (n_tiles - 1 - ((n_tiles - 1 - grp) & 1)) : -1;
      if (last >= grp) fold_o(last);
    }

    // ---- merge the two groups' partial softmax states (group 1 -> shared memory -> group 0) ------------
    // after the first barrier every MMA has completed (each group waited on its last O tile), so the
    // K/V ring is no longer read by the tensor core and can carry the exchange
    float* xch = reinterpret_cast<float*>(smem_raw + (kv_a - smem_u32(smem_raw)));
    constexpr int XS = AT_D + 2;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (grp == 1) {
      float* r = xch + row * XS;
      r[0] = m_run; r[1] = l_run;
#pragma unroll
      for (int i = 0; i < AT_D; ++i) r[2 + i] = o_reg[i];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (grp == 0) {
      const float* r = xch + row * XS;
      const float m1 = r[0], l1 = r[1];
      const float m = fmaxf(m_run, m1);
      const float w0 = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run - m);
      const float w1 = (m1 == -INFINITY) ? 0.f : ex2_approx(m1 - m);
      const float inv = 1.0f / (l_run * w0 + l1 * w1);
      const int qr = q0 + row;
      if (qr < p.T) {
        const int64_t off = ((int64_t)b * p.T + qr) * p.C + head * AT_D;
#pragma unroll
        for (int i = 0; i < AT_D; i += 4) {
          const float4 v = make_float4((o_reg[i] * w0 + r[2 + i] * w1) * inv, (o_reg[i + 1] * w0 + r[3 + i] * w1) * inv,
                                       (o_reg[i + 2] * w0 + r[4 + i] * w1) * inv, (o_reg[i + 3] * w0 + r[5 + i] * w1) * inv);
          if (p.out_f32) st_f4(p.out_f32 + off + i, v);
          if (p.out_hi) {
            uint2 h, l;
            split4(v, h, l);
            *reinterpret_cast<uint2*>(p.out_hi + off + i) = h;
            *reinterpret_cast<uint2*>(p.out_lo + off + i) = l;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// qkv plane [B][T][3C] bf16 -> 3-D map (3C, T, B), box (64, rows, 1), SWIZZLE_128B
static int make_qkv_map(CUtensorMap* m, const void* ptr, int B, int T, int C3, int rows) {
  EncodeTiledFn enc = get_encode();
  BBDM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)C3, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C3 * 2, (cuuint64_t)T * C3 * 2};
  cuuint32_t box[3] = {(cuuint32_t)AT_D, (cuuint32_t)rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BBDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(qkv) failed: %d", (int)r);
  return BBDM_OK;
}

}  // namespace bbdm

using namespace bbdm;

// head_dim 64 only (the template UNets); other head dims are served by bbdm_attention_split.
extern "C" int bbdm_attention_tc(const void* qkv_hi, const void* qkv_lo, int B, int T, int C, int heads, int order,
                                 float* out_f32, void* out_hi, void* out_lo, void* stream) {
  BBDM_REQUIRE(qkv_hi && qkv_lo && (out_f32 || (out_hi && out_lo)), "attention_tc: null pointer");
  BBDM_REQUIRE((out_hi == nullptr) == (out_lo == nullptr), "attention_tc: hi/lo must come in pairs");
  BBDM_REQUIRE(B > 0 && T > 0 && heads > 0 && C % heads == 0 && (order == 0 || order == 1), "attention_tc: bad shape");
  if (C / heads != AT_D) {
    set_error("attention_tc: head_dim %d not supported (64 only)", C / heads);
    return BBDM_E_UNSUPPORTED;
  }
  BBDM_REQUIRE((int64_t)B * heads <= 65535, "attention_tc: B*heads too large");
  CUtensorMap maps[4];
  int rc;
  if ((rc = make_qkv_map(&maps[0], qkv_hi, B, T, 3 * C, AT_BQ))) return rc;
  if ((rc = make_qkv_map(&maps[1], qkv_lo, B, T, 3 * C, AT_BQ))) return rc;
  if ((rc = make_qkv_map(&maps[2], qkv_hi, B, T, 3 * C, AT_BK))) return rc;
  if ((rc = make_qkv_map(&maps[3], qkv_lo, B, T, 3 * C, AT_BK))) return rc;
  AttnParams p;
  p.T = T; p.C = C; p.heads = heads; p.order = order;
  p.scale_log2 = (float)(1.4426950408889634 / sqrt((double)AT_D));
  p.out_f32 = out_f32; p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo;
  p.fault = device_fault_ptr();
  BBDM_REQUIRE(p.fault != nullptr, "attention_tc: device fault word unavailable");
  static DeviceOnce configured;
  if (configured.need()) {
    BBDM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
    configured.mark();
  }
  dim3 grid((T + AT_BQ - 1) / AT_BQ, B * heads);
  attention_tc_kernel<<<grid, 320, AT_SMEM, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3], p);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}
