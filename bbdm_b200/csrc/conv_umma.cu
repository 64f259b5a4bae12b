// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[pixel, co] = sum_{tap, ci} A[pixel + tap, ci] * W[tap, co, ci]  (+ fused 1x1 operand)
//                    + bias (+ bias2) (+ residual)
//
//   * M tile = 128 output pixels = a (TW x TH x TB) box of the NHWC tensor; N tile = BN couts;
//     K block = 64 input channels of one filter tap.
//   * Operands are split-bf16 planes (hi, lo).  passes=3 issues A_hi.W_hi + A_lo.W_hi + A_hi.W_lo
//     into the same TMEM accumulator (fp32-class accuracy); passes=1 issues A_hi.W_hi only.
//   * TMA (cp.async.bulk.tensor, 4-D tiled map, SWIZZLE_128B) stages the shifted pixel box of a
//     tap straight from the activation tensor; out-of-image coordinates are zero-filled by the
//     TMA unit, which IS the conv padding -- no im2col buffer, no halo logic.
//   * Warp-specialised persistent CTA: warp0 = TMA producer, warp1 = tcgen05.mma issuer (one
//     elected lane) + TMEM owner, warps2-5 = epilogue (TMEM -> registers -> bias/residual ->
//     global).  Two TMEM accumulator buffers: chunk promotion and the tile epilogue overlap the
//     tensor-core mainloop.
//   * fp32-faithful accumulation: the tensor core's accumulator add truncates (round-toward-zero),
//     so a K = 9216 chain accumulated entirely in TMEM drifts by ~3e-5.  The K loop is therefore
//     cut into chunks of `kb_per_chunk` K-blocks that alternate between the two TMEM buffers; the
//     epilogue warps drain each finished chunk (tcgen05.ld) and add it into fp32 REGISTER
//     accumulators with round-to-nearest while the tensor core works on the next chunk.
//   * Every mbarrier wait has a clock-based watchdog: on expiry a device fault word is set and
//     the CTA drains without deadlocking the GPU (bbdm_check_device_fault reports it).
#include "tc_common.cuh"
#include <stdlib.h>

namespace bbdm {

constexpr int UM_BM = 128;       // pixels per tile (UMMA M)
// K block = BK bf16 input channels of one tap: 64 (128-byte rows, SWIZZLE_128B, default) or 32
// (64-byte rows, SWIZZLE_64B: half-size stages, 4-deep TMA ring for the 256-wide tiles).

// Column sums over the 32 lanes (rows) of a warp for 32 per-lane values: reduce-scatter
// butterfly, 31 shuffles; afterwards lane l holds the total of column l in v[0].
__device__ __forceinline__ void warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const bool up = (lane & off) != 0;
      const float keep = up ? v[i + off] : v[i];
      const float send = up ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

struct ConvParams {
  int B, H, W, Cout;
  int TW, TH, TB, tiles_w, tiles_h, tiles_b, n_tiles;
  int kb_per_tap;   // Cin / BK
  int K1;           // taps * Cin/64
  int K2;           // Cin2 / BK
  int taps;
  int up2;           // 1: fused nearest-2x upsample (4 output phases x 2x2 taps on the low-res input)
  int passes;
  int w_per_image;   // 1: the weight "tap" index is the image index of the tile (36 Winograd position GEMMs in one launch)
  int f16;           // 1: operand planes are fp16 (hi, lo) instead of bf16
  int kb_per_chunk;  // K blocks accumulated in TMEM before promotion to registers
  const float* bias; const float* bias2;
  const float* residual; int res_mode;
  float* out; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int out_nchw_c;    // > 0: out is NCHW with this many channels (UNet head)
  float* stats;      // fused GroupNorm partial sums (see BbdmConvArgs.stats_partial) or nullptr
  unsigned long long* fault;
};

template <int BN, int BK = 64>
struct UmmaCfg {
  static constexpr uint32_t A_BYTES = UM_BM * BK * 2;     // per plane per stage
  static constexpr uint32_t W_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE3 = 2 * A_BYTES + 2 * W_BYTES;  // hi+lo planes
  static constexpr uint32_t STAGE1 = A_BYTES + W_BYTES;
  static constexpr uint32_t SMEM_BUDGET = 200 * 1024;
  static constexpr int STAGES3 = (SMEM_BUDGET / STAGE3) > 6 ? 6 : (SMEM_BUDGET / STAGE3);
  static constexpr int STAGES1 = (SMEM_BUDGET / STAGE1) > 8 ? 8 : (SMEM_BUDGET / STAGE1);
  static constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // power of two for BN in {64,128,256}
  // epilogue warps: 4 (one per TMEM lane quarter) own BN columns each; for BN = 256 two warps
  // share a lane quarter (128 columns = 128 accumulator registers per thread)
  static constexpr int EPI_WARPS = BN == 256 ? 8 : 4;
  static constexpr int COLS = BN / (EPI_WARPS / 4);
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

// shared-memory matrix descriptor for a K-major tile whose rows are BK bf16 wide
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
  if (BK == 64) return make_sw128_desc(saddr);
  // SWIZZLE_64B: 64-byte rows, 8-row groups 512 B apart, layout type 4
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

template <int BN, int PASSES, int BK>
__global__ void __launch_bounds__(UmmaCfg<BN>::THREADS, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                 const __grid_constant__ CUtensorMap map_a2_hi, const __grid_constant__ CUtensorMap map_a2_lo,
                 const __grid_constant__ CUtensorMap map_w2_hi, const __grid_constant__ CUtensorMap map_w2_lo,
                 const ConvParams p) {
  using Cfg = UmmaCfg<BN, BK>;
  constexpr int STAGES = PASSES == 3 ? Cfg::STAGES3 : Cfg::STAGES1;
  constexpr uint32_t STAGE_BYTES = PASSES == 3 ? Cfg::STAGE3 : Cfg::STAGE1;
  constexpr uint32_t OFF_ALO = Cfg::A_BYTES;
  constexpr uint32_t OFF_WHI = PASSES == 3 ? 2 * Cfg::A_BYTES : Cfg::A_BYTES;
  constexpr uint32_t OFF_WLO = OFF_WHI + Cfg::W_BYTES;
  static_assert(STAGES >= 2, "need at least a double buffer");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * 8 + 4];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_s;
  // per-epilogue-warp staging tile (32 rows x 16 columns, row stride 20 floats) for the coalesced plain store
  __shared__ __align__(16) float stg_s[Cfg::EPI_WARPS][32 * 20];

  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]);        // [STAGES]
  const uint32_t bar_empty = smem_u32(&bars[8]);       // [STAGES]
  const uint32_t bar_tfull = smem_u32(&bars[16]);      // [2]
  const uint32_t bar_tempty = smem_u32(&bars[18]);     // [2]
  volatile int* abort_flag = &abort_s;

  if (threadIdx.x == 0) {
    abort_s = 0;
    for (int i = 0; i < STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, Cfg::EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int n_blocks = p.Cout / BN;
  const int total_tiles = p.n_tiles * n_blocks * (p.up2 ? 4 : 1);
  const int KB = p.K1 + p.K2;

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase_bit = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nb = tile % n_blocks;
        int mt = tile / n_blocks;
        int phase = 0;
        if (p.up2) { phase = mt & 3; mt >>= 2; }
        const int tw = mt % p.tiles_w; mt /= p.tiles_w;
        const int th = mt % p.tiles_h;
        const int tb = mt / p.tiles_h;
        const int w0 = tw * p.TW, h0 = th * p.TH, b0 = tb * p.TB, n0 = nb * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase_bit ^ 1, abort_flag, p.fault, 0xE0000000ull | (unsigned)kb);
          const uint32_t sbase = tiles_base + stage * STAGE_BYTES;
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, STAGE_BYTES);
          if (kb < p.K1) {
            const int tap = kb / p.kb_per_tap, cb = kb - tap * p.kb_per_tap;
            int dy = 0, dx = 0, wtap = tap;
            if (p.up2) {
              // output phase (a, b) = (phase>>1, phase&1); 2x2 taps (r, c) on the low-res source
              const int r = tap >> 1, c = tap & 1;
              dy = (phase >> 1) ? r : r - 1;
              dx = (phase & 1) ? c : c - 1;
              wtap = phase * 4 + tap;
            } else if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
            else if (p.taps == 4) { dy = tap >> 1; dx = tap & 1; }   // 2x2 window at (0..1, 0..1): zero pad bottom/right
            else if (p.w_per_image) wtap = tb;                       // taps == 1: weights of transform position tb
            tma_load_4d(sbase, &map_a_hi, full, cb * BK, w0 + dx, h0 + dy, b0);
            if (PASSES == 3) tma_load_4d(sbase + OFF_ALO, &map_a_lo, full, cb * BK, w0 + dx, h0 + dy, b0);
            tma_load_3d(sbase + OFF_WHI, &map_w_hi, full, cb * BK, n0, wtap);
            if (PASSES == 3) tma_load_3d(sbase + OFF_WLO, &map_w_lo, full, cb * BK, n0, wtap);
          } else {
            const int cb = kb - p.K1;
            tma_load_4d(sbase, &map_a2_hi, full, cb * BK, w0, h0, b0);
            if (PASSES == 3) tma_load_4d(sbase + OFF_ALO, &map_a2_lo, full, cb * BK, w0, h0, b0);
            tma_load_3d(sbase + OFF_WHI, &map_w2_hi, full, cb * BK, n0, 0);
            if (PASSES == 3) tma_load_3d(sbase + OFF_WLO, &map_w2_lo, full, cb * BK, n0, 0);
          }
          if (++stage == STAGES) { stage = 0; phase_bit ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    // Warp-uniform issue: the whole warp runs the loop in uniform control flow and only tcgen05.mma / commit sit
    // under the elected lane's predicate, so ptxas keeps the descriptors in uniform registers (12 MMAs of a K-block:
    // 30 straight-line instructions instead of 91 with an ELECT/BRA loop around each UTCHMMA).  Measured round 2
    // (profiles/r02_uniform_issue_ab.md): cfg2 163.1 -> 160.2 ms, cfg1 4.64 -> 4.32 ms, cfg5 7.32 -> 7.14 ms.
    const bool leader = elect_one_sync();
    {
      // instruction descriptor: D=f32, A=B=bf16 (format 1) or fp16 (format 0), K-major both, N>>3 @17, M>>4 @24
      const uint32_t IDESC = (1u << 4) | (p.f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(UM_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int kb0 = 0; kb0 < KB; kb0 += p.kb_per_chunk) {
          const int kb1 = kb0 + p.kb_per_chunk < KB ? kb0 + p.kb_per_chunk : KB;
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1, abort_flag, p.fault, 0xA0000000ull | (unsigned)tile);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase, abort_flag, p.fault, 0xF0000000ull | (unsigned)kb);
            tc_fence_after();
            const uint32_t sbase = tiles_base + stage * STAGE_BYTES;
            const uint64_t da_hi = make_kmajor_desc<BK>(sbase);
            const uint64_t da_lo = make_kmajor_desc<BK>(sbase + OFF_ALO);
            const uint64_t db_hi = make_kmajor_desc<BK>(sbase + OFF_WHI);
            const uint64_t db_lo = make_kmajor_desc<BK>(sbase + OFF_WLO);
            if (leader) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                const uint64_t ko = (uint64_t)(k * 32 >> 4);   // +32 B per UMMA_K inside the swizzle atom
                const uint32_t first = (kb > kb0 || k > 0) ? 1u : 0u;   // a chunk starts from zero
                if (PASSES == 3) {
                  tc_mma_bf16(d_tmem, da_lo + ko, db_hi + ko, IDESC, first);
                  tc_mma_bf16(d_tmem, da_hi + ko, db_lo + ko, IDESC, 1u);
                  tc_mma_bf16(d_tmem, da_hi + ko, db_hi + ko, IDESC, 1u);
                } else {
                  tc_mma_bf16(d_tmem, da_hi + ko, db_hi + ko, IDESC, first);
                }
              }
              tc_commit(bar_empty + 8 * stage);        // frees the smem slot when these MMAs retire
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (leader) tc_commit(bar_tfull + 8 * acc);  // chunk complete -> promotion warps
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ promotion + epilogue warps ============================
    constexpr int COLS = Cfg::COLS;            // accumulator columns (= registers) per thread
    const int ew = warp - 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int col0 = (ew >> 2) * COLS;         // column half for BN = 256
    const int row = q * 32 + lane;             // tile row == TMEM lane
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nb = tile % n_blocks;
      int mt = tile / n_blocks;
      int phase = 0;
      if (p.up2) { phase = mt & 3; mt >>= 2; }
      const int tw = mt % p.tiles_w; mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int tb = mt / p.tiles_h;
      const int ws = tw * p.TW + row % p.TW;            // coordinates in the conv INPUT grid
      const int hs = th * p.TH + (row / p.TW) % p.TH;
      const int bb = tb * p.TB + row / (p.TW * p.TH);
      const bool valid = ws < p.W && hs < p.H && bb < p.B;
      // output grid: same as the input, or 2x with this tile's phase offset (fused upsample)
      const int OH = p.up2 ? 2 * p.H : p.H, OW = p.up2 ? 2 * p.W : p.W;
      const int hh = p.up2 ? 2 * hs + (phase >> 1) : hs;
      const int ww = p.up2 ? 2 * ws + (phase & 1) : ws;
      const int64_t pix = ((int64_t)bb * OH + hh) * OW + ww;
      const int n0 = nb * BN + col0;

      // plain GEMM-style output (no bias / residual / statistics / split copy: the Winograd position GEMMs): the tile
      // is stored through a shared-memory transpose, 8 rows x 64 contiguous bytes per instruction, instead of one
      // 16-byte piece of 32 different rows (half-used sectors; for K = 512 tiles the store took longer than the two
      // chunks the tensor core may run ahead)
      const bool plain = p.out != nullptr && p.bias == nullptr && p.bias2 == nullptr && p.res_mode == BBDM_RES_NONE &&
                         p.out_hi == nullptr && p.stats == nullptr && p.out_nchw_c == 0;
      int64_t prow[4];
      bool pval[4];
      if (plain) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int src = 8 * i + (lane >> 2);
          prow[i] = __shfl_sync(0xffffffffu, pix, src);
          pval[i] = __shfl_sync(0xffffffffu, valid ? 1 : 0, src) != 0;
        }
      }
      float racc[COLS];
#pragma unroll
      for (int j = 0; j < COLS; ++j) racc[j] = 0.f;
      // ---- drain every K chunk: TMEM -> registers, round-to-nearest fp32 adds -------------------
      for (int kb0 = 0; kb0 < KB; kb0 += p.kb_per_chunk) {
        mbar_wait(bar_tfull + 8 * acc, acc_phase, abort_flag, p.fault, 0xD0000000ull | (unsigned)tile);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + col0;
#pragma unroll
        for (int ch = 0; ch < COLS / 32; ++ch) {
          uint32_t v[32];
          tc_ld32(t_addr + ch * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) racc[ch * 32 + j] += __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // ---- tile epilogue: bias / fused-skip bias / residual / store ------------------------------
      {
#pragma unroll
        for (int ch = 0; ch < COLS / 32; ++ch) {
          const int nc = n0 + ch * 32;
          float* r = &racc[ch * 32];
          if (plain) {
            float* st = stg_s[ew];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(st + lane * 20 + 4 * j) =
                    make_float4(r[16 * h + 4 * j], r[16 * h + 4 * j + 1], r[16 * h + 4 * j + 2], r[16 * h + 4 * j + 3]);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(st + (8 * i + (lane >> 2)) * 20 + (lane & 3) * 4);
                if (pval[i]) st_f4(p.out + prow[i] * p.Cout + nc + 16 * h + (lane & 3) * 4, v);
              }
              __syncwarp();
            }
            continue;
          }
          if (valid) {
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = ld_f4(p.bias + nc + j);
              r[j] += bv.x; r[j + 1] += bv.y; r[j + 2] += bv.z; r[j + 3] += bv.w;
            }
          }
          if (p.bias2) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = ld_f4(p.bias2 + nc + j);
              r[j] += bv.x; r[j + 1] += bv.y; r[j + 2] += bv.z; r[j + 3] += bv.w;
            }
          }
          if (p.res_mode == BBDM_RES_SAME) {
            const float* rp = p.residual + pix * p.Cout + nc;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = ld_f4(rp + j);
              r[j] += t.x; r[j + 1] += t.y; r[j + 2] += t.z; r[j + 3] += t.w;
            }
          } else if (p.res_mode == BBDM_RES_UP2) {
            const float* rp = p.residual + (((int64_t)bb * (OH >> 1) + (hh >> 1)) * (OW >> 1) + (ww >> 1)) * p.Cout + nc;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = ld_f4(rp + j);
              r[j] += t.x; r[j + 1] += t.y; r[j + 2] += t.z; r[j + 3] += t.w;
            }
          } else if (p.res_mode == BBDM_RES_DOWN2) {
            const int64_t W2 = (int64_t)OW * 2;
            const float* rp = p.residual + (((int64_t)bb * OH * 2 + hh * 2) * W2 + ww * 2) * p.Cout + nc;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t0 = ld_f4(rp + j), t1 = ld_f4(rp + p.Cout + j);
              const float4 t2 = ld_f4(rp + W2 * p.Cout + j), t3 = ld_f4(rp + (W2 + 1) * p.Cout + j);
              r[j] += 0.25f * (((t0.x + t1.x) + t2.x) + t3.x);
              r[j + 1] += 0.25f * (((t0.y + t1.y) + t2.y) + t3.y);
              r[j + 2] += 0.25f * (((t0.z + t1.z) + t2.z) + t3.z);
              r[j + 3] += 0.25f * (((t0.w + t1.w) + t2.w) + t3.w);
            }
          }
          if (p.out_nchw_c > 0) {
            // head: first out_nchw_c couts straight into the NCHW result (lanes = consecutive w)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nc + j < p.out_nchw_c)
                p.out[(((int64_t)bb * p.out_nchw_c + nc + j) * OH + hh) * OW + ww] = r[j];
          } else if (p.out) {
            float* op = p.out + pix * p.Cout + nc;
#pragma unroll
            for (int j = 0; j < 32; j += 4) st_f4(op + j, make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]));
          }
          if (p.out_hi) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              uint2 h, l;
              split4(make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]), h, l);
              *reinterpret_cast<uint2*>(p.out_hi + pix * p.Cout + nc + j) = h;
              *reinterpret_cast<uint2*>(p.out_lo + pix * p.Cout + nc + j) = l;
            }
          }
          }   // valid
          if (p.stats) {
            // fused GroupNorm statistics of the tensor just written: per-channel (sum, sum sq) over
            // this warp's 32 pixel rows (tile lies inside one image: TB == 1)
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { if (!valid) r[j] = 0.f; sq[j] = r[j] * r[j]; }
            warp_colsum32(r, lane);
            warp_colsum32(sq, lane);
            const int64_t prow = (((int64_t)tb * p.tiles_w * p.tiles_h + th * p.tiles_w + tw) * (p.up2 ? 4 : 1) + phase) * 4 + q;
            *reinterpret_cast<float2*>(p.stats + (prow * p.Cout + nc + lane) * 2) = make_float2(r[0], sq[0]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------- host side
// bf16 NHWC activation [B,H,W,C] -> 4-D map (C, W, H, B), box (64, TW, TH, TB), SWIZZLE_128B
static int make_act_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int TW, int TH, int TB, int BK,
                        bool f16 = false) {
  EncodeTiledFn enc = get_encode();
  BBDM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BBDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d (B=%d H=%d W=%d C=%d box=%dx%dx%d)",
               (int)r, B, H, W, C, TW, TH, TB);
  return BBDM_OK;
}

// bf16 weights [taps][Cout][Cin] -> 3-D map (Cin, Cout, taps), box (64, BN, 1)
static int make_w_map(CUtensorMap* m, const void* ptr, int taps, int Cout, int Cin, int BN, int BK, bool f16 = false) {
  EncodeTiledFn enc = get_encode();
  BBDM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BBDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed: %d (taps=%d Cout=%d Cin=%d BN=%d)", (int)r,
               taps, Cout, Cin, BN);
  return BBDM_OK;
}

static int pow2_floor(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }
static int pow2_ceil(int x) { int p = 1; while (p < x) p *= 2; return p; }

template <int BN, int PASSES, int BK>
static int launch_conv(const CUtensorMap* maps, const ConvParams& p, int grid, cudaStream_t s) {
  using Cfg = UmmaCfg<BN, BK>;
  constexpr int STAGES = PASSES == 3 ? Cfg::STAGES3 : Cfg::STAGES1;
  constexpr uint32_t STAGE_BYTES = PASSES == 3 ? Cfg::STAGE3 : Cfg::STAGE1;
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  static DeviceOnce configured;
  if (configured.need()) {
    BBDM_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<BN, PASSES, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.mark();
  }
  conv_umma_kernel<BN, PASSES, BK><<<grid, Cfg::THREADS, smem, s>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5],
                                                              maps[6], maps[7], p);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // namespace bbdm

using namespace bbdm;

// BBDM_CONV_BK=32 selects the 32-channel K block for the 256-wide tiles (A/B measurements)
static int conv_bk_override() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("BBDM_CONV_BK"); v = e ? atoi(e) : 0; }
  return v;
}

static void tile_geometry(int H, int W, int* TW, int* TH, int* TB) {
  *TW = pow2_floor(W) < 16 ? pow2_floor(W) : 16;
  const int th = pow2_ceil(H);
  *TH = th < UM_BM / *TW ? th : UM_BM / *TW;
  *TB = UM_BM / (*TW * *TH);
}

// (for an upsample2x conv call this with the INPUT H, W and multiply rows_per_image by 4)
extern "C" int bbdm_conv_umma_geometry(int H, int W, int* TW, int* TH, int* TB, int* rows_per_image) {
  BBDM_REQUIRE(H > 0 && W >= 4, "conv_umma_geometry: need H > 0, W >= 4");
  int tw, th, tb;
  tile_geometry(H, W, &tw, &th, &tb);
  if (TW) *TW = tw;
  if (TH) *TH = th;
  if (TB) *TB = tb;
  if (rows_per_image) *rows_per_image = tb == 1 ? 4 * ((W + tw - 1) / tw) * ((H + th - 1) / th) : 0;
  return BBDM_OK;
}

extern "C" int bbdm_conv_umma(const BbdmConvArgs* a, void* stream) {
  BBDM_REQUIRE(a != nullptr, "conv_umma: null args");
  BBDM_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, "conv_umma: bad spatial shape");
  BBDM_REQUIRE(a->taps == 1 || a->taps == 9 || a->taps == 4, "conv_umma: taps must be 1, 4 or 9 (got %d)", a->taps);
  BBDM_REQUIRE(!a->upsample2x || (a->taps == 4 && a->Cin2 == 0), "conv_umma: upsample2x needs the 16 phase taps and no fused 1x1");
  BBDM_REQUIRE(a->Cin > 0 && a->Cin % 64 == 0, "conv_umma: Cin %% 64 != 0 (Cin=%d)", a->Cin);
  BBDM_REQUIRE(a->Cout > 0 && a->Cout % 64 == 0, "conv_umma: Cout %% 64 != 0 (Cout=%d)", a->Cout);
  BBDM_REQUIRE(a->Cin2 >= 0 && a->Cin2 % 64 == 0, "conv_umma: Cin2 %% 64 != 0 (Cin2=%d)", a->Cin2);
  BBDM_REQUIRE(a->passes == 1 || a->passes == 3, "conv_umma: passes must be 1 or 3");
  BBDM_REQUIRE(a->a_hi && a->w_hi && (a->passes == 1 || (a->a_lo && a->w_lo)), "conv_umma: missing operand plane");
  if (a->Cin2) BBDM_REQUIRE(a->a2_hi && a->w2_hi && (a->passes == 1 || (a->a2_lo && a->w2_lo)), "conv_umma: missing 1x1 operand plane");
  BBDM_REQUIRE(a->out || (a->out_hi && a->out_lo), "conv_umma: no output");
  BBDM_REQUIRE((a->out_hi == nullptr) == (a->out_lo == nullptr), "conv_umma: out hi/lo must come in pairs");
  BBDM_REQUIRE(a->res_mode >= 0 && a->res_mode <= 3 && (a->res_mode == 0 || a->residual), "conv_umma: bad residual");
  if (a->res_mode == BBDM_RES_UP2 && !a->upsample2x) BBDM_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0, "conv_umma: RES_UP2 needs even H, W");
  BBDM_REQUIRE(a->W >= 4, "conv_umma: W < 4 not supported (use conv_direct)");
  const bool wpi = a->weights_per_image != 0, f16 = a->operand_f16 != 0;
  if (wpi) BBDM_REQUIRE(a->taps == 1 && a->Cin2 == 0 && !a->upsample2x, "conv_umma: weights_per_image needs taps == 1 and no fused operand");

  ConvParams p;
  p.B = a->B; p.H = a->H; p.W = a->W; p.Cout = a->Cout;
  tile_geometry(a->H, a->W, &p.TW, &p.TH, &p.TB);
  p.tiles_w = (a->W + p.TW - 1) / p.TW;
  p.tiles_h = (a->H + p.TH - 1) / p.TH;
  p.tiles_b = (a->B + p.TB - 1) / p.TB;
  p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  // N tile: the widest that divides Cout -- unless that leaves most SMs without a tile (small spatial
  // extents / batches): then narrower tiles (more CTAs, each less efficient) finish sooner.
  int BN = (a->Cout % 256 == 0) ? 256 : (a->Cout % 128 == 0 ? 128 : 64);
  {
    const int64_t m_tiles = (int64_t)p.n_tiles * (a->upsample2x ? 4 : 1);
    const int64_t want = (int64_t)(0.7 * num_sms());
    while (BN > 64 && m_tiles * (a->Cout / BN) < want) BN /= 2;
  }
  // BBDM_CONV_BK=32: 256-wide N tiles in split mode use 32-channel K blocks (4-stage TMA ring instead of
  // 2).  Measured on B200 (round 1): the tensor pipe gets busier but the 1 kW power cap lowers the SM
  // clock by the same factor (1522 -> 1335 MHz), identical step time -- so 64 stays the default.
  const int BK = (BN == 256 && a->passes == 3 && conv_bk_override() == 32) ? 32 : 64;
  p.kb_per_tap = a->Cin / BK;
  p.K1 = a->taps * p.kb_per_tap;
  p.K2 = a->Cin2 / BK;
  p.taps = a->taps;
  p.up2 = a->upsample2x ? 1 : 0;
  p.passes = a->passes;
  p.w_per_image = wpi ? 1 : 0;
  p.f16 = f16 ? 1 : 0;
  // chunk length: 4 K-blocks (direct conv); 2 for the Winograd position GEMMs, whose output transform amplifies
  // the truncation error of the TMEM accumulator (tools/studies/tmem_rz_accumulation.py)
  // Winograd position GEMMs: the output transform amplifies the truncation error of the TMEM accumulator, so they
  // promote more often than the direct conv -- every 2 K-blocks below 512 input channels, every 4 from 512 on
  // (measured chain deviation from the fp64 conv, tests/test_gpu_winograd.py: C = 256: 7.6e-6 with 2, 1.4e-5 with 4;
  // C = 1024: 3.9e-6 with 2, 6.2e-6 with 4; model fixtures unchanged at 1.9-2.4e-5; draining 128 KB of TMEM per chunk
  // costs ~2000 cycles of tcgen05.ld against 3072 / 6144 cycles of MMAs, so 4 is 10-15 % faster).
  // BBDM_WINO_CHUNK overrides (A/B switch).
  static int wino_chunk = -1;
  if (wino_chunk < 0) { const char* e = getenv("BBDM_WINO_CHUNK"); wino_chunk = (e && atoi(e) > 0) ? atoi(e) : 0; }
  const int wchunk = wino_chunk ? wino_chunk : (a->Cin >= 512 ? 4 : 2);
  p.kb_per_chunk = (wpi ? wchunk : (a->passes == 3 ? 4 : 8)) * (64 / BK);
  if (wpi) BBDM_REQUIRE(p.TB == 1, "conv_umma: weights_per_image needs 128-pixel tiles inside one image (H*W >= 128)");
  p.bias = a->bias; p.bias2 = a->Cin2 ? a->bias2 : nullptr;
  p.residual = a->residual; p.res_mode = a->res_mode;
  p.out = a->out; p.out_hi = (__nv_bfloat16*)a->out_hi; p.out_lo = (__nv_bfloat16*)a->out_lo;
  p.out_nchw_c = a->out_nchw_channels;
  p.stats = a->stats_partial;
  BBDM_REQUIRE(p.out_nchw_c >= 0 && p.out_nchw_c <= a->Cout && (p.out_nchw_c == 0 || (a->out && !a->out_hi)),
               "conv_umma: bad out_nchw_channels");
  p.fault = device_fault_ptr();
  BBDM_REQUIRE(p.fault != nullptr, "conv_umma: device fault word unavailable");

  BBDM_REQUIRE(p.stats == nullptr || (p.TB == 1 && p.out_nchw_c == 0),
               "conv_umma: stats_partial needs a tile inside one image (H*W >= 128) and NHWC output");
  CUtensorMap maps[8];
  int rc;
  if ((rc = make_act_map(&maps[0], a->a_hi, a->B, a->H, a->W, a->Cin, p.TW, p.TH, p.TB, BK, f16))) return rc;
  if ((rc = make_act_map(&maps[1], a->passes == 3 ? a->a_lo : a->a_hi, a->B, a->H, a->W, a->Cin, p.TW, p.TH, p.TB, BK, f16))) return rc;
  const int wtaps = a->upsample2x ? 16 : (wpi ? a->B : a->taps);
  if ((rc = make_w_map(&maps[2], a->w_hi, wtaps, a->Cout, a->Cin, BN, BK, f16))) return rc;
  if ((rc = make_w_map(&maps[3], a->passes == 3 ? a->w_lo : a->w_hi, wtaps, a->Cout, a->Cin, BN, BK, f16))) return rc;
  if (a->Cin2) {
    if ((rc = make_act_map(&maps[4], a->a2_hi, a->B, a->H, a->W, a->Cin2, p.TW, p.TH, p.TB, BK, f16))) return rc;
    if ((rc = make_act_map(&maps[5], a->passes == 3 ? a->a2_lo : a->a2_hi, a->B, a->H, a->W, a->Cin2, p.TW, p.TH, p.TB, BK, f16))) return rc;
    if ((rc = make_w_map(&maps[6], a->w2_hi, 1, a->Cout, a->Cin2, BN, BK, f16))) return rc;
    if ((rc = make_w_map(&maps[7], a->passes == 3 ? a->w2_lo : a->w2_hi, 1, a->Cout, a->Cin2, BN, BK, f16))) return rc;
  } else {
    maps[4] = maps[0]; maps[5] = maps[1]; maps[6] = maps[2]; maps[7] = maps[3];
  }
  const int64_t total = (int64_t)p.n_tiles * (a->Cout / BN) * (p.up2 ? 4 : 1);
  BBDM_REQUIRE(total < (1ll << 30), "conv_umma: too many tiles");
  const int grid = (int)(total < num_sms() ? total : num_sms());
  cudaStream_t s = (cudaStream_t)stream;
#define BBDM_UL(N)                                                        \
  (a->passes == 3 ? launch_conv<N, 3, 64>(maps, p, grid, s) : launch_conv<N, 1, 64>(maps, p, grid, s))
  if (BN == 256 && BK == 32) return launch_conv<256, 3, 32>(maps, p, grid, s);
  if (BN == 256) return BBDM_UL(256);
  if (BN == 128) return BBDM_UL(128);
  return BBDM_UL(64);
#undef BBDM_UL
}
