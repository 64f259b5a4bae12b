// Weight gradient of the stride-1 "same" convolution on tcgen05 tensor cores (sm_100a).
//
//   dW[tap][co][ci] = sum over pixels p of  dY[p][co] * A[p + tap][ci]
//
// GEMM view per filter tap:  M = Cout (tile 128), N = Cin (tile BN), K = pixels (blocks of 64).
//   * A operand  = dY^T planes [Cout][P] (K-major: 64 consecutive pixels = one 128-byte row), written by
//     bbdm_split_grad; 2-D TMA map, SWIZZLE_128B.
//   * B operand  = the forward conv's input planes [B,H,W,Cin]: the same shifted 64-pixel box the
//     forward kernel loads (4-D TMA map, OOB zero fill = padding), consumed as an MN-major operand
//     (channels contiguous); BN/64 swizzle atoms side by side, LBO = one atom (8 KiB).
//   * split-bf16 x3 products, chunked TMEM -> fp32-register promotion (as conv_umma.cu).
//   * K is split across CTAs (the pixel range); every CTA writes its partial [split][tap][Cout][Cin]
//     tile, bbdm's reduce kernel sums the splits in a fixed order (deterministic) into OIHW.
#include "tc_common.cuh"

namespace bbdm {

constexpr int WG_BM = 128;   // couts per tile
constexpr int WG_BK = 64;    // pixels per K block

struct WgradParams {
  int Cout, Cin, taps;
  int n_co, n_ci;            // tiles along Cout / Cin
  int TW, TH, TB, tiles_w, tiles_h, tiles_b;   // 64-pixel box geometry of a K block
  int kblocks;               // total K blocks (= pixel boxes)
  int splits, kb_per_split;
  int kb_per_chunk;
  float* partial;            // [splits][taps][Cout][Cin]
  unsigned long long* fault;
};

// MN-major SWIZZLE_128B descriptor: rows (K index = pixel) of 128 B, 8-row groups 1024 B apart (SBO),
// 64-element MN atoms `lbo` bytes apart.
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
struct WgCfg {
  static constexpr int EPI_WARPS = BN == 256 ? 8 : 4;      // two warps share a TMEM lane quarter for BN = 256
  static constexpr int COLS = BN / (EPI_WARPS / 4);
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

template <int BN>
__global__ void __launch_bounds__(WgCfg<BN>::THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                  const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const WgradParams p) {
  constexpr uint32_t G_BYTES = WG_BM * WG_BK * 2;       // 16 KiB per plane (dY^T tile)
  constexpr uint32_t A_ATOM = WG_BK * 64 * 2;           // 8 KiB: 64 pixels x 64 channels
  constexpr uint32_t A_BYTES = (BN / 64) * A_ATOM;      // per plane
  constexpr uint32_t STAGE_BYTES = 2 * G_BYTES + 2 * A_BYTES;
  constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static_assert(STAGES >= 2, "stage too large");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * 8 + 4];
  __shared__ uint32_t tmem_base_s;
  __shared__ int abort_s;
  const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[8]);
  const uint32_t bar_tfull = smem_u32(&bars[16]), bar_tempty = smem_u32(&bars[18]);
  volatile int* abort_flag = &abort_s;

  if (threadIdx.x == 0) {
    abort_s = 0;
    for (int i = 0; i < STAGES; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, WgCfg<BN>::EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // work item = (split, tap, co tile, ci tile); persistent over the grid
  const int total = p.splits * p.taps * p.n_co * p.n_ci;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        int r = item;
        const int ci_t = r % p.n_ci; r /= p.n_ci;
        const int co_t = r % p.n_co; r /= p.n_co;
        const int tap = r % p.taps;
        const int split = r / p.taps;
        int dy = 0, dx = 0;
        if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        const int kb0 = split * p.kb_per_split;
        const int kb1 = kb0 + p.kb_per_split < p.kblocks ? kb0 + p.kb_per_split : p.kblocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          int mt = kb;
          const int tw = mt % p.tiles_w; mt /= p.tiles_w;
          const int th = mt % p.tiles_h;
          const int tb = mt / p.tiles_h;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1, abort_flag, p.fault, 0xC1000000ull | (unsigned)kb);
          const uint32_t sb = tiles_base + stage * STAGE_BYTES, full = bar_full + 8 * stage;
          mbar_expect_tx(full, STAGE_BYTES);
          // dY^T tile: rows = 128 couts, 64 consecutive pixels (flattened index kb*64)
          tma_load_2d(sb, &map_g_hi, full, kb * WG_BK, co_t * WG_BM);
          tma_load_2d(sb + G_BYTES, &map_g_lo, full, kb * WG_BK, co_t * WG_BM);
#pragma unroll
          for (int at = 0; at < BN / 64; ++at) {
            const int c0 = ci_t * BN + at * 64;
            tma_load_4d(sb + 2 * G_BYTES + at * A_ATOM, &map_a_hi, full, c0, tw * p.TW + dx, th * p.TH + dy, tb * p.TB);
            tma_load_4d(sb + 2 * G_BYTES + A_BYTES + at * A_ATOM, &map_a_lo, full, c0, tw * p.TW + dx, th * p.TH + dy, tb * p.TB);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // warp-uniform MMA issue (see conv_umma.cu)
    const bool leader = elect_one_sync();
    {
      // D=f32, A=B=bf16, A K-major, B MN-major (bit 16), N=BN, M=128
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                                 ((uint32_t)(WG_BM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int split = item / (p.taps * p.n_co * p.n_ci);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = kb0 + p.kb_per_split < p.kblocks ? kb0 + p.kb_per_split : p.kblocks;
        for (int c0 = kb0; c0 < kb1; c0 += p.kb_per_chunk) {
          const int c1 = c0 + p.kb_per_chunk < kb1 ? c0 + p.kb_per_chunk : kb1;
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1, abort_flag, p.fault, 0xC2000000ull | (unsigned)item);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kb = c0; kb < c1; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase, abort_flag, p.fault, 0xC3000000ull | (unsigned)kb);
            tc_fence_after();
            const uint32_t sb = tiles_base + stage * STAGE_BYTES;
            const uint64_t dg_hi = make_sw128_desc(sb), dg_lo = make_sw128_desc(sb + G_BYTES);
            if (leader) {
#pragma unroll
              for (int k = 0; k < WG_BK / 16; ++k) {
                const uint64_t ka = (uint64_t)(k * 32 >> 4);                   // A: +32 B per 16 pixels (K-major)
                const uint32_t ab = sb + 2 * G_BYTES + k * 16 * 128;            // B: +16 pixel rows of 128 B
                const uint64_t da_hi = make_mn_desc(ab, A_ATOM), da_lo = make_mn_desc(ab + A_BYTES, A_ATOM);
                const uint32_t first = (kb > c0 || k > 0) ? 1u : 0u;
                tc_mma_bf16(d_tmem, dg_lo + ka, da_hi, IDESC, first);
                tc_mma_bf16(d_tmem, dg_hi + ka, da_lo, IDESC, 1u);
                tc_mma_bf16(d_tmem, dg_hi + ka, da_hi, IDESC, 1u);
              }
              tc_commit(bar_empty + 8 * stage);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (leader) tc_commit(bar_tfull + 8 * acc);
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // promotion + store: thread = one cout row, COLS cin columns
    constexpr int COLS = WgCfg<BN>::COLS;
    const int q = warp & 3;
    const int col0 = ((warp - 2) >> 2) * COLS;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      int r = item;
      const int ci_t = r % p.n_ci; r /= p.n_ci;
      const int co_t = r % p.n_co; r /= p.n_co;
      const int tap = r % p.taps;
      const int split = r / p.taps;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = kb0 + p.kb_per_split < p.kblocks ? kb0 + p.kb_per_split : p.kblocks;
      float racc[COLS];
#pragma unroll
      for (int j = 0; j < COLS; ++j) racc[j] = 0.f;
      for (int c0 = kb0; c0 < kb1; c0 += p.kb_per_chunk) {
        mbar_wait(bar_tfull + 8 * acc, acc_phase, abort_flag, p.fault, 0xC4000000ull | (unsigned)item);
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
      const int co = co_t * WG_BM + row;
      if (co < p.Cout) {
        float* op = p.partial + (((int64_t)split * p.taps + tap) * p.Cout + co) * p.Cin + ci_t * BN + col0;
#pragma unroll
        for (int j = 0; j < COLS; j += 4) st_f4(op + j, make_float4(racc[j], racc[j + 1], racc[j + 2], racc[j + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// sum the split-K partials in a fixed order; write OIHW:  dW[co][ci][tap].
// thread = one (co, ci) pair, all taps: partial reads coalesced along ci, each thread writes its
// `taps` consecutive output floats.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int Cout, int Cin, float* __restrict__ dw) {
  const int64_t pairs = (int64_t)Cout * Cin;
  const int64_t n = pairs * taps;
  for (int64_t pr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pr < pairs; pr += (int64_t)gridDim.x * blockDim.x) {
    for (int t = 0; t < taps; ++t) {
      float s = 0.f;
      for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * n + (int64_t)t * pairs + pr];
      dw[pr * taps + t] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 NHWC gradient [P][C] -> split planes in both orientations + per-channel sums (bias grad):
//   hi/lo   [P][C]  (K = channel major: A operand of the data-gradient conv)
//   hi_t/lo_t [C][P] (K = pixel major: A operand of the weight-gradient GEMM)
// 32x32 tiles through shared memory; colsum partials per tile row-block, reduced in fixed order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_grad_kernel(const float* __restrict__ src, int64_t P, int C, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ hi_t, __nv_bfloat16* __restrict__ lo_t,
                  float* __restrict__ colsum_part) {
  __shared__ float tile[64][65];
  const int64_t p0 = (int64_t)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x % 64, ty = threadIdx.x / 64;   // 64 x 4
  for (int r = ty; r < 64; r += 4) {
    const int64_t pp = p0 + r;
    float v = 0.f;
    if (pp < P && c0 + tx < C) v = src[pp * C + c0 + tx];
    tile[r][tx] = v;
    if (pp < P && c0 + tx < C && hi) {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      hi[pp * C + c0 + tx] = h;
      lo[pp * C + c0 + tx] = l;
    }
  }
  __syncthreads();
  // transposed planes: thread tx walks pixels (contiguous in the [C][P] layout)
  for (int r = ty; r < 64; r += 4) {
    const int c = c0 + r;
    const int64_t pp = p0 + tx;
    if (c < C && pp < P) {
      __nv_bfloat16 h, l;
      split_bf16(tile[tx][r], h, l);
      hi_t[(int64_t)c * P + pp] = h;
      lo_t[(int64_t)c * P + pp] = l;
    }
  }
  if (colsum_part && ty == 0 && c0 + tx < C) {
    float s = 0.f;
    for (int r = 0; r < 64; ++r) s += tile[r][tx];
    colsum_part[(int64_t)blockIdx.x * C + c0 + tx] = s;
  }
}

// one CTA per 32 channels: 8 row-lanes x 32 channels, fixed-order fp64 tree => deterministic
__global__ void __launch_bounds__(256)
colsum_reduce_kernel(const float* __restrict__ part, int64_t nblk, int C, float* __restrict__ out) {
  __shared__ double sm[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double s = 0.0;
  if (c < C)
    for (int64_t b = ry; b < nblk; b += 8) s += (double)part[b * C + c];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    double t = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += sm[r][cx];
    out[c] = (float)t;
  }
}

static int make_gt_map(CUtensorMap* m, const void* ptr, int Cout, int64_t P) {
  EncodeTiledFn enc = get_encode();
  BBDM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)P * 2};
  cuuint32_t box[2] = {(cuuint32_t)WG_BK, (cuuint32_t)WG_BM};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BBDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(dY^T) failed: %d", (int)r);
  return BBDM_OK;
}

static int make_act64_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int TW, int TH, int TB) {
  EncodeTiledFn enc = get_encode();
  BBDM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BBDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad activation) failed: %d", (int)r);
  return BBDM_OK;
}

template <int BN>
static int launch_wgrad(const CUtensorMap* maps, const WgradParams& p, int grid, cudaStream_t s) {
  constexpr uint32_t STAGE_BYTES = 2 * (WG_BM * WG_BK * 2) + 2 * (BN / 64) * (WG_BK * 64 * 2);
  constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
  static DeviceOnce configured;
  if (configured.need()) {
    BBDM_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.mark();
  }
  conv_wgrad_kernel<BN><<<grid, WgCfg<BN>::THREADS, smem, s>>>(maps[0], maps[1], maps[2], maps[3], p);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // namespace bbdm

using namespace bbdm;

extern "C" {

int bbdm_split_grad(const float* src, int64_t P, int C, void* hi, void* lo, void* hi_t, void* lo_t,
                    float* colsum, float* workspace, void* stream) {
  BBDM_REQUIRE(src && hi_t && lo_t && P > 0 && C > 0, "split_grad: bad args");
  BBDM_REQUIRE((hi == nullptr) == (lo == nullptr), "split_grad: hi/lo in pairs");
  BBDM_REQUIRE(colsum == nullptr || workspace != nullptr, "split_grad: colsum needs a workspace of ceil(P/64)*C floats");
  const int64_t nb = (P + 63) / 64;
  BBDM_REQUIRE(nb < (1ll << 31), "split_grad: too many pixels");
  dim3 grid((unsigned)nb, (C + 63) / 64);
  cudaStream_t s = (cudaStream_t)stream;
  split_grad_kernel<<<grid, 256, 0, s>>>(src, P, C, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, (__nv_bfloat16*)hi_t,
                                         (__nv_bfloat16*)lo_t, colsum ? workspace : nullptr);
  BBDM_LAUNCH_CHECK();
  if (colsum) {
    colsum_reduce_kernel<<<(C + 31) / 32, 256, 0, s>>>(workspace, nb, C, colsum);
    BBDM_LAUNCH_CHECK();
  }
  return BBDM_OK;
}

int bbdm_conv_wgrad_workspace(int B, int H, int W, int Cin, int Cout, int taps, int* splits, int64_t* floats) {
  BBDM_REQUIRE(B > 0 && H > 0 && W > 0 && Cin % 64 == 0 && Cout % 64 == 0 && (taps == 1 || taps == 9), "wgrad_workspace: bad shape");
  const int64_t P = (int64_t)B * H * W;
  const int64_t kblocks = (P + 63) / 64;
  const int BN = Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64);
  const int64_t tiles = (int64_t)taps * ((Cout + 127) / 128) * (Cin / BN);
  int64_t sp = (3 * (int64_t)num_sms() + tiles - 1) / tiles;
  if (sp > kblocks / 8) sp = kblocks / 8;
  if (sp < 1) sp = 1;
  if (sp > 64) sp = 64;
  if (splits) *splits = (int)sp;
  if (floats) *floats = sp * taps * (int64_t)Cout * Cin;
  return BBDM_OK;
}

int bbdm_conv_wgrad(const void* g_hi_t, const void* g_lo_t, const void* a_hi, const void* a_lo, int B, int H,
                    int W, int Cin, int Cout, int taps, float* dw, float* workspace, void* stream) {
  BBDM_REQUIRE(g_hi_t && g_lo_t && a_hi && a_lo && dw && workspace, "conv_wgrad: null pointer");
  BBDM_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0 && (taps == 1 || taps == 9) && W >= 4, "conv_wgrad: unsupported shape");
  const int64_t P = (int64_t)B * H * W;
  BBDM_REQUIRE(P % 64 == 0 && (P * 2) % 16 == 0, "conv_wgrad: B*H*W must be a multiple of 64");
  WgradParams p;
  p.Cout = Cout; p.Cin = Cin; p.taps = taps;
  // 64-pixel box: as wide as possible in w, then h, then b (same order as the flattened pixel index)
  int tw = 1; while (tw * 2 <= W && tw * 2 <= 64 && W % (tw * 2) == 0) tw *= 2;
  int th = 1; while (tw * th * 2 <= 64 && th * 2 <= H && H % (th * 2) == 0) th *= 2;
  int tb = 64 / (tw * th);
  BBDM_REQUIRE(W % tw == 0 && H % th == 0 && (tw == W || th == 1) && (th == H || tb == 1) && B % tb == 0,
               "conv_wgrad: a 64-pixel run must be a box (W=%d H=%d B=%d)", W, H, B);
  p.TW = tw; p.TH = th; p.TB = tb;
  p.tiles_w = W / tw; p.tiles_h = H / th; p.tiles_b = B / tb;
  p.kblocks = (int)(P / 64);
  int splits; int64_t fl;
  int rc = bbdm_conv_wgrad_workspace(B, H, W, Cin, Cout, taps, &splits, &fl);
  if (rc) return rc;
  p.splits = splits;
  p.kb_per_split = (p.kblocks + splits - 1) / splits;
  p.splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;      // no empty splits
  p.kb_per_chunk = 4;
  p.partial = workspace;
  p.fault = device_fault_ptr();
  BBDM_REQUIRE(p.fault != nullptr, "conv_wgrad: device fault word unavailable");
  const int BN = Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64);
  p.n_co = (Cout + WG_BM - 1) / WG_BM;
  p.n_ci = Cin / BN;
  CUtensorMap maps[4];
  if ((rc = make_gt_map(&maps[0], g_hi_t, Cout, P))) return rc;
  if ((rc = make_gt_map(&maps[1], g_lo_t, Cout, P))) return rc;
  if ((rc = make_act64_map(&maps[2], a_hi, B, H, W, Cin, tw, th, tb))) return rc;
  if ((rc = make_act64_map(&maps[3], a_lo, B, H, W, Cin, tw, th, tb))) return rc;
  const int64_t total = (int64_t)p.splits * taps * p.n_co * p.n_ci;
  const int grid = (int)(total < num_sms() ? total : num_sms());
  cudaStream_t s = (cudaStream_t)stream;
  if (BN == 256) rc = launch_wgrad<256>(maps, p, grid, s);
  else if (BN == 128) rc = launch_wgrad<128>(maps, p, grid, s);
  else rc = launch_wgrad<64>(maps, p, grid, s);
  if (rc) return rc;
  const int64_t n = (int64_t)Cout * Cin;
  int64_t g = (n + 255) / 256;
  if (g > (int64_t)num_sms() * 8) g = (int64_t)num_sms() * 8;
  wgrad_reduce_kernel<<<(int)g, 256, 0, s>>>(workspace, p.splits, taps, Cout, Cin, dw);
  BBDM_LAUNCH_CHECK();
  return BBDM_OK;
}

}  // extern "C"
