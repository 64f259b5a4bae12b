// Host-side check of the Winograd output transform's tile routine (bbdm_b200/csrc/winograd.cu,
// wino_output_tile<RES>, a __host__ __device__ function): the SAME source the kernel runs is executed on the CPU for
// every (sample, tile, channel pair) and compared with a direct fp64 evaluation of
//     out = 2^-8 * A^T M A + bias + residual(same | nearest-up | 2x2-average addressed)
// plus the per-thread partial sums that feed the fused GroupNorm statistics.  No GPU and no CUDA runtime call is
// involved.  Build + run (tests/test_wino_output_host.py does this):
//     nvcc -std=c++17 --expt-relaxed-constexpr -I include -o /tmp/host_check_wino_output tools/host_check_wino_output.cu
#include "../bbdm_b200/csrc/winograd.cu"

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

static const double AT[4][6] = {{1, 1, 1, 1, 1, 0}, {0, 1, -1, 2, -2, 0}, {0, 1, 1, 4, 4, 0}, {0, 1, -1, 8, -8, 1}};

template <int RES>
static int run(int B, int H, int W, int Cout, bool with_bias) {
  using namespace bbdm;
  std::mt19937 rng(1234 + RES * 7 + H);
  std::normal_distribution<float> nd(0.f, 1.f);
  const int th = H / 4, tw = W / 4;
  const int64_t Mtot = (int64_t)B * th * tw;
  std::vector<float> M((size_t)36 * Mtot * Cout), bias(Cout), out((size_t)B * H * W * Cout, -777.f);
  for (auto& v : M) v = 40.f * nd(rng);
  for (auto& v : bias) v = nd(rng);
  int RH = H, RW = W;
  if (RES == BBDM_RES_UP2) { RH = H / 2; RW = W / 2; }
  if (RES == BBDM_RES_DOWN2) { RH = H * 2; RW = W * 2; }
  std::vector<float> res((size_t)B * RH * RW * Cout);
  for (auto& v : res) v = nd(rng);
  WinoOutParams p;
  p.m = M.data(); p.Mtot = Mtot; p.B = B; p.H = H; p.W = W; p.Cout = Cout; p.th = th; p.tw = tw;
  p.bias = with_bias ? bias.data() : nullptr;
  p.residual = RES == BBDM_RES_NONE ? nullptr : res.data(); p.res_mode = RES;
  p.out = out.data(); p.stats = nullptr;
  std::vector<double> s1((size_t)Cout, 0.0), s2((size_t)Cout, 0.0);
  for (int b = 0; b < B; ++b)
    for (int ty = 0; ty < th; ++ty)
      for (int tx = 0; tx < tw; ++tx)
        for (int c = 0; c < Cout; c += 2) {
          float a0 = 0, a1 = 0, q0 = 0, q1 = 0;
          const float2 bv = with_bias ? make_float2(bias[c], bias[c + 1]) : make_float2(0.f, 0.f);
          wino_output_tile<RES>(p, b, ty, tx, c, bv, a0, a1, q0, q1);
          s1[c] += a0; s1[c + 1] += a1; s2[c] += q0; s2[c + 1] += q1;
        }
  double worst = 0, scale = 0, r1 = 0, r2 = 0;
  std::vector<double> w1((size_t)Cout, 0.0), w2((size_t)Cout, 0.0);
  for (int b = 0; b < B; ++b)
    for (int hh = 0; hh < H; ++hh)
      for (int ww = 0; ww < W; ++ww)
        for (int c = 0; c < Cout; ++c) {
          const int ty = hh / 4, i = hh % 4, tx = ww / 4, j = ww % 4;
          const int64_t m = ((int64_t)b * th + ty) * tw + tx;
          double y = 0;
          for (int k = 0; k < 6; ++k)
            for (int l = 0; l < 6; ++l) y += AT[i][k] * (double)M[((size_t)(k * 6 + l) * Mtot + m) * Cout + c] * AT[j][l];
          y = y / 256.0 + (with_bias ? (double)bias[c] : 0.0);
          if (RES == BBDM_RES_SAME) y += res[(((size_t)b * H + hh) * W + ww) * Cout + c];
          if (RES == BBDM_RES_UP2) y += res[(((size_t)b * RH + hh / 2) * RW + ww / 2) * Cout + c];
          if (RES == BBDM_RES_DOWN2) {
            double a = 0;
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx) a += res[(((size_t)b * RH + 2 * hh + dy) * RW + 2 * ww + dx) * Cout + c];
            y += 0.25 * a;
          }
          const double got = out[(((size_t)b * H + hh) * W + ww) * Cout + c];
          worst = std::fmax(worst, std::fabs(got - y));
          scale = std::fmax(scale, std::fabs(y));
          w1[c] += got; w2[c] += got * got;
        }
  for (int c = 0; c < Cout; ++c) {
    r1 = std::fmax(r1, std::fabs(s1[c] - w1[c]) / (1.0 + std::fabs(w1[c])));
    r2 = std::fmax(r2, std::fabs(s2[c] - w2[c]) / (1.0 + std::fabs(w2[c])));
  }
  const bool ok = worst <= 2e-6 * scale && r1 < 1e-4 && r2 < 1e-4;
  std::printf("RES=%d B=%d H=%d W=%d Cout=%d bias=%d: max abs dev %.3e (scale %.3e), partial sums %.1e / %.1e -> %s\n", RES, B, H,
              W, Cout, (int)with_bias, worst, scale, r1, r2, ok ? "ok" : "FAIL");
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<BBDM_RES_NONE>(2, 8, 12, 64, true);
  bad += run<BBDM_RES_NONE>(1, 4, 4, 128, false);
  bad += run<BBDM_RES_SAME>(2, 8, 12, 64, true);
  bad += run<BBDM_RES_SAME>(3, 16, 8, 128, false);
  bad += run<BBDM_RES_UP2>(2, 8, 12, 64, true);
  bad += run<BBDM_RES_UP2>(1, 16, 16, 128, false);
  bad += run<BBDM_RES_DOWN2>(2, 8, 12, 64, true);
  bad += run<BBDM_RES_DOWN2>(1, 4, 8, 128, false);
  std::printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
