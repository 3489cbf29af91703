// CPU emulator of the narrow-layer convolution kernels: compiles the SAME per-thread bodies as the CUDA kernels
// (pytorch-gan_b200/csrc/conv_narrow.cuh) with g++ and runs every (class, block, thread) of a launch sequentially.
// Test infrastructure only (tests/test_cpu_narrow_emulation.py); it pins the index arithmetic, not the speed.
#include "../../pytorch-gan_b200/csrc/conv_narrow.cuh"

using namespace b200gan::narrow;

namespace {
Geom make(int N, int H, int W, int C, int P, int Q, int K, int R, int S, int stride, int pad_t, int pad_l, int mode,
          const float *bias, const float *chan_scale, int act, float slope, int rtf) {
  Geom g;
  g.N = N; g.H = H; g.W = W; g.C = C; g.P = P; g.Q = Q; g.K = K; g.R = R; g.S = S;
  g.stride = stride; g.pad_t = pad_t; g.pad_l = pad_l; g.mode = mode;
  g.bias = bias; g.chan_scale = chan_scale; g.act = act; g.slope = slope; g.rtf = rtf;
  return g;
}

template <int KT>
void run_gather(const Geom &g, const float *x, const float *wp, float *y, int block) {
  for (int cls = 0; cls < num_classes(g); ++cls) {
    const int64_t threads = gather_threads(g, KT, cls);
    const int64_t blocks = (threads + block - 1) / block;
    for (int64_t b = 0; b < blocks; ++b)
      for (int tid = 0; tid < block; ++tid) {  // includes the out-of-range tail threads of the last block
        int64_t m;
        int kg;
        gather_coords(g, KT, b * block + tid, &m, &kg);
        gather_thread<KT>(g, x, wp, y, m, kg, cls);
      }
  }
}

template <int DT>
void run_wgrad(const Geom &g, const float *xg, const float *dn, float *dw, int Cd, int splits, int block) {
  const int64_t nout = (int64_t)g.R * g.S * g.C * (Cd / DT);
  const int64_t M = (int64_t)g.N * g.P * g.Q;
  const int64_t per = (M + splits - 1) / splits;
  const int64_t blocks = (nout + block - 1) / block;
  auto add = [](float *p, float v) { *p += v; };
  for (int z = 0; z < splits; ++z)
    for (int64_t b = 0; b < blocks; ++b)
      for (int tid = 0; tid < block; ++tid)
        wgrad_thread<DT>(g, xg, dn, dw, Cd, b * block + tid, z * per, (z + 1) * per, add);
}
}  // namespace

extern "C" int emu_narrow_gather(int N, int H, int W, int C, int P, int Q, int K, int R, int S, int stride, int pad_t,
                                 int pad_l, int mode, const float *bias, const float *chan_scale, int act, float slope,
                                 int rtf, const float *x, const float *wp, float *y, int KT, int block) {
  const Geom g = make(N, H, W, C, P, Q, K, R, S, stride, pad_t, pad_l, mode, bias, chan_scale, act, slope, rtf);
  if (K % KT) return -1;
  if (KT == 8) run_gather<8>(g, x, wp, y, block);
  else if (KT == 4) run_gather<4>(g, x, wp, y, block);
  else if (KT == 1) run_gather<1>(g, x, wp, y, block);
  else return -1;
  return 0;
}

extern "C" int emu_narrow_wgrad(int N, int H, int W, int C, int P, int Q, int Cd, int R, int S, int stride, int pad_t,
                                int pad_l, const float *xg, const float *dn, float *dw, int DT, int splits, int block) {
  const Geom g = make(N, H, W, C, P, Q, Cd, R, S, stride, pad_t, pad_l, 0, nullptr, nullptr, 0, 0.f, 0);
  if (Cd % DT) return -1;
  if (DT == 4) run_wgrad<4>(g, xg, dn, dw, Cd, splits, block);
  else if (DT == 1) run_wgrad<1>(g, xg, dn, dw, Cd, splits, block);
  else return -1;
  return 0;
}
