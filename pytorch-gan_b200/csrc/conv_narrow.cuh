// conv_narrow.cuh -- per-thread bodies of the narrow-layer convolution kernels (EXPERIMENTAL, off by default).
//
// The DCGAN discriminator's 1->16->32(->64) stride-2 layers (dcgan.py:77-88) move ~10 MB per launch but take 24-115 us
// in the generic 64x64x16 gather-GEMM (profiles/r1_kernels_final.txt): they are latency- and index-math-bound, not
// HBM-bound.  These kernels map ONE THREAD to one output pixel x KT output channels, gather with 16-byte loads, read the
// weights straight from L1/L2 (<= 18 KB per layer, warp-coalesced) and handle the stride-2 transposed gather (Conv2d
// dgrad) by output parity class, so every thread of a block walks the same taps.
//
// The bodies are plain C++ (`NB_HD`): the same source is compiled by g++ into a CPU emulator that runs every
// (block, thread) of a launch sequentially (tests/emu/narrow_emu.cpp, tests/test_cpu_narrow_emulation.py) and is checked
// against torch on CPU.  That pins the index arithmetic without a GPU; it says nothing about speed.  Dispatch is
// behind B200GAN_NARROW=1 (conv_simt.cu) until the kernels have been validated and timed on hardware.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif

namespace b200gan {
namespace narrow {

struct alignas(16) F4 {
  float x, y, z, w;
};

// out[n][p][q][k] = act(bias[k] + sum_{r,s,c} in[n][ih][iw][c] * w[(r*S+s)][c][k]) * chan_scale[n][k]
//   mode 0 (Conv2d fprop, ConvTranspose2d dgrad):  ih = p*stride - pad_t + r
//   mode 1 (Conv2d dgrad, ConvTranspose2d fprop):  ih = (p + pad_t - r) / stride where that divides
// zero padding only; activations NHWC fp32; weights packed [tap][contraction channel][output channel].
struct Geom {
  int N, H, W, C;  // gathered tensor
  int P, Q, K;     // produced tensor
  int R, S, stride, pad_t, pad_l;
  int mode;
  const float *bias;
  const float *chan_scale;
  int act;  // B200GAN_ACT_*
  float slope;
  int rtf;  // round the result to TF32 (it feeds a tensor-core conv)
};

NB_HD F4 ld4(const float *p) {
#if defined(__CUDA_ARCH__)
  const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
  return F4{v.x, v.y, v.z, v.w};
#else
  return *reinterpret_cast<const F4 *>(p);
#endif
}
NB_HD float ld1(const float *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits
NB_HD float rtf32(float v) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
#else
  union {
    float f;
    uint32_t u;
  } c;
  c.f = v;
  if ((c.u & 0x7f800000u) != 0x7f800000u) c.u = (c.u + 0x1000u) & ~0x1fffu;
  return c.f;
#endif
}

NB_HD float act_apply(float v, int act, float slope) {
  if (act == 0) return v;                          // B200GAN_ACT_NONE
  if (act == 1) return v > 0.f ? v : v * slope;    // LRELU
  if (act == 2) return fmaxf(v, 0.f);              // RELU
  if (act == 3) return tanhf(v);                   // TANH
  return 1.f / (1.f + expf(-v));                   // SIGMOID
}

// Number of parity classes of a launch: stride^2 for the strided transposed gather, else 1.
NB_HD int num_classes(const Geom &g) { return (g.mode == 1 && g.stride > 1) ? g.stride * g.stride : 1; }

struct ClassInfo {
  int pa, pb;      // first output row / column of the class
  int step;        // output pixel step (stride for a class launch, 1 otherwise)
  int r0, s0;      // first tap of the class
  int Rc, Sc;      // taps of the class per dimension (tap = r0 + step*i)
  int Pc, Qc;      // output pixels of the class per dimension
};

NB_HD ClassInfo class_info(const Geom &g, int cls) {
  ClassInfo ci;
  const int nc = num_classes(g);
  if (nc == 1) {
    ci.pa = ci.pb = 0; ci.step = 1; ci.r0 = ci.s0 = 0; ci.Rc = g.R; ci.Sc = g.S; ci.Pc = g.P; ci.Qc = g.Q;
    return ci;
  }
  const int st = g.stride;
  ci.step = st;
  ci.pa = cls / st;
  ci.pb = cls % st;
  // (p + pad - r) divisible by stride  <=>  r == (p + pad) mod stride, and p == pa mod stride for the whole class
  ci.r0 = (ci.pa + g.pad_t) % st;
  ci.s0 = (ci.pb + g.pad_l) % st;
  ci.Rc = ci.r0 < g.R ? (g.R - ci.r0 + st - 1) / st : 0;
  ci.Sc = ci.s0 < g.S ? (g.S - ci.s0 + st - 1) / st : 0;
  ci.Pc = ci.pa < g.P ? (g.P - ci.pa + st - 1) / st : 0;
  ci.Qc = ci.pb < g.Q ? (g.Q - ci.pb + st - 1) / st : 0;
  return ci;
}

// Flattened launch coordinates: thread t -> (pixel index inside the class, channel group).  Consecutive threads take
// consecutive channel groups of one pixel: coalesced stores and weight loads, broadcast activation loads.
NB_HD int64_t gather_threads(const Geom &g, int KT, int cls) {
  const ClassInfo ci = class_info(g, cls);
  return (int64_t)g.N * ci.Pc * ci.Qc * (g.K / KT);
}
NB_HD void gather_coords(const Geom &g, int KT, int64_t t, int64_t *m, int *kg) {
  const int KG = g.K / KT;
  *m = t / KG;
  *kg = (int)(t % KG);
}

// One thread: output pixel `m` (index inside its parity class), output channels [kg*KT, kg*KT + KT).
// Requirements checked by the launcher: K % KT == 0; C % 4 == 0 or the scalar contraction path (C < 4 or odd C);
// x, wp, y 16-byte aligned when the vector paths are used.
template <int KT>
NB_HD void gather_thread(const Geom &g, const float *x, const float *wp, float *y, int64_t m, int kg, int cls) {
  const ClassInfo ci = class_info(g, cls);
  const int64_t Mc = (int64_t)g.N * ci.Pc * ci.Qc;
  if (m >= Mc || kg * KT >= g.K) return;
  const int q = ci.pb + ci.step * (int)(m % ci.Qc);
  const int64_t t = m / ci.Qc;
  const int p = ci.pa + ci.step * (int)(t % ci.Pc);
  const int n = (int)(t / ci.Pc);
  const int k0 = kg * KT;

  float acc[KT];
#pragma unroll
  for (int j = 0; j < KT; ++j) acc[j] = 0.f;

  for (int ir = 0; ir < ci.Rc; ++ir) {
    const int r = ci.r0 + ci.step * ir;
    int ih;
    if (g.mode == 0) {
      ih = p * g.stride - g.pad_t + r;
    } else {
      const int th = p + g.pad_t - r;
      if (th < 0) continue;
      ih = th / g.stride;  // divisible by construction of the class (stride 1: trivially)
    }
    if (ih < 0 || ih >= g.H) continue;
    for (int is = 0; is < ci.Sc; ++is) {
      const int s = ci.s0 + ci.step * is;
      int iw;
      if (g.mode == 0) {
        iw = q * g.stride - g.pad_l + s;
      } else {
        const int tw = q + g.pad_l - s;
        if (tw < 0) continue;
        iw = tw / g.stride;
      }
      if (iw < 0 || iw >= g.W) continue;
      const float *xp = x + ((int64_t)(n * g.H + ih) * g.W + iw) * g.C;
      const float *wt = wp + (int64_t)(r * g.S + s) * g.C * g.K + k0;
      int c = 0;
      if ((g.C & 3) == 0) {
        for (; c < g.C; c += 4) {
          const F4 xv = ld4(xp + c);
          const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float *wr = wt + (int64_t)(c + j) * g.K;
            if (KT % 4 == 0) {
#pragma unroll
              for (int u = 0; u < KT / 4; ++u) {
                const F4 w4 = ld4(wr + 4 * u);
                acc[4 * u + 0] = fmaf(xs[j], w4.x, acc[4 * u + 0]);
                acc[4 * u + 1] = fmaf(xs[j], w4.y, acc[4 * u + 1]);
                acc[4 * u + 2] = fmaf(xs[j], w4.z, acc[4 * u + 2]);
                acc[4 * u + 3] = fmaf(xs[j], w4.w, acc[4 * u + 3]);
              }
            } else {
#pragma unroll
              for (int u = 0; u < KT; ++u) acc[u] = fmaf(xs[j], ld1(wr + u), acc[u]);
            }
          }
        }
      } else {
        for (; c < g.C; ++c) {
          const float xs = ld1(xp + c);
          const float *wr = wt + (int64_t)c * g.K;
          if (KT % 4 == 0) {
#pragma unroll
            for (int u = 0; u < KT / 4; ++u) {
              const F4 w4 = ld4(wr + 4 * u);
              acc[4 * u + 0] = fmaf(xs, w4.x, acc[4 * u + 0]);
              acc[4 * u + 1] = fmaf(xs, w4.y, acc[4 * u + 1]);
              acc[4 * u + 2] = fmaf(xs, w4.z, acc[4 * u + 2]);
              acc[4 * u + 3] = fmaf(xs, w4.w, acc[4 * u + 3]);
            }
          } else {
#pragma unroll
            for (int u = 0; u < KT; ++u) acc[u] = fmaf(xs, ld1(wr + u), acc[u]);
          }
        }
      }
    }
  }

  float *yo = y + ((int64_t)(n * g.P + p) * g.Q + q) * g.K + k0;
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    float v = acc[j];
    if (g.bias) v += ld1(g.bias + k0 + j);
    v = act_apply(v, g.act, g.slope);
    if (g.chan_scale) v *= ld1(g.chan_scale + (int64_t)n * g.K + k0 + j);
    if (g.rtf) v = rtf32(v);
    acc[j] = v;
  }
  if (KT % 4 == 0) {
#pragma unroll
    for (int u = 0; u < KT / 4; ++u)
      *reinterpret_cast<F4 *>(yo + 4 * u) = F4{acc[4 * u], acc[4 * u + 1], acc[4 * u + 2], acc[4 * u + 3]};
  } else {
#pragma unroll
    for (int j = 0; j < KT; ++j) yo[j] = acc[j];
  }
}

// Weight gradient.  xg: gathered tensor [N][H][W][C] (mode 0 addressing), dn: dense tensor [N][P][Q][Cd];
//   dw[cd][c][r][s] += sum_{pixels m in [m_begin, m_end)} xg[n][p*stride-pad+r][q*stride-pad+s][c] * dn[n][p][q][cd]
// One thread: output group o = ((r*S + s)*C + c)*DG + dg with DG = Cd/DT, i.e. DT consecutive dn channels of one
// (tap, c).  Consecutive threads read consecutive dn channels (coalesced) and the same xg scalar (broadcast).
// `add` receives (pointer, value): atomicAdd on the device, += in the emulator.
template <int DT, class Add>
NB_HD void wgrad_thread(const Geom &g, const float *xg, const float *dn, float *dw, int Cd, int64_t o, int64_t m_begin,
                        int64_t m_end, Add add) {
  const int DG = Cd / DT;
  const int64_t nout = (int64_t)g.R * g.S * g.C * DG;
  const int64_t M = (int64_t)g.N * g.P * g.Q;
  if (o >= nout) return;
  if (m_end > M) m_end = M;
  if (m_begin >= m_end) return;
  const int dg = (int)(o % DG);
  int64_t t = o / DG;
  const int c = (int)(t % g.C);
  t /= g.C;
  const int s = (int)(t % g.S);
  const int r = (int)(t / g.S);
  const int cd0 = dg * DT;

  int q = (int)(m_begin % g.Q);
  int64_t t2 = m_begin / g.Q;
  int p = (int)(t2 % g.P);
  int n = (int)(t2 / g.P);

  float acc[DT];
#pragma unroll
  for (int j = 0; j < DT; ++j) acc[j] = 0.f;

  for (int64_t m = m_begin; m < m_end; ++m) {
    const int ih = p * g.stride - g.pad_t + r;
    const int iw = q * g.stride - g.pad_l + s;
    if (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) {
      const float xv = ld1(xg + ((int64_t)(n * g.H + ih) * g.W + iw) * g.C + c);
      const float *dp = dn + m * Cd + cd0;
      if (DT % 4 == 0) {
#pragma unroll
        for (int u = 0; u < DT / 4; ++u) {
          const F4 d4 = ld4(dp + 4 * u);
          acc[4 * u + 0] = fmaf(xv, d4.x, acc[4 * u + 0]);
          acc[4 * u + 1] = fmaf(xv, d4.y, acc[4 * u + 1]);
          acc[4 * u + 2] = fmaf(xv, d4.z, acc[4 * u + 2]);
          acc[4 * u + 3] = fmaf(xv, d4.w, acc[4 * u + 3]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < DT; ++u) acc[u] = fmaf(xv, ld1(dp + u), acc[u]);
      }
    }
    if (++q == g.Q) {
      q = 0;
      if (++p == g.P) {
        p = 0;
        ++n;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < DT; ++j)
    add(dw + (((int64_t)(cd0 + j) * g.C + c) * g.R + r) * g.S + s, acc[j]);
}

}  // namespace narrow
}  // namespace b200gan
