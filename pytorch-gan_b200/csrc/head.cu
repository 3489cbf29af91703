// head.cu -- the Discriminator head and the adversarial loss of the DCGAN step as single small kernels.
//
// Reference call sites:
//   self.adv_layer = nn.Sequential(nn.Linear(128 * ds_size ** 2, 1), nn.Sigmoid())      dcgan.py:92
//   adversarial_loss = torch.nn.BCELoss()                                                dcgan.py:103
//   g_loss = adversarial_loss(discriminator(gen_imgs), valid)                            dcgan.py:166, 178-180
// Stock torch runs these as a cuBLASLt GEMV (+ split-K reduce + bias epilogue), a sigmoid kernel, a BCE kernel and a
// mean reduction forward, and five more backward -- ~12 launches per discriminator pass for 128 x 2048 numbers.
// Here: one kernel per direction for Linear(K -> 1) [+ Sigmoid], one per direction for the BCE mean.
#include "common.cuh"

namespace b200gan {

__device__ __forceinline__ float block_sum_128(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  __syncthreads();
  return t;
}

// y[n] = act(dot(x[n], w) + b).  One block (128 threads) per sample.
__global__ void __launch_bounds__(128)
linear1_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                   float *__restrict__ y, int K, int act) {
  __shared__ float red[4];
  const float *xr = x + (int64_t)blockIdx.x * K;
  float s = 0.f;
  if ((K & 3) == 0 && (((uintptr_t)x | (uintptr_t)w) & 15) == 0) {
    const float4 *x4 = reinterpret_cast<const float4 *>(xr), *w4 = reinterpret_cast<const float4 *>(w);
    for (int i = threadIdx.x; i < (K >> 2); i += 128) {
      const float4 a = __ldg(x4 + i), c = __ldg(w4 + i);
      s += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    }
  } else {
    for (int i = threadIdx.x; i < K; i += 128) s = fmaf(__ldg(xr + i), __ldg(w + i), s);
  }
  s = block_sum_128(s, red);
  if (threadIdx.x == 0) y[blockIdx.x] = apply_act(s + (b ? __ldg(b) : 0.f), act, 0.f);
}

// dl[n] = dy[n] * act'(y[n]);  dw[k] = sum_n dl[n] x[n][k];  dx[n][k] = dl[n] w[k];  db = sum_n dl[n].
// grid = ceil(K / 128); thread = one k.  N <= 4096 (dl staged in shared memory).
__global__ void __launch_bounds__(128)
linear1_bwd_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ y,
                   const float *__restrict__ dy, float *__restrict__ dx, float *__restrict__ dw,
                   float *__restrict__ db, int N, int K, int act) {
  extern __shared__ float dl[];  // [N]
  __shared__ float red[4];
  float part = 0.f;
  for (int n = threadIdx.x; n < N; n += 128) {
    const float v = __ldg(dy + n) * act_grad_from_out(__ldg(y + n), act, 0.f);
    dl[n] = v;
    part += v;
  }
  __syncthreads();
  if (db && blockIdx.x == 0) {
    const float t = block_sum_128(part, red);
    if (threadIdx.x == 0) *db = t;
  }
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= K) return;
  const float wk = __ldg(w + k);
  float acc = 0.f;
  for (int n = 0; n < N; ++n) {
    const float d = dl[n];
    acc = fmaf(d, __ldg(x + (int64_t)n * K + k), acc);
    if (dx) dx[(int64_t)n * K + k] = d * wk;
  }
  dw[k] = acc;
}

// torch.nn.BCELoss(reduction='mean'): log terms clamped at -100 (torch/aten binary_cross_entropy semantics).
__global__ void __launch_bounds__(128)
bce_fwd_kernel(const float *__restrict__ v, const float *__restrict__ t, float *__restrict__ loss, int64_t n) {
  __shared__ float red[4];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 128) {
    const float p = __ldg(v + i), y = __ldg(t + i);
    const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(log1pf(-p), -100.f);
    s += (y - 1.f) * lq - y * lp;
  }
  s = block_sum_128(s, red);
  if (threadIdx.x == 0) *loss = s / (float)n;
}
// d loss / d v = gout / n * (v - t) / max((1 - v) v, 1e-12)
__global__ void bce_bwd_kernel(const float *__restrict__ v, const float *__restrict__ t, const float *__restrict__ gout,
                               float *__restrict__ dv, int64_t n) {
  const float g = __ldg(gout) / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = __ldg(v + i), y = __ldg(t + i);
    dv[i] = g * (p - y) / fmaxf((1.f - p) * p, 1e-12f);
  }
}

}  // namespace b200gan

using namespace b200gan;

extern "C" int b200gan_linear1_fwd(const float *x, const float *w, const float *b, float *y, int32_t N, int32_t K,
                                   int32_t act, void *stream) {
  B2_CHECK_ARG(x && w && y && N >= 0 && K > 0, "linear1_fwd: bad arguments");
  if (N == 0) return B200GAN_OK;
  linear1_fwd_kernel<<<(unsigned)N, 128, 0, as_stream(stream)>>>(x, w, b, y, K, act);
  B2_LAUNCH_CHECK();
  return B200GAN_OK;
}

extern "C" int b200gan_linear1_bwd(const float *x, const float *w, const float *y, const float *dy, float *dx, float *dw,
                                   float *db, int32_t N, int32_t K, int32_t act, void *stream) {
  B2_CHECK_ARG(x && w && y && dy && dw && N > 0 && K > 0, "linear1_bwd: bad arguments");
  B2_CHECK_ARG(N <= 4096, "linear1_bwd: batch %d > 4096", N);
  linear1_bwd_kernel<<<(unsigned)ceil_div(K, 128), 128, (size_t)N * sizeof(float), as_stream(stream)>>>(
      x, w, y, dy, dx, dw, db, N, K, act);
  B2_LAUNCH_CHECK();
  return B200GAN_OK;
}

extern "C" int b200gan_bce_fwd(const float *v, const float *t, float *loss, int64_t n, void *stream) {
  B2_CHECK_ARG(v && t && loss && n > 0, "bce_fwd: bad arguments");
  bce_fwd_kernel<<<1, 128, 0, as_stream(stream)>>>(v, t, loss, n);
  B2_LAUNCH_CHECK();
  return B200GAN_OK;
}

extern "C" int b200gan_bce_bwd(const float *v, const float *t, const float *gout, float *dv, int64_t n, void *stream) {
  B2_CHECK_ARG(v && t && gout && dv && n > 0, "bce_bwd: bad arguments");
  int64_t blocks = ceil_div64(n, 256);
  if (blocks > 1184) blocks = 1184;
  bce_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(v, t, gout, dv, n);
  B2_LAUNCH_CHECK();
  return B200GAN_OK;
}
