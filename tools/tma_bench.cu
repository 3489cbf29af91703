// tma_bench.cu -- how fast can one SM pull 128-byte-row boxes through TMA?  (bring-up microbenchmark)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pytorch-gan_b200/csrc tools/tma_bench.cu \
//        -L pytorch-gan_b200/b200gan -lb200gan -Xlinker -rpath=$PWD/pytorch-gan_b200/b200gan -o gpurun_out/tma_bench
//
// Each CTA streams `iters` boxes of `rows` x 128 B through a ring of `stages` smem buffers; a consumer thread
// frees a slot as soon as it is full.  Variants: tensor rank (2 vs 5), rows per box, stages, CTAs per SM.
#include "tc_common.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>

using namespace b200gan;

struct BP {
  int rank, iters, stages, rows_w, rows_h;  // 5D: box {32, rows_w, 1, rows_h, 1}
  int tiles_w, tiles_h, N, chunks;
};

__global__ void __launch_bounds__(64) tma_stream_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ BP p,
                                                        unsigned long long *sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int box_bytes = p.rows_w * p.rows_h * 128;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.stages * box_bytes);
  uint64_t *empty = full + p.stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int ntiles = p.tiles_w * p.tiles_h * p.N;
  if (threadIdx.x == 0) {
    int stage = 0;
    uint32_t phase = 0;
    int t = (blockIdx.x * 7) % ntiles, c = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], box_bytes);
      int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, n = t / (p.tiles_w * p.tiles_h);
      if (p.rank == 5)
        tma_load_5d(smem + stage * box_bytes, &tm, &full[stage], c * 32, tw * p.rows_w, 0, th * p.rows_h, n);
      else
        tma_load_2d(smem + stage * box_bytes, &tm, &full[stage], c * 32, t * p.rows_w * p.rows_h);
      if (++c == p.chunks) {
        c = 0;
        t = (t + 1) % ntiles;
      }
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long acc = 0;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(&full[stage], phase);
      acc += *reinterpret_cast<volatile unsigned int *>(smem + stage * box_bytes);
      mbar_arrive(&empty[stage]);
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (acc == 0x12345678ull) *sink = acc;
  }
}

int main() {
  const int N = 128, H = 32, W = 32, C = 128;  // DCGAN conv2 input: 67 MB
  float *x;
  unsigned long long *sink;
  cudaMalloc(&x, (size_t)N * H * W * C * 4);
  cudaMemset(x, 0, (size_t)N * H * W * C * 4);
  cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int rank, rw, rh, stages, ctas_per_sm; };
  std::vector<Cfg> cfgs = {
      {5, 32, 4, 3, 2}, {5, 32, 4, 4, 2}, {5, 32, 4, 8, 1}, {5, 32, 4, 6, 2}, {2, 32, 4, 3, 2}, {2, 32, 4, 8, 1},
      {5, 32, 8, 3, 2}, {5, 32, 8, 6, 1}, {2, 32, 8, 6, 1}, {5, 32, 1, 8, 2}, {5, 32, 2, 8, 2}, {5, 16, 8, 4, 2},
  };
  printf("%-6s %-10s %-7s %-8s %10s %12s %14s\n", "rank", "box rows", "stages", "CTA/SM", "us", "TB/s total", "B/clk/SM@1.9G");
  for (auto c : cfgs) {
    BP p;
    p.rank = c.rank; p.iters = 512; p.stages = c.stages; p.rows_w = c.rw; p.rows_h = c.rh;
    p.tiles_w = W / c.rw; p.tiles_h = H / c.rh; p.N = N; p.chunks = C / 32;
    CUtensorMap tm;
    if (c.rank == 5) {
      uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, 1, (uint64_t)H, (uint64_t)N};
      uint64_t st[4] = {(uint64_t)C * 4, (uint64_t)W * C * 4, (uint64_t)W * C * 4, (uint64_t)H * W * C * 4};
      uint32_t box[5] = {32, (uint32_t)c.rw, 1, (uint32_t)c.rh, 1};
      if (make_tmap_f32(&tm, x, 5, dims, st, box)) { printf("tmap fail: %s\n", b200gan_last_error()); return 1; }
    } else {
      uint64_t dims[2] = {(uint64_t)C, (uint64_t)N * H * W};
      uint64_t st[1] = {(uint64_t)C * 4};
      uint32_t box[2] = {32, (uint32_t)(c.rw * c.rh)};
      if (make_tmap_f32(&tm, x, 2, dims, st, box)) { printf("tmap fail: %s\n", b200gan_last_error()); return 1; }
    }
    int grid = 148 * c.ctas_per_sm;
    size_t smem = (size_t)c.stages * c.rw * c.rh * 128 + 1024 + 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    tma_stream_kernel<<<grid, 64, smem>>>(tm, p, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    tma_stream_kernel<<<grid, 64, smem>>>(tm, p, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double bytes = (double)grid * p.iters * c.rw * c.rh * 128;
    printf("%-6d %2dx%-7d %-7d %-8d %10.1f %12.2f %14.1f   %s\n", c.rank, c.rw, c.rh, c.stages, c.ctas_per_sm, ms * 1e3,
           bytes / ms / 1e9, bytes / 148 / (ms * 1e-3 * 1.9e9), err == cudaSuccess ? "" : cudaGetErrorString(err));
  }
  return 0;
}
