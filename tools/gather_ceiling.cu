// gather_ceiling.cu -- what HBM3e gives a kernel shaped like the replay kernels, with NO tape logic at all:
// every warp gathers random 512-byte segments of random rows of a [rows][K] double matrix (K*8-byte rows, the
// workspace layout g[slot*K + d]), eight 16-byte loads per lane in flight, adds them up, and stores one 512-byte
// segment per `ops` gathered (ops = 4: the forward sweep of BASELINE config 4; its algorithmic traffic is exactly
// these bytes).  The achieved GB/s is the ceiling for forward_level_kernel on that workload; the copy bandwidth in
// MEASURED_PEAKS.json (sequential read+write) is not reachable by a random 512-byte gather.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gather_ceiling tools/gather_ceiling.cu
//   tools/gather_ceiling [log2_rows=20] [K=8192] [ops=4] [seg_bytes=512]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
  return z;
}

// one warp per (statement, column block); V doubles per lane
template <int V>
__global__ void __launch_bounds__(256, 3) gather_kernel(double* g, int64_t K, uint32_t row_mask, int64_t n_stmts, int ops,
                                                        int n_cb, uint64_t seed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int cb = blockIdx.y * nw + warp;
  if (cb >= n_cb) return;
  double* col = g + (int64_t(cb) * 32 + lane) * V;
  for (int64_t s = blockIdx.x; s < n_stmts; s += gridDim.x) {
    double acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.0;
    for (int o0 = 0; o0 < ops; o0 += 8) {
      double x[8][V];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (o0 + u < ops) {
          const uint32_t row = uint32_t(mix(seed + uint64_t(s) * 64 + o0 + u)) & row_mask;
          const double* p = col + int64_t(row) * K;
          if (V == 2) { const double2 t = *reinterpret_cast<const double2*>(p); x[u][0] = t.x; x[u][V - 1] = t.y; }
          else if (V == 4) { const double4 t = *reinterpret_cast<const double4*>(p); x[u][0] = t.x; x[u][1 % V] = t.y; x[u][2 % V] = t.z; x[u][3 % V] = t.w; }
          else x[u][0] = *p;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (o0 + u < ops)
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] += 0.5 * x[u][v];
    }
    const uint32_t row = uint32_t(mix(seed ^ (uint64_t(s) * 977 + 13))) & row_mask;
    double* p = col + int64_t(row) * K;
    if (V == 2) *reinterpret_cast<double2*>(p) = make_double2(acc[0], acc[V - 1]);
    else if (V == 4) *reinterpret_cast<double4*>(p) = make_double4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
    else *p = acc[0];
  }
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 20;
  const int64_t K = argc > 2 ? atoll(argv[2]) : 8192;
  const int ops = argc > 3 ? atoi(argv[3]) : 4;
  const int seg = argc > 4 ? atoi(argv[4]) : 512;
  const int64_t rows = int64_t(1) << lg;
  double* g = nullptr;
  if (cudaMalloc(&g, size_t(rows) * K * 8) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
  cudaMemset(g, 0, size_t(rows) * K * 8);
  const int V = seg / 256;   // 256 B -> 1, 512 B -> 2, 1024 B -> 4
  const int n_cb = int(K / (32 * V));
  const int64_t n_stmts = 20000;   // per launch, like a wide step of config 4 (10^4 statements)
  const dim3 grid(unsigned(n_stmts / 4), unsigned((n_cb + 7) / 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f, total = 0;
  const int reps = 20;
  for (int r = 0; r < reps + 3; ++r) {
    cudaEventRecord(e0);
    if (V == 1) gather_kernel<1><<<grid, 256>>>(g, K, uint32_t(rows - 1), n_stmts, ops, n_cb, 1234567ull * (r + 1));
    else if (V == 2) gather_kernel<2><<<grid, 256>>>(g, K, uint32_t(rows - 1), n_stmts, ops, n_cb, 1234567ull * (r + 1));
    else gather_kernel<4><<<grid, 256>>>(g, K, uint32_t(rows - 1), n_stmts, ops, n_cb, 1234567ull * (r + 1));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 3) { best = ms < best ? ms : best; total += ms; }
  }
  if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
  const double bytes = double(n_stmts) * (ops + 1) * double(K) * 8.0;   // ops reads + 1 write of a K-double row per statement
  printf("{\"rows\": %lld, \"K\": %lld, \"ops\": %d, \"segment_bytes\": %d, \"bytes_per_launch\": %.0f, \"best_ms\": %.4f, \"avg_ms\": %.4f, "
         "\"best_GBps\": %.1f, \"avg_GBps\": %.1f}\n", (long long)rows, (long long)K, ops, seg, bytes, best, total / reps,
         bytes / best / 1e6, bytes / (total / reps) / 1e6);
  cudaFree(g);
  return 0;
}
