// Micro-benchmark deciding the intermediate layout of the fused 2-D engine (DESIGN.md "Layout of the mixed-space
// intermediates"): effective HBM bandwidth of a transpose-free "tile layout" hand-off between a column kernel
// (owns C adjacent kr columns, all y) and a row kernel (owns G adjacent y rows, all kr).
//   layout T(C): element (r, k) at ((k / C) * R + r) * C + (k % C)      [r = row (y), k = column (kr)], 16-byte elements
//   gather : read T(C) row-wise (what the row kernel does), write row-major contiguous
//   scatter: read row-major... no: read column-major contiguous (what the column kernel holds), write T(C)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o strided_bw strided_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_copy(const double2* __restrict__ in, double2* __restrict__ out, long n) {
  long i = (long)blockIdx.x * blockDim.x * 8 + threadIdx.x;
  double2 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) if (i + u * 256 < n) v[u] = in[i + u * 256];
#pragma unroll
  for (int u = 0; u < 8; ++u) if (i + u * 256 < n) out[i + u * 256] = v[u];
}

// row kernel pattern: CTA owns G adjacent rows; for each row reads all K columns from T(C); writes row-major.
template <int C>
__global__ void __launch_bounds__(256) k_gather(const double2* __restrict__ in, double2* __restrict__ out, int R, int K, int G) {
  int r0 = blockIdx.x * G;
  for (int g = 0; g < G; ++g) {
    int r = r0 + g;
    for (int kb = 0; kb < K; kb += 256 * 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int k = kb + u * 256 + threadIdx.x;
        if (k < K) v[u] = in[((long)(k / C) * R + r) * C + (k % C)];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int k = kb + u * 256 + threadIdx.x;
        if (k < K) out[(long)r * K + k] = v[u];
      }
    }
  }
}

// column kernel pattern: CTA owns C adjacent columns (processed one after the other); reads column-major contiguous
// (in[k * R + r]); writes T(C).
template <int C>
__global__ void __launch_bounds__(256) k_scatter(const double2* __restrict__ in, double2* __restrict__ out, int R, int K) {
  int k0 = blockIdx.x * C;
  for (int c = 0; c < C; ++c) {
    int k = k0 + c;
    if (k >= K) break;
    for (int rb = 0; rb < R; rb += 256 * 8) {
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int r = rb + u * 256 + threadIdx.x;
        if (r < R) v[u] = in[(long)k * R + r];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        int r = rb + u * 256 + threadIdx.x;
        if (r < R) out[((long)(k / C) * R + r) * C + (k % C)] = v[u];
      }
    }
  }
}

template <class F>
float timeit(F f, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  const int R = 4096, K = 2048;          // 134 MB per array (> L2)
  long n = (long)R * K;
  double2 *a, *b;
  CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
  CK(cudaMemset(a, 0, n * 16)); CK(cudaMemset(b, 0, n * 16));
  double gb = 2.0 * n * 16 / 1e9;
  float ms = timeit([&] { k_copy<<<(unsigned)((n + 2047) / 2048), 256>>>(a, b, n); });
  printf("copy            : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
  for (int G : {1, 2, 4, 8}) {
    ms = timeit([&] { k_gather<1><<<R / G, 256>>>(a, b, R, K, G); }); printf("gather C=1 G=%d  : %.3f ms  %.0f GB/s\n", G, ms, gb / ms * 1e3);
    ms = timeit([&] { k_gather<2><<<R / G, 256>>>(a, b, R, K, G); }); printf("gather C=2 G=%d  : %.3f ms  %.0f GB/s\n", G, ms, gb / ms * 1e3);
    ms = timeit([&] { k_gather<4><<<R / G, 256>>>(a, b, R, K, G); }); printf("gather C=4 G=%d  : %.3f ms  %.0f GB/s\n", G, ms, gb / ms * 1e3);
    ms = timeit([&] { k_gather<8><<<R / G, 256>>>(a, b, R, K, G); }); printf("gather C=8 G=%d  : %.3f ms  %.0f GB/s\n", G, ms, gb / ms * 1e3);
  }
  ms = timeit([&] { k_scatter<1><<<K / 1, 256>>>(a, b, R, K); }); printf("scatter C=1     : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
  ms = timeit([&] { k_scatter<2><<<K / 2, 256>>>(a, b, R, K); }); printf("scatter C=2     : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
  ms = timeit([&] { k_scatter<4><<<K / 4, 256>>>(a, b, R, K); }); printf("scatter C=4     : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
  ms = timeit([&] { k_scatter<8><<<K / 8, 256>>>(a, b, R, K); }); printf("scatter C=8     : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
  return 0;
}
