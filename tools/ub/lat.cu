// Dependent-chain latencies of the warp primitives the stream reducer's round is made of (one warp, clock64 around 256 ops).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 256
__global__ void k(unsigned* out, long long* t, const unsigned* g, int nwarps_bar) {
  __shared__ unsigned sm[1024];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + 1) & 1023;
  __syncthreads();
  unsigned x = threadIdx.x * 2654435761u + out[0];
  long long a, b;
  int s = 0;
#define T0 a = clock64();
#define T1 b = clock64(); if (threadIdx.x == 0) t[s] = b - a; s++;
  T0 for (int i = 0; i < N; i++) x = __reduce_min_sync(0xffffffffu, x + lane) + 1; T1                       // 0 redux.min
  T0 for (int i = 0; i < N; i++) x = __reduce_add_sync(0xffffffffu, x ^ lane); T1                           // 1 redux.add
  T0 for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1 + (i & 15)) + 1; T1                  // 2 shfl
  T0 for (int i = 0; i < N; i++) x = __ballot_sync(0xffffffffu, (x >> lane) & 1) + i; T1                    // 3 ballot
  T0 for (int i = 0; i < N; i++) x = sm[x & 1023]; T1                                                       // 4 lds chain
  T0 for (int i = 0; i < N; i++) x = g[x & 1023]; T1                                                        // 5 ldg chain (L1)
  T0 for (int i = 0; i < N; i++) x = __ldcg(g + (x & 1023)); T1                                             // 6 ldg.cg chain (L2)
  T0 for (int i = 0; i < N; i++) { __syncthreads(); } T1                                                     // 7 bar
  T0 for (int i = 0; i < N; i++) { if (lane == 0) sm[threadIdx.x >> 5] = x; __syncthreads(); x += sm[lane & 7]; } T1  // 8 sts+bar+lds
  T0 for (int i = 0; i < N; i++) x = x * 2654435761u + 12345u; T1                                           // 9 imad chain
  T0 for (int i = 0; i < N; i++) x = __umulhi(x, 0x9e3779b9u) + x; T1                                       // 10 umulhi chain
  T0 for (int i = 0; i < N; i++) x = __match_any_sync(0xffffffffu, x & 3) + i; T1                           // 11 match
  { unsigned long long y = x; y = (y << 32) | x;
    T0 for (int i = 0; i < N; i++) {   // 12: 64-bit min by two redux + sum redux (the round's fold)
      const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(y >> 32));
      const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(y >> 32) == hi ? (unsigned)y : 0xffffffffu);
      const unsigned sum = __reduce_add_sync(0xffffffffu, ((unsigned)(y >> 32) == hi && (unsigned)y == lo) ? lane : 0u);
      y += (((unsigned long long)hi << 32) | lo) + sum + lane;
    } T1
    T0 for (int i = 0; i < N; i++) {   // 13: 64-bit (min, sum) butterfly by shuffles
      unsigned long long k2 = y; unsigned sm2 = lane;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, k2, o); const unsigned os = __shfl_xor_sync(0xffffffffu, sm2, o);
        if (ok < k2) { k2 = ok; sm2 = os; } else if (ok == k2) sm2 += os;
      }
      y += k2 + sm2 + lane;
    } T1
    x += (unsigned)y; }
  out[threadIdx.x] = x;
}
int main() {
  unsigned *out, *g; long long* t;
  cudaMalloc(&out, 4096); cudaMemset(out, 0, 4096); cudaMalloc(&g, 4096); cudaMallocManaged(&t, 64 * 8);
  unsigned hg[1024]; for (int i = 0; i < 1024; i++) hg[i] = (i * 13 + 5) & 1023; cudaMemcpy(g, hg, 4096, cudaMemcpyHostToDevice);
  const char* nm[] = {"redux.min", "redux.add", "shfl", "ballot", "lds", "ldg L1", "ldg.cg L2", "bar", "sts+bar+lds", "imad", "umulhi+add", "match_any", "min64+sum 3xredux", "min64+sum shfl butterfly"};
  for (int w : {1, 4, 8}) {
    for (int rep = 0; rep < 2; rep++) { k<<<1, 32 * w>>>(out, t, g, w); cudaDeviceSynchronize(); }
    printf("warps %d:", w);
    for (int i = 0; i < 14; i++) printf("  %s %.1f", nm[i], (double)t[i] / N);
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
