// microbench.cu — throughput of the shared-memory / warp primitives the count path is built from,
// measured on the GPU box (B200). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
constexpr int NB = 2400;  // counters (static smem <= 48 KB)

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint64_t* gbuf, uint32_t nbuck) {
  __shared__ uint32_t sh[NB];
  __shared__ unsigned long long sh64[NB];
  __shared__ uint8_t wp[8][NB];   // warp-private u8 counters
  for (int i = threadIdx.x; i < NB; i += blockDim.x) { sh[i] = 0; sh64[i] = ~0ull; }
  for (int i = threadIdx.x; i < 8 * NB; i += blockDim.x) ((uint8_t*)wp)[i] = 0;
  __syncthreads();
  uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1;
  uint32_t acc = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = 0; it < ITER; ++it) {
    const uint32_t b = lcg(s) % nbuck;
    if (MODE == 0) { atomicAdd(&sh[b], 1u); }                        // ATOMS.ADD no return
    if (MODE == 1) { acc += atomicAdd(&sh[b], 1u); }                 // ATOMS.ADD with return
    if (MODE == 2) { acc += __match_any_sync(0xffffffffu, b); }      // MATCH.ANY
    if (MODE == 3) {                                                  // match + warp-private u8 RMW (proposed ranking)
      const unsigned peers = __match_any_sync(0xffffffffu, b);
      const uint32_t old = wp[warp][b];
      __syncwarp();
      if ((peers & ((1u << lane) - 1)) == 0) wp[warp][b] = (uint8_t)(old + __popc(peers));
      __syncwarp();
      acc += old + __popc(peers & ((1u << lane) - 1));
    }
    if (MODE == 4) {                                                  // CAS64 + add (tile insert, miss path)
      unsigned long long key = ((unsigned long long)b << 20) | 5;
      unsigned long long cur = sh64[b];
      if (cur == ~0ull) cur = atomicCAS(&sh64[b], ~0ull, key);
      atomicAdd(&sh[b], 1u);
      acc += (uint32_t)cur;
    }
    if (MODE == 5) {                                                  // LDS64 compare + ATOMS.ADD (tile insert, hit path)
      unsigned long long cur = sh64[b];
      if (cur != 12345ull) atomicAdd(&sh[b], 1u);
    }
    if (MODE == 6) {                                                  // scattered 8-byte global stores, frontier pattern
      const uint32_t pos = (blockIdx.x * 256u + threadIdx.x) + it;   // advancing frontier per bucket
      gbuf[(uint64_t)b * 65536u + (pos & 65535u)] = b;
    }
    if (MODE == 7) {                                                  // LDS64 compare + non-atomic u32 RMW (if races were excluded)
      unsigned long long cur = sh64[b];
      if (cur != 12345ull) sh[b] = sh[b] + 1;
    }
    if (MODE == 8) {                                                  // ballot-based conflict detect fast path
      wp[warp][b] = (uint8_t)lane;
      __syncwarp();
      const bool lost = wp[warp][b] != lane;
      acc += __ballot_sync(0xffffffffu, lost);
    }
  }
  if (acc == 0xdeadbeef) out[0] = acc + sh[threadIdx.x] + (uint32_t)sh64[threadIdx.x];
  if (MODE <= 1 || MODE == 4 || MODE == 5 || MODE == 7) { __syncthreads(); if (threadIdx.x == 0 && sh[1] == 0xffffffff) out[1] = 1; }
}

template <int MODE>
void run(const char* name, uint32_t nbuck, uint32_t* out, uint64_t* gbuf, int sms, double ghz) {
  const int grid = sms * 4;
  k<MODE><<<grid, 256>>>(out, gbuf, nbuck);
  cudaDeviceSynchronize();
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<grid, 256>>>(out, gbuf, nbuck);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double lane_ops = (double)grid * 256 * ITER;
  const double cyc_per_lane_per_sm = ms * 1e-3 * ghz * 1e9 / (lane_ops / sms);
  printf("%-44s nbuck=%5u  %8.3f ms  %6.3f cyc/lane-op/SM  %7.1f Gop/s  err=%s\n", name, nbuck, ms, cyc_per_lane_per_sm,
         lane_ops / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double ghz = p.clockRate * 1e-6;
  printf("%s SMs=%d clock=%.3f GHz\n", p.name, sms, ghz);
  uint32_t* out; cudaMalloc(&out, 1024);
  uint64_t* gbuf; cudaMalloc(&gbuf, (uint64_t)4096 * 65536 * 8);
  for (uint32_t nb : {2368u, 256u, 32u}) {
    run<0>("ATOMS.ADD no-return", nb, out, gbuf, sms, ghz);
    run<1>("ATOMS.ADD with return", nb, out, gbuf, sms, ghz);
    run<2>("MATCH.ANY b32", nb, out, gbuf, sms, ghz);
    run<3>("MATCH + warp-private u8 LDS/STS rank", nb, out, gbuf, sms, ghz);
    run<8>("tag STS/LDS + ballot conflict detect", nb, out, gbuf, sms, ghz);
  }
  run<4>("LDS64 + CAS64(if empty) + ATOMS.ADD", 2400, out, gbuf, sms, ghz);
  run<5>("LDS64 compare + ATOMS.ADD", 2400, out, gbuf, sms, ghz);
  run<7>("LDS64 compare + non-atomic LDS/STS u32", 2400, out, gbuf, sms, ghz);
  run<6>("scattered 8B STG, 2368 frontiers", 2368, out, gbuf, sms, ghz);
  run<6>("scattered 8B STG, 148 frontiers", 148, out, gbuf, sms, ghz);
  return 0;
}
