// Development aid (not a test): what does a chain of small tcgen05.mma kind::tf32 instructions cost on a B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/mma_issue tests/ubench/mma_issue.cu && gpurun_out/mma_issue
// For each variant one CTA issues NMMA = 38 MMAs (the K = 300 reduction of the GRU forward step) from one elected thread, commits and waits;
// clock64 deltas: issue = until the last MMA has been issued, total = until the commit's mbarrier fires.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../gesture-generation-from-trimodal-context_b200/csrc/umma.cuh"
using namespace umma;

constexpr int NMMA = 38;

template <int M, int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 210 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 210 * 1024 / 16; i += 128) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
  constexpr uint32_t idesc = idesc_tf32(M, N, 0, 0);
  constexpr uint32_t hi = desc_hi(1024, 2);
  constexpr uint32_t A0 = 64u * NACC >= (uint32_t)N * NACC ? 64u * NACC : (uint32_t)N * NACC;   // accumulators first, then A (TS)
  constexpr uint32_t ACC_STRIDE = N > 64 ? N : 64;
  if (warp == 1) {
    long long t_issue = 0, t_total = 0;
    for (int r = 0; r < reps; ++r) {
      __syncwarp();
      if (elect_one()) {
        const uint32_t a_lo = desc_lo(smem_u32(smem), 16), b_lo = desc_lo(smem_u32(smem) + 144 * 1024, 16);      // A: 9 chunks x 16 KB; B: 2 chunks, reused
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          const uint32_t d = tmem_base + ACC_STRIDE * (i % NACC);
          const uint32_t acc = i >= NACC ? 1u : 0u;
          // K-step i: chunk i / 4 (16 KB apart for A, N * 128 B for B), 32 B per K-step inside the chunk
          if constexpr (TS) mma_tf32_ts(d, tmem_base + A0 + 8u * i, b_lo + ((i / 4) % 2) * (N * 128 / 16) + (i % 4) * 2, hi, idesc, acc);
          else mma_tf32_lohi(d, a_lo + (i / 4) * (16384 / 16) + (i % 4) * 2, hi, b_lo + ((i / 4) % 2) * (N * 128 / 16) + (i % 4) * 2, hi, idesc, acc);
        }
        const long long t1 = clock64();
        tc_commit(bar);
        t_issue += t1 - t0;
        out[2] = t0;
      }
      __syncwarp();
      mbar_wait(bar, (uint32_t)(r & 1));
      const long long t2 = clock64();
      if (elect_one()) t_total += t2 - out[2];
      __syncwarp();
    }
    if (elect_one()) { out[0] = t_issue / reps; out[1] = t_total / reps; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int M, int N, bool TS, int NACC>
void run(const char* name, long long* dbuf) {
  cudaFuncSetAttribute(k<M, N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 211 * 1024);
  k<M, N, TS, NACC><<<1, 128, 211 * 1024>>>(dbuf, 50);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[3] = {0, 0, 0};
  cudaMemcpy(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s M=%3d N=%3d %s nacc=%d : issue %6lld cyc (%5.1f / MMA)   total %6lld cyc (%5.1f / MMA)   %s\n", name, M, N, TS ? "TS" : "SS", NACC, h[0],
         (double)h[0] / NMMA, h[1], (double)h[1] / NMMA, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* dbuf;
  cudaMalloc(&dbuf, 64);
  run<128, 64, false, 1>("SS one accumulator", dbuf);
  run<128, 32, false, 1>("SS one accumulator", dbuf);
  run<128, 16, false, 1>("SS one accumulator", dbuf);
  run<64, 24, false, 1>("SS one accumulator", dbuf);
  run<64, 64, false, 1>("SS one accumulator", dbuf);
  run<128, 128, false, 1>("SS one accumulator", dbuf);
  run<128, 256, false, 1>("SS one accumulator", dbuf);
  run<128, 64, true, 1>("TS one accumulator", dbuf);
  run<128, 32, true, 1>("TS one accumulator", dbuf);
  run<128, 16, true, 1>("TS one accumulator", dbuf);
  run<128, 128, true, 1>("TS one accumulator", dbuf);
  run<128, 64, true, 2>("TS two accumulators", dbuf);
  run<128, 64, true, 3>("TS three accumulators", dbuf);
  run<128, 32, true, 3>("TS three accumulators", dbuf);
  run<128, 64, false, 2>("SS two accumulators", dbuf);
  run<128, 64, false, 3>("SS three accumulators", dbuf);
  run<64, 24, false, 3>("SS three accumulators", dbuf);
  return 0;
}
