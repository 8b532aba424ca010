// Shared helpers for libtg_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tg_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtg_b200 is written for sm_100a (B200) only"
#endif

void tg_set_error(const char* fmt, ...);

#define TG_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      tg_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));          \
      return -2;                                                                     \
    }                                                                                \
  } while (0)

#define TG_REQUIRE(cond, name)                                                       \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      tg_set_error("%s: requirement failed: %s", name, #cond);                       \
      return -1;                                                                     \
    }                                                                                \
  } while (0)

static inline int tg_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

int tg_num_sms();
int tg_max_smem_optin();

__device__ __forceinline__ float tg_act(float v, int act, float slope) {
  if (act == 1) return v > 0.f ? v : 0.f;
  if (act == 2) return v >= 0.f ? v : v * slope;
  if (act == 3) return 1.f / (1.f + __expf(-v));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Philox-4x32-10
struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  __device__ Philox(unsigned long long seed, unsigned long long subseq, unsigned long long offset) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = (uint32_t)offset; ctr[1] = (uint32_t)(offset >> 32);
    ctr[2] = (uint32_t)subseq; ctr[3] = (uint32_t)(subseq >> 32);
  }
  __device__ uint4 next() {
    uint32_t k0 = key[0], k1 = key[1];
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    if (++ctr[0] == 0) ++ctr[1];
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)
