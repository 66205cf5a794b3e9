// Shared helpers for libtdrb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/tdrb200.h"

namespace tdr {

void set_error(const char* fmt, ...);

#define TDR_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            tdr::set_error(__VA_ARGS__);    \
            return TDR_E_INVALID;           \
        }                                   \
    } while (0)

#define TDR_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            tdr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                           __FILE__, __LINE__);                                     \
            return TDR_E_CUDA;                                                      \
        }                                                                           \
    } while (0)

#define TDR_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            tdr::set_error("kernel launch failed: %s (%s:%d)",                      \
                           cudaGetErrorString(_e), __FILE__, __LINE__);             \
            return TDR_E_CUDA;                                                      \
        }                                                                           \
    } while (0)

constexpr int kNumSMs = 148;  // B200

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: used for in-kernel negative sampling.
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        return rounds<10>(c0, c1, c2, c3);
    }
    // Philox4x32-R; R = 7 is the smallest round count of the Random123 paper that passes BigCrush
    template <int R>
    __device__ __forceinline__ uint4 rounds(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

}  // namespace tdr
