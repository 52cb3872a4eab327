// ck_common.cuh -- shared host/device utilities for libckb200 (errors, RNG, small helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/ckb200.h"

namespace ck {

// ---- error plumbing (host) -------------------------------------------------------------
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define CK_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            return ck::fail(CK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                             " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
        }                                                                                   \
    } while (0)

#define CK_CUDA_PTR(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ck::fail(CK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +      \
                                      " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return nullptr;                                                                 \
        }                                                                                   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int num_sms(int device);

// ---- Philox4x32-10 counter-based RNG ---------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    __host__ __device__ Philox(uint64_t key) : k0((uint32_t)key), k1((uint32_t)(key >> 32)) {}
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo) {
        const uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
    }
    __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0, l0, h1, l1;
            mulhilo(0xD2511F53u, c0, h0, l0);
            mulhilo(0xCD9E8D57u, c2, h1, l1);
            const uint32_t n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

__host__ __device__ inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// uniform successor choice of a playout: one Philox block serves four plies (np.random.randint, MCTS.py:141)
struct PhiloxChoice {
    Philox rng;
    uint32_t tag;
    mutable uint32_t r[4];
    template <typename Pos>
    __host__ __device__ int operator()(const Pos &, int ply, int cnt) const {
        if ((ply & 3) == 0) rng((uint32_t)(ply >> 2), 0u, 0u, tag, r);
        return (int)(((uint64_t)r[ply & 3] * (uint64_t)cnt) >> 32);
    }
};

// uniform in (0,1] from two 32-bit words (53 bits)
__host__ __device__ inline double u01(uint32_t a, uint32_t b) {
    const uint64_t v = (((uint64_t)a << 32) | b) >> 11;
    return (double)(v + 1) * (1.0 / 9007199254740992.0);
}

}  // namespace ck
