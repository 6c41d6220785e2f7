// Shared helpers of the sucre_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "sucre_b200.h"

namespace sucre {

constexpr int kTile = SUCRE_TILE_PIXELS;  // 32 consecutive flat target pixels = one warp
constexpr unsigned kFull = 0xffffffffu;

// thread-local error text behind sucre_last_error()
int set_error(const char* fmt, ...);
void clear_error();

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("%s: %s", what, cudaGetErrorString(e));
    return 0;
}

#define SUCRE_REQUIRE(cond, ...)                          \
    do {                                                  \
        if (!(cond)) return ::sucre::set_error(__VA_ARGS__); \
    } while (0)

#define SUCRE_CUDA(call)                                                                 \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return ::sucre::set_error(#call ": %s", cudaGetErrorString(e_)); \
    } while (0)

// global tile of local tile k of a band (include/sucre_b200.h)
__host__ __device__ inline int band_tile(const sucre_band& b, int k) {
    return b.first_tile + (k / b.chunk_tiles) * b.stride_tiles + k % b.chunk_tiles;
}

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace sucre
