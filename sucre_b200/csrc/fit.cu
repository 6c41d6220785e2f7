// Stage 2 — per-pixel fit of the underwater image formation model for sm_100a.
//
// One warp owns one tile (32 consecutive target pixels), one lane one pixel.  The tile's observations are a
// contiguous run of 16-byte records {z, I_r, I_g, I_b}; per block (= source view) the matched lanes read
// consecutive records, i.e. one coalesced 128-bit load per lane.  Sweep 1 forms the closed-form J of the
// lane's pixel (sucre.py:66-77), sweep 2 re-reads the same records (L1/L2 hits, the tile was just streamed)
// and accumulates the residual sums that are the gradients of B, beta, gamma with J held constant
// (sucre.py:79-82, 144-145).  Per-thread fp32 sums over one tile are promoted to double per tile, reduced
// with warp shuffles, then one double partial per CTA; a single small CTA finishes the reduction in a fixed
// order and applies torch.optim.Adam's update to the 9 scalars on the device, so 200 iterations need no host
// round trip.
#include "common.cuh"

namespace sucre {

constexpr int kFitThreads = 256;
constexpr int kMaxFitCtas = 2048;
constexpr int kSums = 10;

struct Params9 {
    float B[3], beta[3], gamma[3];
};

__device__ __forceinline__ Params9 load_params(const float* __restrict__ p) {
    Params9 q;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        q.B[c] = p[c];
        q.beta[c] = p[3 + c];
        q.gamma[c] = p[6 + c];
    }
    return q;
}

// Walks the blocks of one tile; F(record) is called for the lanes whose pixel is matched in the block.
template <class F>
__device__ __forceinline__ void for_each_record(const float4* __restrict__ records, const uint32_t* __restrict__ blk_mask,
                                                long long rec, long long b0, int nb, int lane, F&& f) {
    const uint32_t lt = (1u << lane) - 1u;
    for (int j0 = 0; j0 < nb; j0 += 32) {
        const int nj = min(32, nb - j0);
        const uint32_t mload = lane < nj ? __ldg(blk_mask + b0 + j0 + lane) : 0u;  // 32 block masks per coalesced load
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
            const uint32_t m = __shfl_sync(kFull, mload, j);
            if ((m >> lane) & 1u) f(__ldg(records + rec + __popc(m & lt)));
            rec += __popc(m);
        }
    }
}

// sweep 1: closed-form J of this lane's pixel.  0/0 = NaN when the pixel has no observation (sucre.py:77).
__device__ __forceinline__ void closed_form_J(const float4* __restrict__ records, const uint32_t* __restrict__ blk_mask,
                                              long long rec, long long b0, int nb, int lane, const Params9& q, float J[3]) {
    float num[3] = {0.f, 0.f, 0.f}, den[3] = {0.f, 0.f, 0.f};
    for_each_record(records, blk_mask, rec, b0, nb, lane, [&](const float4 r) {
        const float I[3] = {r.y, r.z, r.w};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = expf(-q.beta[c] * r.x);
            const float bs = q.B[c] * (1.0f - expf(-q.gamma[c] * r.x));
            num[c] += (I[c] - bs) * a;
            den[c] += a * a;
        }
    });
#pragma unroll
    for (int c = 0; c < 3; ++c) J[c] = num[c] / den[c];
}

__global__ void __launch_bounds__(kFitThreads)
fit_sums_kernel(const float4* __restrict__ records, const long long* __restrict__ rec_off,
                const long long* __restrict__ blk_off, const uint32_t* __restrict__ blk_mask, int n_tiles,
                const float* __restrict__ params, double* __restrict__ partials) {
    const Params9 q = load_params(params);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_warps = gridDim.x * (kFitThreads / 32);
    double acc[kSums];
#pragma unroll
    for (int i = 0; i < kSums; ++i) acc[i] = 0.0;

    for (int tile = blockIdx.x * (kFitThreads / 32) + warp; tile < n_tiles; tile += n_warps) {
        const long long rec = rec_off[tile], b0 = blk_off[tile];
        const int nb = (int)(blk_off[tile + 1] - b0);
        if (nb == 0) continue;
        float J[3];
        closed_form_J(records, blk_mask, rec, b0, nb, lane, q, J);
        float s[kSums];
#pragma unroll
        for (int i = 0; i < kSums; ++i) s[i] = 0.f;
        for_each_record(records, blk_mask, rec, b0, nb, lane, [&](const float4 r) {
            const float I[3] = {r.y, r.z, r.w};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = expf(-q.beta[c] * r.x);
                const float e = expf(-q.gamma[c] * r.x);
                const float res = I[c] - (J[c] * a + q.B[c] * (1.0f - e));  // sucre.py:81
                s[c] += res * (1.0f - e);
                s[3 + c] += res * J[c] * r.x * a;
                s[6 + c] += res * q.B[c] * r.x * e;
                s[9] += res * res;
            }
        });
#pragma unroll
        for (int i = 0; i < kSums; ++i) acc[i] += (double)s[i];
    }

    // warp tree, then one slot per warp, then one partial row per CTA
    __shared__ double sm[kFitThreads / 32][kSums];
#pragma unroll
    for (int i = 0; i < kSums; ++i) {
        double v = acc[i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kSums) {
        double v = 0.0;
        for (int wi = 0; wi < kFitThreads / 32; ++wi) v += sm[wi][threadIdx.x];
        partials[(size_t)blockIdx.x * kSums + threadIdx.x] = v;
    }
}

// fixed-order reduction of the per-CTA partial rows: warp i sums column i
__device__ __forceinline__ double reduce_column(const double* __restrict__ partials, int n_rows, int col, int lane) {
    double v = 0.0;
    for (int r = lane; r < n_rows; r += 32) v += partials[(size_t)r * kSums + col];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__global__ void __launch_bounds__(kSums * 32)
reduce_partials_kernel(const double* __restrict__ partials, int n_rows, double* __restrict__ sums) {
    const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
    const double v = reduce_column(partials, n_rows, col, lane);
    if (lane == 0) sums[col] = v;
}

// torch.optim.Adam (single-tensor, non-capturable CPU branch the reference runs): fp32 state and params,
// python-float (double) scalars rounded to fp32 where they meet a tensor.
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, int t, double lr) {
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    m = m + (float)(1.0 - b1) * (g - m);
    v = v * (float)b2 + (float)(1.0 - b2) * (g * g);
    const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
    const double step_size = lr / bc1, bc2_sqrt = sqrt(bc2);
    const float denom = sqrtf(v) / (float)bc2_sqrt + (float)eps;
    return p + (float)(-step_size) * (m / denom);
}

// reduces `n_rows` partial rows (n_rows == 1: already reduced sums) and steps the 9 parameters
__global__ void __launch_bounds__(kSums * 32)
adam_step_kernel(const double* __restrict__ partials, int n_rows, long long n_obs, int t, double lr,
                 float* __restrict__ params, float* __restrict__ state, float* __restrict__ history_row) {
    __shared__ double sums[kSums];
    const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
    const double v = reduce_column(partials, n_rows, col, lane);
    if (lane == 0) sums[col] = v;
    __syncthreads();
    if (threadIdx.x < 9) {
        const int i = threadIdx.x;
        // d/dtheta [ sum r^2 / n_obs / 3 ] (sucre.py:145): B: -2 r (1-e), beta: +2 r J z a, gamma: -2 r B z e
        const double sc = 2.0 / (3.0 * (double)n_obs);
        const float g = (float)((i >= 3 && i < 6 ? sc : -sc) * sums[i]);
        const float p = adam_update(params[i], g, state[i], state[9 + i], t, lr);
        params[i] = p;
        if (history_row) history_row[i] = p;
    }
    if (threadIdx.x == 9 && history_row) history_row[9] = (float)sums[9];
}

__global__ void __launch_bounds__(kFitThreads)
write_J_kernel(const float4* __restrict__ records, const long long* __restrict__ rec_off,
               const long long* __restrict__ blk_off, const uint32_t* __restrict__ blk_mask, int n_tiles,
               long long pixels, const float* __restrict__ params, float* __restrict__ Jout) {
    const Params9 q = load_params(params);
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (kFitThreads / 32) + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const long long b0 = blk_off[tile];
    float J[3];
    closed_form_J(records, blk_mask, rec_off[tile], b0, (int)(blk_off[tile + 1] - b0), lane, q, J);
    const long long p = (long long)tile * kTile + lane;
    if (p < pixels) {
        Jout[3 * p + 0] = J[0];
        Jout[3 * p + 1] = J[1];
        Jout[3 * p + 2] = J[2];
    }
}

static int fit_grid() {
    static int ctas = 0;
    if (ctas == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_sums_kernel, kFitThreads, 0) != cudaSuccess || per_sm <= 0)
            per_sm = 4;
        ctas = min(kMaxFitCtas, num_sms() * per_sm);
    }
    return ctas;
}

static int check_store(const float* records, const int64_t* rec_off, const int64_t* blk_off, const uint32_t* blk_mask,
                       int n_tiles, const char* who) {
    SUCRE_REQUIRE(records && rec_off && blk_off && blk_mask, "%s: null pointer", who);
    SUCRE_REQUIRE(n_tiles > 0, "%s: n_tiles = %d", who, n_tiles);
    SUCRE_REQUIRE((reinterpret_cast<uintptr_t>(records) & 15) == 0, "%s: records must be 16-byte aligned", who);
    return 0;
}

}  // namespace sucre

using namespace sucre;

extern "C" size_t sucre_fit_workspace_bytes(void) { return sizeof(double) * kSums * kMaxFitCtas; }

extern "C" int sucre_fit_sums_closed_form(const float* records, const int64_t* rec_off, const int64_t* blk_off,
                                          const uint32_t* blk_mask, int n_tiles, const float* params, double* sums,
                                          void* workspace, void* stream) {
    clear_error();
    if (check_store(records, rec_off, blk_off, blk_mask, n_tiles, "sucre_fit_sums_closed_form")) return 1;
    SUCRE_REQUIRE(params && sums && workspace, "sucre_fit_sums_closed_form: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = fit_grid();
    fit_sums_kernel<<<ctas, kFitThreads, 0, st>>>(reinterpret_cast<const float4*>(records), (const long long*)rec_off,
                                                  (const long long*)blk_off, blk_mask, n_tiles, params, (double*)workspace);
    reduce_partials_kernel<<<1, kSums * 32, 0, st>>>((const double*)workspace, ctas, sums);
    return check_launch("fit_sums_kernel");
}

extern "C" int sucre_adam_step(float* params, float* adam_state, const double* sums, int64_t n_obs, int t, double lr,
                               float* history_row, void* stream) {
    clear_error();
    SUCRE_REQUIRE(params && adam_state && sums, "sucre_adam_step: null pointer");
    SUCRE_REQUIRE(n_obs > 0 && t >= 1, "sucre_adam_step: n_obs = %lld, t = %d", (long long)n_obs, t);
    adam_step_kernel<<<1, kSums * 32, 0, (cudaStream_t)stream>>>(sums, 1, n_obs, t, lr, params, adam_state, history_row);
    return check_launch("adam_step_kernel");
}

extern "C" int sucre_fit_closed_form(const float* records, const int64_t* rec_off, const int64_t* blk_off,
                                     const uint32_t* blk_mask, int n_tiles, int64_t n_obs, float* params,
                                     float* adam_state, int first_step, int num_iter, double lr, float* history,
                                     void* workspace, void* stream) {
    clear_error();
    if (check_store(records, rec_off, blk_off, blk_mask, n_tiles, "sucre_fit_closed_form")) return 1;
    SUCRE_REQUIRE(params && adam_state && workspace, "sucre_fit_closed_form: null pointer");
    SUCRE_REQUIRE(n_obs > 0 && first_step >= 1 && num_iter >= 0, "sucre_fit_closed_form: bad n_obs/first_step/num_iter");
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = fit_grid();
    for (int it = 0; it < num_iter; ++it) {
        fit_sums_kernel<<<ctas, kFitThreads, 0, st>>>(reinterpret_cast<const float4*>(records), (const long long*)rec_off,
                                                      (const long long*)blk_off, blk_mask, n_tiles, params, (double*)workspace);
        adam_step_kernel<<<1, kSums * 32, 0, st>>>((const double*)workspace, ctas, n_obs, first_step + it, lr, params,
                                                   adam_state, history ? history + (size_t)it * kSums : nullptr);
    }
    return check_launch("sucre_fit_closed_form kernels");
}

extern "C" int sucre_fit_write_J(const float* records, const int64_t* rec_off, const int64_t* blk_off,
                                 const uint32_t* blk_mask, int n_tiles, int64_t target_pixels, const float* params,
                                 float* J, void* stream) {
    clear_error();
    if (check_store(records, rec_off, blk_off, blk_mask, n_tiles, "sucre_fit_write_J")) return 1;
    SUCRE_REQUIRE(params && J && target_pixels > 0, "sucre_fit_write_J: bad arguments");
    write_J_kernel<<<(n_tiles + kFitThreads / 32 - 1) / (kFitThreads / 32), kFitThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(records), (const long long*)rec_off, (const long long*)blk_off, blk_mask, n_tiles,
        target_pixels, params, J);
    return check_launch("write_J_kernel");
}
