// Stage 2 — per-pixel fit of the underwater image formation model for sm_100a.
//
// One warp owns one tile (32 consecutive target pixels), one lane one pixel.  The tile's observations are a
// contiguous run of 16-byte records {z, I_r, I_g, I_b}; per block (= source view) the matched lanes read
// consecutive records: one coalesced 128-bit load per lane, four blocks in flight per warp.
//
// Every Adam iteration reads every record exactly ONCE.  The reference makes two passes (update_J, then
// forward/backward with J held constant, sucre.py:141-146); here both come out of one sweep through per-pixel
// sufficient statistics of the SHIFTED residual D' = I - B(1 - e^{-gamma z}) - Jref e^{-beta z}, where Jref is
// the pixel's J of the previous iteration (kept in HBM, 12 B/pixel):
//     S1 = sum D' a   S2 = sum a^2   S3 = sum D' h   S4 = sum a h   S5 = sum D' z a   S6 = sum z a^2
//     S7 = sum D' z g S8 = sum a z g S9 = sum D'^2            (a = e^{-beta z}, g = e^{-gamma z}, h = 1 - g)
//     delta = S1 / S2,  J = Jref + delta                      (closed form, sucre.py:66-77)
//     sum r h = S3 - delta S4, sum r z a = S5 - delta S6, sum r z g = S7 - delta S8, sum r^2 = S9 - delta S1
// with r = I - (J a + B h) the reference's residual (sucre.py:81).  Shifting by Jref keeps every product at
// residual scale, so the subtraction-of-sums above does not cancel (an unshifted one-sweep form loses ~3 digits).
// In the default mode (J is itself an Adam parameter, sucre.py:47-50) Jref IS J, delta = 0, and the pixel's own
// Adam update is applied in the same kernel.
//
// Global sums: per-pixel values are promoted to double per thread, reduced by warp shuffles, one double row per
// CTA; the last CTA to finish (ticket counter) reduces the rows in a fixed order and applies torch.optim.Adam's
// update to the 9 scalars, so an iteration is ONE kernel and 200 iterations need no host round trip.
// Tiles are statically partitioned over the resident warps by block count (sucre_fit_prepare), so the summation
// order — and therefore every bit of the result — is reproducible run to run.
#include "common.cuh"

namespace sucre {

constexpr int kFitThreads = 256;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kMaxFitCtas = 2048;
constexpr int kSums = 10;
constexpr float kLog2e = 1.4426950408889634f;

// workspace layout (bytes)
constexpr size_t kWsPartials = 0;                                                   // double[kMaxFitCtas][kSums]
constexpr size_t kWsPartition = kWsPartials + sizeof(double) * kSums * kMaxFitCtas;  // int[kMaxFitCtas*kFitWarps + 1]
constexpr size_t kWsTicket = kWsPartition + sizeof(int) * (kMaxFitCtas * kFitWarps + 1 + 3);  // unsigned, 16-aligned
constexpr size_t kWsBytes = kWsTicket + 16;

enum FitMode { kClosedForm = 0, kParamJ = 1, kWriteJ = 2 };

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Coef {
    float B[3], kb[3], kg[3];  // kb = -beta log2(e), kg = -gamma log2(e): e^{-beta z} = 2^{kb z}
    float beta[3], gamma[3];
};

__device__ __forceinline__ Coef load_coef(const float* __restrict__ p) {
    Coef q;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        q.B[c] = p[c];
        q.beta[c] = p[3 + c];
        q.gamma[c] = p[6 + c];
        q.kb[c] = -q.beta[c] * kLog2e;
        q.kg[c] = -q.gamma[c] * kLog2e;
    }
    return q;
}

// torch.optim.Adam (single-tensor, non-capturable CPU branch the reference runs): fp32 state and params,
// python-float (double) scalars rounded to fp32 where they meet a tensor.  step_size = lr / (1 - 0.9^t),
// bc2_sqrt = sqrt(1 - 0.999^t) are evaluated in double by the caller.
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float neg_step_size, float bc2_sqrt) {
    m = m + (float)(1.0 - 0.9) * (g - m);                 // exp_avg.lerp_(grad, 1 - beta1)
    v = v * (float)0.999 + (float)(1.0 - 0.999) * (g * g);  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float denom = sqrtf(v) / bc2_sqrt + (float)1e-8;
    return p + neg_step_size * (m / denom);               // param.addcdiv_(exp_avg, denom, value=-step_size)
}

struct AdamScalars {
    float neg_step_size, bc2_sqrt;
    double grad_scale;  // 2 / (3 n_obs): d/dtheta of sum r^2 / n_obs / 3 (sucre.py:145)
};

static AdamScalars adam_scalars(int t, double lr, long long n_obs) {
    const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
    AdamScalars s;
    s.neg_step_size = (float)(-(lr / bc1));
    s.bc2_sqrt = (float)sqrt(bc2);
    s.grad_scale = 2.0 / (3.0 * (double)n_obs);
    return s;
}

template <int MODE, bool PRECISE>
struct PixelStats {
    // closed form: 9 statistics per channel; J parameter: S2, S4, S6, S8 are not needed; write-J: S1, S2 only
    float S1[3], S2[3], S3[3], S4[3], S5[3], S6[3], S7[3], S8[3], S9[3];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < 3; ++c) S1[c] = S2[c] = S3[c] = S4[c] = S5[c] = S6[c] = S7[c] = S8[c] = S9[c] = 0.f;
    }

    __device__ __forceinline__ void add(const float4 r, const Coef& q, const float Jref[3]) {
        const float z = r.x;
        const float I[3] = {r.y, r.z, r.w};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = PRECISE ? expf(-q.beta[c] * z) : fast_exp2(q.kb[c] * z);
            const float g = PRECISE ? expf(-q.gamma[c] * z) : fast_exp2(q.kg[c] * z);
            const float D = fmaf(q.B[c], g, I[c] - q.B[c]);  // I - B (1 - g)
            const float Dp = fmaf(-Jref[c], a, D);           // shifted residual
            S1[c] = fmaf(Dp, a, S1[c]);
            if (MODE != kParamJ) S2[c] = fmaf(a, a, S2[c]);
            if (MODE == kWriteJ) continue;
            const float h = 1.0f - g, za = z * a, zg = z * g;
            S3[c] = fmaf(Dp, h, S3[c]);
            S5[c] = fmaf(Dp, za, S5[c]);
            S7[c] = fmaf(Dp, zg, S7[c]);
            S9[c] = fmaf(Dp, Dp, S9[c]);
            if (MODE == kClosedForm) {
                S4[c] = fmaf(a, h, S4[c]);
                S6[c] = fmaf(a, za, S6[c]);
                S8[c] = fmaf(a, zg, S8[c]);
            }
        }
    }
};

// Streams the records of one tile through `st`.  Four blocks (source views) per step: their masks are
// warp-uniform loads, the four 128-bit record loads are issued before any arithmetic.
template <class Stats>
__device__ __forceinline__ int sweep_tile(const float4* __restrict__ records, const uint32_t* __restrict__ blk_mask,
                                          long long rec, long long b0, int nb, int lane, const Coef& q,
                                          const float Jref[3], Stats& st) {
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t* __restrict__ mk = blk_mask + b0;
    int seen = 0;
    int j = 0;
    for (; j + 4 <= nb; j += 4) {
        const uint32_t m0 = __ldg(mk + j), m1 = __ldg(mk + j + 1), m2 = __ldg(mk + j + 2), m3 = __ldg(mk + j + 3);
        const int n0 = __popc(m0), n1 = __popc(m1), n2 = __popc(m2), n3 = __popc(m3);
        const bool a0 = (m0 >> lane) & 1u, a1 = (m1 >> lane) & 1u, a2 = (m2 >> lane) & 1u, a3 = (m3 >> lane) & 1u;
        float4 r0, r1, r2, r3;
        if (a0) r0 = __ldcs(records + rec + __popc(m0 & lt));
        if (a1) r1 = __ldcs(records + rec + n0 + __popc(m1 & lt));
        if (a2) r2 = __ldcs(records + rec + n0 + n1 + __popc(m2 & lt));
        if (a3) r3 = __ldcs(records + rec + n0 + n1 + n2 + __popc(m3 & lt));
        rec += n0 + n1 + n2 + n3;
        if (a0) st.add(r0, q, Jref);
        if (a1) st.add(r1, q, Jref);
        if (a2) st.add(r2, q, Jref);
        if (a3) st.add(r3, q, Jref);
        seen += (int)a0 + (int)a1 + (int)a2 + (int)a3;
    }
    for (; j < nb; ++j) {
        const uint32_t m = __ldg(mk + j);
        if ((m >> lane) & 1u) {
            st.add(__ldcs(records + rec + __popc(m & lt)), q, Jref);
            ++seen;
        }
        rec += __popc(m);
    }
    return seen;
}

struct FitArgs {
    const float4* records;
    const long long* rec_off;
    const long long* blk_off;
    const uint32_t* blk_mask;
    int n_tiles;
    long long pixels;
    float* params;        // 9: B, beta, gamma (read at start; written by the last CTA when do_step)
    float* moments;       // 18: Adam state of the 9 scalars
    float* J;             // pixels*3: Jref (closed form, in/out) or the J parameter (in/out)
    float* J_moments;     // pixels*6: per pixel {m[3], v[3]} (J parameter mode)
    const int* partition; // per global warp: first tile; [n_warps] = n_tiles
    double* partials;     // gridDim.x rows of kSums
    unsigned* ticket;
    double* sums_out;     // if non-null the last CTA stores the reduced sums here
    float* history_row;   // if non-null: params after the step + cost
    int do_step;          // apply Adam to the 9 scalars in the last CTA
    AdamScalars adam;
};

// ---- per-warp bulk-copy ring --------------------------------------------------------------------------------
// The tiles of one warp are consecutive, so its records are ONE contiguous byte range in HBM.  The warp streams
// that range through a private shared-memory ring with cp.async.bulk (TMA 1-D copies of 4 KB, kStages slots,
// one mbarrier per slot): HBM latency is covered by the copies in flight instead of by occupancy, and the
// arithmetic reads 16-byte records from shared memory.
constexpr int kChunkRecs = 256;                   // records per bulk copy (4 KB)
constexpr int kStages = 3;                        // ring slots per warp
constexpr int kRingRecs = kChunkRecs * kStages;   // 768 records = 12 KB per warp, 96 KB per CTA, 2 CTAs per SM
constexpr size_t kFitSmem = (size_t)kFitWarps * kRingRecs * sizeof(float4);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): two records per instruction ------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) {
    u64 v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ float lo(u64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi(u64 v) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <bool PRECISE>
__device__ __forceinline__ u64 exp2_pair(u64 x) {  // PRECISE: x holds natural-log exponents, else base-2 exponents
    return PRECISE ? pk(expf(lo(x)), expf(hi(x))) : pk(fast_exp2(lo(x)), fast_exp2(hi(x)));
}

// Per-pixel statistics, one packed accumulator per statistic and channel: the low half sums the records of
// even blocks, the high half those of odd blocks (two source views are processed per step).
template <int MODE, bool PRECISE>
struct PairStats {
    u64 S1[3], S2[3], S3[3], S4[3], S5[3], S6[3], S7[3], S8[3], S9[3];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < 3; ++c) S1[c] = S2[c] = S3[c] = S4[c] = S5[c] = S6[c] = S7[c] = S8[c] = S9[c] = 0ull;
    }

    // rA / rB: the lane's record in the even / odd block; w = (1|0, 1|0) says which of the two exist
    __device__ __forceinline__ void add(const float4 rA, const float4 rB, u64 w, const u64 kb[3], const u64 kg[3],
                                        const u64 Bp[3], const u64 Bn[3], const u64 Jn[3]) {
        const u64 z = pk(rA.x, rB.x);
        const u64 I[3] = {pk(rA.y, rB.y), pk(rA.z, rB.z), pk(rA.w, rB.w)};
        const u64 one = pk(1.f, 1.f), neg1 = pk(-1.f, -1.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const u64 a = exp2_pair<PRECISE>(mul2(kb[c], z));   // e^{-beta z}
            const u64 g = exp2_pair<PRECISE>(mul2(kg[c], z));   // e^{-gamma z}
            const u64 D = fma2(Bp[c], g, add2(I[c], Bn[c]));    // I - B (1 - g)
            const u64 Dp = fma2(Jn[c], a, D);                   // shifted residual D - Jref a
            const u64 Dw = mul2(Dp, w);
            S1[c] = fma2(Dw, a, S1[c]);
            if (MODE == kParamJ) {
                const u64 h = fma2(g, neg1, one), za = mul2(z, a), zg = mul2(z, g);
                S3[c] = fma2(Dw, h, S3[c]);
                S5[c] = fma2(Dw, za, S5[c]);
                S7[c] = fma2(Dw, zg, S7[c]);
                S9[c] = fma2(Dw, Dp, S9[c]);
            } else {
                const u64 aw = mul2(a, w);
                S2[c] = fma2(aw, a, S2[c]);
                const u64 h = fma2(g, neg1, one), za = mul2(z, a), zg = mul2(z, g);
                S3[c] = fma2(Dw, h, S3[c]);
                S4[c] = fma2(aw, h, S4[c]);
                S5[c] = fma2(Dw, za, S5[c]);
                S6[c] = fma2(aw, za, S6[c]);
                S7[c] = fma2(Dw, zg, S7[c]);
                S8[c] = fma2(aw, zg, S8[c]);
                S9[c] = fma2(Dw, Dp, S9[c]);
            }
        }
    }
};

__device__ __forceinline__ float sum2(u64 v) { return lo(v) + hi(v); }

template <int MODE, bool PRECISE>
__global__ void __launch_bounds__(kFitThreads, 2)
fit_kernel(const __grid_constant__ FitArgs A) {
    extern __shared__ __align__(128) unsigned char fit_smem[];
    __shared__ __align__(8) unsigned long long bars[kFitWarps][kStages];
    const Coef q = load_coef(A.params);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kFitWarps + warp;
    const float4* ring = reinterpret_cast<const float4*>(fit_smem) + warp * kRingRecs;
    const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(&bars[warp][0]);
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    // packed per-channel constants: exponent scales, +B, -B
    u64 kb[3], kg[3], Bp[3], Bn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        kb[c] = PRECISE ? pk(-q.beta[c], -q.beta[c]) : pk(q.kb[c], q.kb[c]);
        kg[c] = PRECISE ? pk(-q.gamma[c], -q.gamma[c]) : pk(q.kg[c], q.kg[c]);
        Bp[c] = pk(q.B[c], q.B[c]);
        Bn[c] = pk(-q.B[c], -q.B[c]);
    }

    double acc[kSums];
#pragma unroll
    for (int i = 0; i < kSums; ++i) acc[i] = 0.0;

    const int t_begin = A.partition[gw], t_end = A.partition[gw + 1];
    const long long R0 = A.rec_off[t_begin], B0 = A.blk_off[t_begin];
    const int n_rec = (int)(A.rec_off[t_end] - R0), n_blk = (int)(A.blk_off[t_end] - B0);
    const int n_chunks = (n_rec + kChunkRecs - 1) / kChunkRecs;
    const float4* src = A.records + R0;

    // ring bookkeeping, all warp-uniform
    int next_issue = 0, issue_slot = 0;          // next chunk to copy and the slot it goes to
    int wait_slot = 0;                           // slot of the next chunk to wait for
    uint32_t wait_parity = 0;
    int avail = 0;                               // records that have landed
    int free_at = kChunkRecs;                    // the oldest slot is recyclable once `rel` reaches this
    int rel = 0, rpos = 0;                       // records consumed; same, modulo the ring size
    auto issue = [&]() {  // lane 0: arm the slot's barrier and start the copy of chunk `next_issue`
        const int first = next_issue * kChunkRecs;
        const uint32_t bytes = (uint32_t)min(kChunkRecs, n_rec - first) * (uint32_t)sizeof(float4);
        mbar_expect_tx(bar_s + 8 * issue_slot, bytes);
        bulk_load(ring_s + issue_slot * kChunkRecs * (uint32_t)sizeof(float4), src + first, bytes, bar_s + 8 * issue_slot);
    };
    for (int c = 0; c < min(kStages, n_chunks); ++c) {
        if (lane == 0) issue();
        ++next_issue;
        issue_slot = issue_slot + 1 == kStages ? 0 : issue_slot + 1;
    }

    // block masks: lane j of `mcur` holds the mask of block 32*batch + j; `mnext` is the batch after it
    const uint32_t* __restrict__ mk = A.blk_mask + B0;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t mcur = lane < n_blk ? __ldg(mk + lane) : 0u;
    uint32_t mnext = 32 + lane < n_blk ? __ldg(mk + 32 + lane) : 0u;
    int b = 0;  // block index within the warp's range

    long long p_next = (long long)t_begin * kTile + lane;
    float Jnext[3] = {0.f, 0.f, 0.f};
    if (t_begin < t_end && p_next < A.pixels) {
        Jnext[0] = A.J[3 * p_next], Jnext[1] = A.J[3 * p_next + 1], Jnext[2] = A.J[3 * p_next + 2];
    }
    int b_end_next = t_begin < t_end ? (int)(A.blk_off[t_begin + 1] - B0) : 0;

#pragma unroll 1
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int b_end = b_end_next;
        const long long p = p_next;
        const float Jref[3] = {Jnext[0], Jnext[1], Jnext[2]};
        if (tile + 1 < t_end) {  // prefetch the next tile's extent and reference J
            b_end_next = (int)(A.blk_off[tile + 2] - B0);
            p_next = p + kTile;
            if (p_next < A.pixels) {
                Jnext[0] = A.J[3 * p_next], Jnext[1] = A.J[3 * p_next + 1], Jnext[2] = A.J[3 * p_next + 2];
            }
        }
        if (b_end == b) continue;  // warp-uniform
        const u64 Jn[3] = {pk(-Jref[0], -Jref[0]), pk(-Jref[1], -Jref[1]), pk(-Jref[2], -Jref[2])};
        PairStats<MODE, PRECISE> st;
        st.clear();
        int seen = 0;
#pragma unroll 1
        while (b < b_end) {
            // two blocks (source views) per step, unless the pair would straddle a mask batch or the tile end
            const int j = b & 31;
            const bool two = j != 31 && b + 1 < b_end;
            const uint32_t m0 = __shfl_sync(kFull, mcur, j);
            const uint32_t m1x = __shfl_sync(kFull, mcur, (j + 1) & 31);
            const uint32_t m1 = two ? m1x : 0u;
            const int n0 = __popc(m0), n = n0 + __popc(m1);
            while (rel + n > avail) {  // the records of this step must have landed
                mbar_wait(bar_s + 8 * wait_slot, wait_parity);
                avail += kChunkRecs;
                if (++wait_slot == kStages) {
                    wait_slot = 0;
                    wait_parity ^= 1u;
                }
            }
            const uint32_t w0 = (m0 >> lane) & 1u, w1 = (m1 >> lane) & 1u;
            if (w0 | w1) {
                int i0 = rpos + (w0 ? __popc(m0 & lt) : 0);
                int i1 = rpos + (w1 ? n0 + __popc(m1 & lt) : 0);
                i0 -= i0 >= kRingRecs ? kRingRecs : 0;
                i1 -= i1 >= kRingRecs ? kRingRecs : 0;
                st.add(ring[i0], ring[i1], pk((float)w0, (float)w1), kb, kg, Bp, Bn, Jn);
                seen += (int)(w0 + w1);
            }
            rel += n;
            rpos += n;
            rpos -= rpos >= kRingRecs ? kRingRecs : 0;
            b += two ? 2 : 1;
            if (rel >= free_at) {  // every lane is done with the oldest slot: refill it
                __syncwarp();
                if (next_issue < n_chunks) {
                    if (lane == 0) issue();
                    ++next_issue;
                    issue_slot = issue_slot + 1 == kStages ? 0 : issue_slot + 1;
                }
                free_at += kChunkRecs;
            }
            if ((b & 31) < 2 && (b >> 5) != ((b - (two ? 2 : 1)) >> 5)) {  // crossed into the next mask batch
                mcur = mnext;
                const int nb2 = ((b >> 5) + 1) * 32 + lane;
                mnext = nb2 < n_blk ? __ldg(mk + nb2) : 0u;
            }
        }
        if (seen) {  // lanes whose pixel has no observation in any kept view contribute nothing and keep their J
            float Jout[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float S1 = sum2(st.S1[c]), S3 = sum2(st.S3[c]), S5 = sum2(st.S5[c]), S7 = sum2(st.S7[c]), S9 = sum2(st.S9[c]);
                float delta = 0.f, rh = S3, rza = S5, rzg = S7, rr = S9;
                if (MODE == kClosedForm) {
                    delta = S1 / sum2(st.S2[c]);
                    rh = fmaf(-delta, sum2(st.S4[c]), S3);
                    rza = fmaf(-delta, sum2(st.S6[c]), S5);
                    rzg = fmaf(-delta, sum2(st.S8[c]), S7);
                    rr = fmaf(-delta, S1, S9);
                }
                Jout[c] = Jref[c] + delta;
                acc[c] += (double)rh;                    // sum r (1 - e^{-gamma z})
                acc[3 + c] += (double)(Jout[c] * rza);   // sum r J z e^{-beta z}
                acc[6 + c] += (double)(q.B[c] * rzg);    // sum r B z e^{-gamma z}
                acc[9] += (double)rr;                    // sum r^2
                if (MODE == kParamJ) {
                    // dL/dJ = -(2 / 3N) sum r a; the pixel's own Adam step with the pre-step B, beta, gamma (sucre.py:144-148)
                    float* mv = A.J_moments + 6 * p;
                    float m = mv[c], v = mv[3 + c];
                    Jout[c] = adam_update(Jref[c], (float)(-A.adam.grad_scale) * S1, m, v, A.adam.neg_step_size, A.adam.bc2_sqrt);
                    mv[c] = m;
                    mv[3 + c] = v;
                }
            }
            A.J[3 * p + 0] = Jout[0];
            A.J[3 * p + 1] = Jout[1];
            A.J[3 * p + 2] = Jout[2];
        }
    }

    // warp tree -> one slot per warp -> one row per CTA
    __shared__ double sm[kFitWarps][kSums];
    __shared__ unsigned s_ticket;
#pragma unroll
    for (int i = 0; i < kSums; ++i) {
        double v = acc[i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kSums) {
        double v = 0.0;
        for (int wi = 0; wi < kFitWarps; ++wi) v += sm[wi][threadIdx.x];
        A.partials[(size_t)blockIdx.x * kSums + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(A.ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;

    // last CTA: fixed-order reduction of the rows, then the Adam step of the 9 scalars
    __threadfence();
    __shared__ double tot[kSums];
    for (int col = warp; col < kSums; col += kFitWarps) {
        double v = 0.0;
        for (int r = lane; r < (int)gridDim.x; r += 32) v += __ldcg(A.partials + (size_t)r * kSums + col);
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) tot[col] = v;
    }
    __syncthreads();
    if (threadIdx.x < kSums && A.sums_out) A.sums_out[threadIdx.x] = tot[threadIdx.x];
    if (A.do_step && threadIdx.x < 9) {
        const int i = threadIdx.x;
        // B: -2 r (1-g); beta: +2 r J z a; gamma: -2 r B z g   (all times 1 / 3N)
        const float g = (float)((i >= 3 && i < 6 ? A.adam.grad_scale : -A.adam.grad_scale) * tot[i]);
        float m = A.moments[i], v = A.moments[9 + i];
        const float pnew = adam_update(A.params[i], g, m, v, A.adam.neg_step_size, A.adam.bc2_sqrt);
        A.moments[i] = m;
        A.moments[9 + i] = v;
        A.params[i] = pnew;
        if (A.history_row) A.history_row[i] = pnew;
    }
    if (threadIdx.x == 9 && A.history_row) A.history_row[9] = (float)tot[9];
    if (threadIdx.x == 0) *A.ticket = 0u;
}

// Adam step of the 9 scalars from already reduced sums (multi-GPU: after the all-reduce)
__global__ void adam_step_kernel(const double* __restrict__ sums, AdamScalars ad, float* __restrict__ params,
                                 float* __restrict__ moments, float* __restrict__ history_row) {
    const int i = threadIdx.x;
    if (i < 9) {
        const float g = (float)((i >= 3 && i < 6 ? ad.grad_scale : -ad.grad_scale) * sums[i]);
        float m = moments[i], v = moments[9 + i];
        const float pnew = adam_update(params[i], g, m, v, ad.neg_step_size, ad.bc2_sqrt);
        moments[i] = m;
        moments[9 + i] = v;
        params[i] = pnew;
        if (history_row) history_row[i] = pnew;
    }
    if (i == 9 && history_row) history_row[9] = (float)sums[9];
}

// Final update_J (sucre.py:156): J = Jref + sum(D' a) / sum(a^2) for observed pixels, NaN elsewhere (0/0, sucre.py:77)
template <bool PRECISE>
__global__ void __launch_bounds__(kFitThreads)
write_J_kernel(const float4* __restrict__ records, const long long* __restrict__ rec_off,
               const long long* __restrict__ blk_off, const uint32_t* __restrict__ blk_mask, int n_tiles,
               long long pixels, const float* __restrict__ params, const float* __restrict__ Jref_in,
               float* __restrict__ Jout) {
    const Coef q = load_coef(params);
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kFitWarps + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const long long p = (long long)tile * kTile + lane;
    const long long b0 = blk_off[tile];
    const int nb = (int)(blk_off[tile + 1] - b0);
    float Jref[3] = {0.f, 0.f, 0.f};
    if (Jref_in && p < pixels && nb > 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float j = Jref_in[3 * p + c];
            Jref[c] = j == j ? j : 0.f;  // a NaN reference (never observed so far) is no reference
        }
    }
    PixelStats<kWriteJ, PRECISE> st;
    st.clear();
    const int seen = nb > 0 ? sweep_tile(records, blk_mask, rec_off[tile], b0, nb, lane, q, Jref, st) : 0;
    if (p < pixels) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Jout[3 * p + c] = seen ? Jref[c] + st.S1[c] / st.S2[c] : __int_as_float(0x7fc00000);
    }
}

// first tile of every global warp: tiles are split so that every warp gets the same weight sum(blocks + 2)
__global__ void partition_kernel(const long long* __restrict__ blk_off, int n_tiles, int n_warps, int* __restrict__ partition) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_warps) return;
    const long long total = blk_off[n_tiles] + 2LL * n_tiles;
    const long long target = (total * w + n_warps - 1) / n_warps;
    int lo = 0, hi = n_tiles;  // smallest t with weight_prefix(t) >= target
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (blk_off[mid] + 2LL * mid >= target) hi = mid; else lo = mid + 1;
    }
    partition[w] = w == n_warps ? n_tiles : lo;
}

static bool precise_exp() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SUCRE_PRECISE_EXP");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

static int fit_grid() {
    static int ctas = 0;
    if (ctas == 0) {
        int per_sm = 0;
        cudaFuncSetAttribute(fit_kernel<kClosedForm, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFitSmem);
        cudaFuncSetAttribute(fit_kernel<kClosedForm, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFitSmem);
        cudaFuncSetAttribute(fit_kernel<kParamJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFitSmem);
        cudaFuncSetAttribute(fit_kernel<kParamJ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFitSmem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_kernel<kClosedForm, false>, kFitThreads, kFitSmem) != cudaSuccess || per_sm <= 0)
            per_sm = 2;
        ctas = min(kMaxFitCtas, num_sms() * per_sm);
    }
    return ctas;
}

template <int MODE>
static void launch_fit(const FitArgs& a, int ctas, cudaStream_t st) {
    if (precise_exp()) fit_kernel<MODE, true><<<ctas, kFitThreads, kFitSmem, st>>>(a);
    else fit_kernel<MODE, false><<<ctas, kFitThreads, kFitSmem, st>>>(a);
}

static int check_store(const float* records, const int64_t* rec_off, const int64_t* blk_off, const uint32_t* blk_mask,
                       int n_tiles, const char* who) {
    SUCRE_REQUIRE(records && rec_off && blk_off && blk_mask, "%s: null pointer", who);
    SUCRE_REQUIRE(n_tiles > 0, "%s: n_tiles = %d", who, n_tiles);
    SUCRE_REQUIRE((reinterpret_cast<uintptr_t>(records) & 15) == 0, "%s: records must be 16-byte aligned", who);
    return 0;
}

static FitArgs base_args(const float* records, const int64_t* rec_off, const int64_t* blk_off, const uint32_t* blk_mask,
                         int n_tiles, int64_t pixels, void* workspace) {
    FitArgs a{};
    a.records = reinterpret_cast<const float4*>(records);
    a.rec_off = (const long long*)rec_off;
    a.blk_off = (const long long*)blk_off;
    a.blk_mask = blk_mask;
    a.n_tiles = n_tiles;
    a.pixels = pixels;
    char* ws = (char*)workspace;
    a.partials = (double*)(ws + kWsPartials);
    a.partition = (const int*)(ws + kWsPartition);
    a.ticket = (unsigned*)(ws + kWsTicket);
    return a;
}

}  // namespace sucre

using namespace sucre;

extern "C" size_t sucre_fit_workspace_bytes(void) { return kWsBytes; }

extern "C" int sucre_fit_prepare(const int64_t* blk_off, int n_tiles, void* workspace, void* stream) {
    clear_error();
    SUCRE_REQUIRE(blk_off && workspace && n_tiles > 0, "sucre_fit_prepare: bad arguments");
    SUCRE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sucre_fit_prepare: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_warps = fit_grid() * kFitWarps;
    char* ws = (char*)workspace;
    SUCRE_CUDA(cudaMemsetAsync(ws + kWsTicket, 0, 16, st));
    partition_kernel<<<(n_warps + 1 + 255) / 256, 256, 0, st>>>((const long long*)blk_off, n_tiles, n_warps, (int*)(ws + kWsPartition));
    return check_launch("partition_kernel");
}

extern "C" int sucre_fit_sums(int mode, const float* records, const int64_t* rec_off, const int64_t* blk_off,
                              const uint32_t* blk_mask, int n_tiles, int64_t target_pixels, const float* params, float* J,
                              float* J_moments, int64_t n_obs, int t, double lr, double* sums, void* workspace,
                              void* stream) {
    clear_error();
    if (check_store(records, rec_off, blk_off, blk_mask, n_tiles, "sucre_fit_sums")) return 1;
    SUCRE_REQUIRE(params && J && sums && workspace, "sucre_fit_sums: null pointer");
    SUCRE_REQUIRE(mode == kClosedForm || (mode == kParamJ && J_moments && n_obs > 0 && t >= 1), "sucre_fit_sums: bad mode/arguments");
    FitArgs a = base_args(records, rec_off, blk_off, blk_mask, n_tiles, target_pixels, workspace);
    a.params = const_cast<float*>(params);
    a.J = J;
    a.J_moments = J_moments;
    a.sums_out = sums;
    a.do_step = 0;
    if (mode == kParamJ) a.adam = adam_scalars(t, lr, n_obs);
    if (mode == kClosedForm) launch_fit<kClosedForm>(a, fit_grid(), (cudaStream_t)stream);
    else launch_fit<kParamJ>(a, fit_grid(), (cudaStream_t)stream);
    return check_launch("fit_kernel");
}

extern "C" int sucre_adam_step(float* params, float* adam_state, const double* sums, int64_t n_obs, int t, double lr,
                               float* history_row, void* stream) {
    clear_error();
    SUCRE_REQUIRE(params && adam_state && sums, "sucre_adam_step: null pointer");
    SUCRE_REQUIRE(n_obs > 0 && t >= 1, "sucre_adam_step: n_obs = %lld, t = %d", (long long)n_obs, t);
    adam_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, adam_scalars(t, lr, n_obs), params, adam_state, history_row);
    return check_launch("adam_step_kernel");
}

extern "C" int sucre_fit(int mode, const float* records, const int64_t* rec_off, const int64_t* blk_off,
                         const uint32_t* blk_mask, int n_tiles, int64_t target_pixels, int64_t n_obs, float* params,
                         float* adam_state, float* J, float* J_moments, int first_step, int num_iter, double lr,
                         float* history, void* workspace, void* stream) {
    clear_error();
    if (check_store(records, rec_off, blk_off, blk_mask, n_tiles, "sucre_fit")) return 1;
    SUCRE_REQUIRE(params && adam_state && J && workspace, "sucre_fit: null pointer");
    SUCRE_REQUIRE(mode == kClosedForm || (mode == kParamJ && J_moments), "sucre_fit: bad mode %d", mode);
    SUCRE_REQUIRE(n_obs > 0 && first_step >= 1 && num_iter >= 0, "sucre_fit: bad n_obs/first_step/num_iter");
    FitArgs a = base_args(records, rec_off, blk_off, blk_mask, n_tiles, target_pixels, workspace);
    a.params = params;
    a.moments = adam_state;
    a.J = J;
    a.J_moments = J_moments;
    a.do_step = 1;
    const int ctas = fit_grid();
    for (int it = 0; it < num_iter; ++it) {
        a.adam = adam_scalars(first_step + it, lr, n_obs);
        a.history_row = history ? history + (size_t)it * kSums : nullptr;
        if (mode == kClosedForm) launch_fit<kClosedForm>(a, ctas, (cudaStream_t)stream);
        else launch_fit<kParamJ>(a, ctas, (cudaStream_t)stream);
    }
    return check_launch("sucre_fit kernels");
}

extern "C" int sucre_fit_write_J(const float* records, const int64_t* rec_off, const int64_t* blk_off,
                                 const uint32_t* blk_mask, int n_tiles, int64_t target_pixels, const float* params,
                                 const float* J_ref, float* J, void* stream) {
    clear_error();
    if (check_store(records, rec_off, blk_off, blk_mask, n_tiles, "sucre_fit_write_J")) return 1;
    SUCRE_REQUIRE(params && J && target_pixels > 0, "sucre_fit_write_J: bad arguments");
    const int grid = (n_tiles + kFitWarps - 1) / kFitWarps;
    if (precise_exp())
        write_J_kernel<true><<<grid, kFitThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(records), (const long long*)rec_off,
                                                                             (const long long*)blk_off, blk_mask, n_tiles, target_pixels, params, J_ref, J);
    else
        write_J_kernel<false><<<grid, kFitThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(records), (const long long*)rec_off,
                                                                              (const long long*)blk_off, blk_mask, n_tiles, target_pixels, params, J_ref, J);
    return check_launch("write_J_kernel");
}
