// Stage 2 — per-pixel fit of the underwater image formation model for sm_100a.
//
// One warp owns one tile (32 consecutive target pixels), one lane one pixel.  The observation store is a stream
// of 16-byte cells, tile after tile; within a tile, segments of up to 15 source views: 2 header cells (32 per-lane
// record counts) then the records {z, I_r, I_g, I_b} LANE-MAJOR, so every lane walks its own contiguous run and
// the inner loop carries no mask / popcount / shuffle addressing at all.
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
// Tiles are statically partitioned over the resident warps by an instruction-cost model (sucre_fit_prepare), so the
// summation order — and therefore every bit of the result — is reproducible run to run.  Inside the Adam loop the
// launches are chained with programmatic dependent launch, and for a target sharded over several GPUs the all-reduce
// of the sums runs inside the last CTA over NVLink peer memory (sucre_fit_sharded).
#include "common.cuh"

namespace sucre {

#ifndef SUCRE_FIT_THREADS
#define SUCRE_FIT_THREADS 512
#endif
constexpr int kFitThreads = SUCRE_FIT_THREADS;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kMaxFitCtas = 2048;
constexpr int kSums = 10;
constexpr float kLog2e = 1.4426950408889634f;

// workspace layout (bytes)
constexpr size_t kWsPartials = 0;                                                   // double[kMaxFitCtas][kSums]
constexpr size_t kWsPartition = kWsPartials + sizeof(double) * kSums * kMaxFitCtas;  // int[kMaxFitCtas*kFitWarps + 1]
constexpr size_t kWsTicket = kWsPartition + sizeof(int) * (kMaxFitCtas * kFitWarps + 1 + 3);  // unsigned, 16-aligned
#ifdef SUCRE_FIT_TIMING  // developer build: %globaltimer at CTA start and at the end of every warp's tile loop (tools/fit_timing.py)
constexpr size_t kWsTiming = kWsTicket + 16;   // u64[kMaxFitCtas] CTA start, u64[kMaxFitCtas*kFitWarps] warp end, u64 last CTA end
constexpr size_t kWsBytes = kWsTiming + 8 * (kMaxFitCtas + kMaxFitCtas * kFitWarps + 1);
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#else
constexpr size_t kWsBytes = kWsTicket + 16;
#endif

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

// ---- per-warp bulk-copy ring --------------------------------------------------------------------------------
// The tiles of one warp are consecutive, so its cells are ONE contiguous byte range in HBM.  The warp streams that
// range through a private shared-memory ring with cp.async.bulk (TMA 1-D copies of 4 KB, kStages slots, one
// mbarrier per slot): HBM latency is covered by the copies in flight instead of by occupancy, and the arithmetic
// reads 16-byte records from shared memory.
#ifndef SUCRE_FIT_CTAS
#define SUCRE_FIT_CTAS 1                            // resident CTAs per SM the kernel is shaped for
#endif
#ifndef SUCRE_FIT_CHUNK
#define SUCRE_FIT_CHUNK 256
#endif
constexpr int kChunkCells = SUCRE_FIT_CHUNK;                 // cells per bulk copy (2 or 4 KB)
#ifndef SUCRE_FIT_STAGES
#define SUCRE_FIT_STAGES 3
#endif
constexpr int kStages = SUCRE_FIT_STAGES;                    // ring slots per warp
constexpr int kRingCells = kChunkCells * kStages;   // 12 KB per warp: 192 KB for the one 16-warp CTA of an SM
constexpr int kSegViews = SUCRE_SEGMENT_VIEWS;
#ifndef SUCRE_FIT_ILP
#define SUCRE_FIT_ILP 2
#endif
#ifndef SUCRE_FIT_OVERSUB
#define SUCRE_FIT_OVERSUB 1
#endif
constexpr int kOversub = SUCRE_FIT_OVERSUB;           // CTAs launched per resident CTA slot (finer static partition)
constexpr int kIlp = SUCRE_FIT_ILP;                   // records of one lane in flight per step
// The first kIlp-1 cells of the ring are mirrored behind its end (a second, tiny bulk copy whenever slot 0 is
// filled), so that the kIlp consecutive records of a step are always at p[0..kIlp-1] and only the walking pointer
// wraps — no per-record modulo in the inner loop.
constexpr int kMirrorCells = kIlp - 1;
constexpr int kRingStride = kRingCells + kMirrorCells;   // cells per warp in shared memory
constexpr size_t kFitSmem = (size_t)kFitWarps * kRingStride * sizeof(float4);
static_assert(kMirrorCells >= 0 && kMirrorCells < kChunkCells, "mirror must be a prefix of one chunk");
constexpr int kSegHeaderCells = SUCRE_SEGMENT_HEADER_CELLS;
static_assert(kSegHeaderCells + 32 * kSegViews <= kRingCells - kChunkCells, "a segment must fit in the ring next to one copy in flight");

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
// Per-pixel statistics of one channel, two per packed accumulator so that one FFMA2 updates both:
//   S12 = (sum D'a, sum a^2)   S34 = (sum D'h, sum a h)   S56 = (sum D'za, sum a za)   S78 = (sum D'zg, sum a zg)
//   S9  = sum D'^2             (a = e^{-beta z}, g = e^{-gamma z}, h = 1 - g, D' the shifted residual)
template <int MODE, bool PRECISE>
struct PixelStats {
    u64 S12[3], S34[3], S56[3], S78[3];
    float S9[3];

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            S12[c] = S34[c] = S56[c] = S78[c] = 0ull;
            S9[c] = 0.f;
        }
    }

    // kbg[c] = (kb, kg) exponent scales, Bc / nJ = B and -Jref of the channel
    __device__ __forceinline__ void add(const float4 r, const u64 kbg[3], const float Bc[3], const float nJ[3]) {
        const float z = r.x;
        const float I[3] = {r.y, r.z, r.w};
        const u64 zz = pk(z, z);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const u64 e = mul2(kbg[c], zz);
            const float a = PRECISE ? expf(lo(e)) : fast_exp2(lo(e));   // e^{-beta z}
            const float g = PRECISE ? expf(hi(e)) : fast_exp2(hi(e));   // e^{-gamma z}
            const float h = 1.0f - g;
            const float D = fmaf(-Bc[c], h, I[c]);                      // I - B (1 - g)
            const float Dp = fmaf(nJ[c], a, D);                         // shifted residual D - Jref a
            const u64 Da = pk(Dp, a);
            S12[c] = fma2(Da, pk(a, a), S12[c]);
            if (MODE == kWriteJ) continue;
            const u64 Dz = mul2(Da, zz);                                // (D' z, a z): every broadcast factor below is a scalar
            S34[c] = fma2(Da, pk(h, h), S34[c]);
            S56[c] = fma2(Dz, pk(a, a), S56[c]);
            S78[c] = fma2(Dz, pk(g, g), S78[c]);
            S9[c] = fmaf(Dp, Dp, S9[c]);
        }
    }
};

struct FitArgs {
    const float4* cells;
    const long long* rec_off;
    const long long* blk_off;
    const long long* seg_off;
    int n_tiles;
    int seg_views;
    long long pixels;
    float* params;         // 9: B, beta, gamma (read at start; written by the last CTA when do_step)
    float* moments;        // 18: Adam state of the 9 scalars
    float* J;              // pixels*3: Jref (closed form, in/out), the J parameter (in/out), or Jref (write-J, may be null)
    float* J_out;          // pixels*3: write-J mode output
    float* J_moments;      // pixels*6: per pixel {m[3], v[3]} (J parameter mode)
    const int* partition;  // per global warp: first tile; [n_warps] = n_tiles
    double* partials;      // gridDim.x rows of kSums
    unsigned* ticket;
#ifdef SUCRE_FIT_TIMING
    unsigned long long* timing;
#endif
    double* sums_out;      // if non-null the last CTA stores the reduced sums here
    float* history_row;    // if non-null: params after the step + cost
    int do_step;           // apply Adam to the 9 scalars in the last CTA
    AdamScalars adam;
    // pixel-band sharding over several GPUs: one-shot all-reduce of the 10 sums over NVLink peer memory, fused
    // into the last CTA (world == 1: single GPU, nothing exchanged)
    int rank, world;
    unsigned epoch;                           // unique, increasing tag of this launch on every rank
    unsigned long long peer[SUCRE_MAX_PEERS]; // peer[p] = address of rank p's exchange buffer (PeerSlot[2][SUCRE_MAX_PEERS])
};

// what rank r leaves in every peer's buffer at [epoch & 1][r]
struct PeerSlot {
    double sums[kSums];
    unsigned flag;
    unsigned pad[3];
};
static_assert(sizeof(PeerSlot) == 96 && sizeof(PeerSlot) * 2 * SUCRE_MAX_PEERS == SUCRE_PEER_BUFFER_BYTES, "peer buffer layout");

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE, bool PRECISE>
__global__ void __launch_bounds__(kFitThreads, SUCRE_FIT_CTAS)
fit_kernel(const __grid_constant__ FitArgs A) {
    extern __shared__ __align__(128) unsigned char fit_smem[];
    __shared__ __align__(8) unsigned long long bars[kFitWarps][kStages];
    // Programmatic dependent launch: let the next iteration's kernel be scheduled as soon as SM resources free up;
    // everything up to griddepcontrol.wait below touches only data that no iteration writes (offsets, partition,
    // cells), so this prologue overlaps the previous iteration's tail.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if defined(SUCRE_FIT_WARP_REMAP) && SUCRE_FIT_WARP_REMAP
    // untested tuning variant (tools/fit_variants.sh): the four warps of one scheduler (warp % 4) take four CONSECUTIVE
    // ranges of the partition, whose whole-tile rounding errors cancel pairwise, instead of ranges 4 apart
    const int gw = blockIdx.x * kFitWarps + ((warp & 3) * (kFitWarps / 4) + (warp >> 2));
#else
    const int gw = blockIdx.x * kFitWarps + warp;
#endif
#ifdef SUCRE_FIT_TIMING
    if (threadIdx.x == 0) A.timing[blockIdx.x] = globaltimer();
#endif
    const float4* ring = reinterpret_cast<const float4*>(fit_smem) + warp * kRingStride;
    const uint8_t* ring_bytes = reinterpret_cast<const uint8_t*>(ring);
    const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(&bars[warp][0]);
    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    double acc[kSums];
#pragma unroll
    for (int i = 0; i < kSums; ++i) acc[i] = 0.0;

    const int t_begin = A.partition[gw], t_end = A.partition[gw + 1];
    const long long C0 = A.rec_off[t_begin] + kSegHeaderCells * A.seg_off[t_begin];
    const int n_cells = (int)(A.rec_off[t_end] + kSegHeaderCells * A.seg_off[t_end] - C0);
    const int n_chunks = (n_cells + kChunkCells - 1) / kChunkCells;
    const float4* src = A.cells + C0;

    // ring bookkeeping, all warp-uniform
    int next_issue = 0, issue_slot = 0;  // next chunk to copy and the slot it goes to
    int wait_slot = 0;                   // slot of the next chunk to wait for
    uint32_t wait_parity = 0;
    int avail = 0;                       // cells that have landed
    int free_at = kChunkCells;           // the oldest slot is recyclable once `pos` reaches this
    int pos = 0, rpos = 0;               // cells consumed; same, modulo the ring size
    auto issue = [&]() {  // lane 0: arm the slot's barrier and start the copy of chunk `next_issue`
        const int first = next_issue * kChunkCells;
        const uint32_t bytes = (uint32_t)min(kChunkCells, n_cells - first) * (uint32_t)sizeof(float4);
        // slot 0 also refreshes the mirror of the ring's first cells behind its end (same barrier)
        const uint32_t mirror = issue_slot == 0 ? min(bytes, (uint32_t)(kMirrorCells * sizeof(float4))) : 0u;
        mbar_expect_tx(bar_s + 8 * issue_slot, bytes + mirror);
        bulk_load(ring_s + issue_slot * kChunkCells * (uint32_t)sizeof(float4), src + first, bytes, bar_s + 8 * issue_slot);
        if (mirror) bulk_load(ring_s + kRingCells * (uint32_t)sizeof(float4), src + first, mirror, bar_s + 8 * issue_slot);
    };
    auto acquire = [&](int upto) {  // cells [0, upto) of the warp's stream have landed
        while (upto > avail) {
            mbar_wait(bar_s + 8 * wait_slot, wait_parity);
            avail += kChunkCells;
            if (++wait_slot == kStages) {
                wait_slot = 0;
                wait_parity ^= 1u;
            }
        }
    };
    for (int c = 0; c < min(kStages, n_chunks); ++c) {
        if (lane == 0) issue();
        ++next_issue;
        issue_slot = issue_slot + 1 == kStages ? 0 : issue_slot + 1;
    }

    // the previous iteration (parameters, J, ticket) must be complete and visible from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const Coef q = load_coef(A.params);
    // per-channel exponent scales packed (beta, gamma): one FMUL2 forms both exponents of a record
    u64 kbg[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) kbg[c] = PRECISE ? pk(-q.beta[c], -q.gamma[c]) : pk(q.kb[c], q.kg[c]);

    long long p_next = (long long)t_begin * kTile + lane;
    float Jnext[3] = {0.f, 0.f, 0.f};
    if (A.J && t_begin < t_end && p_next < A.pixels) {
        Jnext[0] = A.J[3 * p_next], Jnext[1] = A.J[3 * p_next + 1], Jnext[2] = A.J[3 * p_next + 2];
    }
    long long blk_next = t_begin < t_end ? A.blk_off[t_begin + 1] : 0;
    long long blk_cur = t_begin < t_end ? A.blk_off[t_begin] : 0;

#pragma unroll 1
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int nb = (int)(blk_next - blk_cur);
        const long long p = p_next;
        float Jref[3] = {Jnext[0], Jnext[1], Jnext[2]};
        blk_cur = blk_next;
        if (tile + 1 < t_end) {  // prefetch the next tile's extent and reference J
            blk_next = A.blk_off[tile + 2];
            p_next = p + kTile;
            if (A.J && p_next < A.pixels) {
                Jnext[0] = A.J[3 * p_next], Jnext[1] = A.J[3 * p_next + 1], Jnext[2] = A.J[3 * p_next + 2];
            }
        }
        if (MODE == kWriteJ) {
#pragma unroll
            for (int c = 0; c < 3; ++c) Jref[c] = Jref[c] == Jref[c] ? Jref[c] : 0.f;  // a NaN reference is no reference
        }
        PixelStats<MODE, PRECISE> st;
        st.clear();
        int seen = 0;
        if (nb > 0) {  // warp-uniform
            const float nJ[3] = {-Jref[0], -Jref[1], -Jref[2]};
            const int nseg = (nb + A.seg_views - 1) / A.seg_views;
#pragma unroll 1
            for (int s = 0; s < nseg; ++s) {
                acquire(pos + kSegHeaderCells);
                int hcell = rpos + (lane >> 4);
                hcell -= hcell >= kRingCells ? kRingCells : 0;
                const int cnt = ring_bytes[hcell * 16 + (lane & 15)];  // records of this lane's pixel in the segment
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(kFull, incl, o);
                    incl += lane >= o ? up : 0;
                }
                const int n = __shfl_sync(kFull, incl, 31);
                acquire(pos + kSegHeaderCells + n);
                int first = rpos + kSegHeaderCells + (incl - cnt);
                first -= first >= kRingCells ? kRingCells : 0;
                // every lane walks its own run (divergent trip count), two records per step for instruction-level
                // parallelism (their exp / residual chains are independent until the accumulators); the records of a
                // step are contiguous thanks to the mirror cells, only the walking pointer wraps
                {
                    const float4* rp = ring + first;
                    int k = 0;
#pragma unroll 1
                    for (; k + kIlp <= cnt; k += kIlp) {
                        float4 r[kIlp];
#pragma unroll
                        for (int u = 0; u < kIlp; ++u) r[u] = rp[u];
                        rp += kIlp;
                        rp -= rp >= ring + kRingCells ? kRingCells : 0;
#pragma unroll
                        for (int u = 0; u < kIlp; ++u) st.add(r[u], kbg, q.B, nJ);
                    }
#pragma unroll 1
                    for (; k < cnt; ++k) {  // fewer than kIlp left: they cannot reach past the mirror
                        st.add(*rp, kbg, q.B, nJ);
                        ++rp;
                    }
                }
                __syncwarp();
                seen += cnt;
                pos += kSegHeaderCells + n;
                rpos += kSegHeaderCells + n;
                rpos -= rpos >= kRingCells ? kRingCells : 0;
                if (pos >= free_at) {  // every lane is done with the oldest slot(s): refill
                    __syncwarp();
                    while (pos >= free_at) {
                        if (next_issue < n_chunks) {
                            if (lane == 0) issue();
                            ++next_issue;
                            issue_slot = issue_slot + 1 == kStages ? 0 : issue_slot + 1;
                        }
                        free_at += kChunkCells;
                    }
                }
            }
        }
        if (MODE == kWriteJ) {
            if (p < A.pixels) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    A.J_out[3 * p + c] = seen ? Jref[c] + lo(st.S12[c]) / hi(st.S12[c]) : __int_as_float(0x7fc00000);
            }
        } else if (seen) {  // lanes whose pixel has no observation in any kept view contribute nothing and keep their J
            float Jout[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float S1 = lo(st.S12[c]), S3 = lo(st.S34[c]), S5 = lo(st.S56[c]), S7 = lo(st.S78[c]), S9 = st.S9[c];
                float delta = 0.f, rh = S3, rza = S5, rzg = S7, rr = S9;
                if (MODE == kClosedForm) {
                    delta = S1 / hi(st.S12[c]);
                    rh = fmaf(-delta, hi(st.S34[c]), S3);
                    rza = fmaf(-delta, hi(st.S56[c]), S5);
                    rzg = fmaf(-delta, hi(st.S78[c]), S7);
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
    if (MODE == kWriteJ) return;
#ifdef SUCRE_FIT_TIMING
    if (lane == 0) A.timing[kMaxFitCtas + gw] = globaltimer();
#endif

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
    if (A.world > 1) {
        // One-shot all-reduce over NVLink: every rank stores its 10 sums into every peer's buffer (slot [epoch&1]
        // [rank]), publishes them with a system-scope release of the epoch tag, waits for the tags of all ranks in
        // its own buffer, and adds the rows in rank order — the same order on every rank, so all ranks take the
        // identical Adam step without any host or NCCL round trip.  Two parities: a rank can be at most one
        // launch ahead of the slowest reader of its previous message.
        const unsigned par = A.epoch & 1u;
        if (threadIdx.x < A.world * kSums) {
            const int p = threadIdx.x / kSums, i = threadIdx.x % kSums;
            PeerSlot* dst = reinterpret_cast<PeerSlot*>(A.peer[p]) + par * SUCRE_MAX_PEERS + A.rank;
            dst->sums[i] = tot[i];
        }
        __threadfence_system();
        __syncthreads();
        PeerSlot* mine = reinterpret_cast<PeerSlot*>(A.peer[A.rank]) + par * SUCRE_MAX_PEERS;
        if (threadIdx.x < A.world) {
            st_release_sys(&(reinterpret_cast<PeerSlot*>(A.peer[threadIdx.x]) + par * SUCRE_MAX_PEERS + A.rank)->flag, A.epoch);
            while (ld_acquire_sys(&mine[threadIdx.x].flag) != A.epoch) __nanosleep(64);
        }
        __syncthreads();
        if (threadIdx.x < kSums) {
            double v = 0.0;
            for (int p = 0; p < A.world; ++p) v += *reinterpret_cast<volatile double*>(&mine[p].sums[threadIdx.x]);
            tot[threadIdx.x] = v;
        }
        __syncthreads();
    }
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
#ifdef SUCRE_FIT_TIMING
    if (threadIdx.x == 0) A.timing[kMaxFitCtas + kMaxFitCtas * kFitWarps] = globaltimer();
#endif
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

// first tile of every global warp: tiles are split so that every warp gets the same estimated cost.  Cost model
// (instructions, from the ncu source view): ~55 per block (one record step of the slowest lane), ~110 per segment
// (header, scan, ring bookkeeping), ~250 per tile (finalisation, J load/store, double accumulation).
#ifndef SUCRE_COST_BLOCK
#define SUCRE_COST_BLOCK 4
#define SUCRE_COST_SEGMENT 8
#define SUCRE_COST_TILE 18
#endif
__device__ __forceinline__ long long cost_prefix(const long long* __restrict__ blk_off, const long long* __restrict__ seg_off, int t) {
    return SUCRE_COST_BLOCK * blk_off[t] + SUCRE_COST_SEGMENT * seg_off[t] + (long long)SUCRE_COST_TILE * t;
}

__global__ void partition_kernel(const long long* __restrict__ blk_off, const long long* __restrict__ seg_off, int n_tiles,
                                 int n_warps, int* __restrict__ partition) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_warps) return;
    const long long total = cost_prefix(blk_off, seg_off, n_tiles);
    const long long target = (total * w + n_warps - 1) / n_warps;
    int lo = 0, hi = n_tiles;  // smallest t with cost_prefix(t) >= target
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cost_prefix(blk_off, seg_off, mid) >= target) hi = mid; else lo = mid + 1;
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

template <int MODE, bool PRECISE>
static int occupancy() {
    cudaFuncSetAttribute(fit_kernel<MODE, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFitSmem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_kernel<MODE, PRECISE>, kFitThreads, kFitSmem) != cudaSuccess)
        per_sm = 0;
    return per_sm;
}

// persistent grid shared by every mode (the warp->tile partition is computed for it): resident CTAs of the
// most register-hungry instantiation
static int fit_grid() {
    static int ctas = 0;
    if (ctas == 0) {
        int per_sm = min(min(occupancy<kClosedForm, false>(), occupancy<kClosedForm, true>()),
                         min(min(occupancy<kParamJ, false>(), occupancy<kParamJ, true>()),
                             min(occupancy<kWriteJ, false>(), occupancy<kWriteJ, true>())));
        if (per_sm <= 0) per_sm = 1;
        ctas = min(kMaxFitCtas, num_sms() * per_sm * kOversub);
    }
    return ctas;
}

// pdl = true: programmatic dependent launch — only when the preceding kernel in the stream is another fit_kernel of
// the same loop, because the prologue before griddepcontrol.wait reads cells / offsets / partition, which must not
// have been written by the kernel just before (gather_sample, partition_kernel).
template <int MODE>
static void launch_fit(const FitArgs& a, int ctas, cudaStream_t st, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kFitThreads);
    cfg.dynamicSmemBytes = kFitSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // pairs with griddepcontrol.* in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (precise_exp()) cudaLaunchKernelEx(&cfg, fit_kernel<MODE, true>, a);
    else cudaLaunchKernelEx(&cfg, fit_kernel<MODE, false>, a);
}

static int check_store(const sucre_store* s, const char* who) {
    SUCRE_REQUIRE(s != nullptr, "%s: null store", who);
    SUCRE_REQUIRE(s->cells && s->rec_off && s->blk_off && s->seg_off, "%s: null pointer in store", who);
    SUCRE_REQUIRE(s->n_tiles > 0 && s->pixels > 0 && s->pixels <= (int64_t)s->n_tiles * kTile, "%s: bad store sizes", who);
    SUCRE_REQUIRE((reinterpret_cast<uintptr_t>(s->cells) & 15) == 0, "%s: cells must be 16-byte aligned", who);
    SUCRE_REQUIRE(s->record_cells == 1, "%s: this entry point reads {z, I} stores (record_cells == 1), got %d", who, s->record_cells);
    SUCRE_REQUIRE(s->seg_views >= 1 && kSegHeaderCells + 32 * s->seg_views <= kRingCells - kChunkCells,
                  "%s: segments of %d views do not fit the shared-memory ring", who, s->seg_views);
    return 0;
}

static FitArgs base_args(const sucre_store* s, void* workspace) {
    FitArgs a{};
    a.cells = reinterpret_cast<const float4*>(s->cells);
    a.rec_off = (const long long*)s->rec_off;
    a.blk_off = (const long long*)s->blk_off;
    a.seg_off = (const long long*)s->seg_off;
    a.n_tiles = s->n_tiles;
    a.seg_views = s->seg_views;
    a.pixels = s->pixels;
    char* ws = (char*)workspace;
    a.partials = (double*)(ws + kWsPartials);
    a.partition = (const int*)(ws + kWsPartition);
    a.ticket = (unsigned*)(ws + kWsTicket);
#ifdef SUCRE_FIT_TIMING
    a.timing = (unsigned long long*)(ws + kWsTiming);
#endif
    a.rank = 0;
    a.world = 1;
    return a;
}

}  // namespace sucre

using namespace sucre;

extern "C" size_t sucre_fit_workspace_bytes(void) { return kWsBytes; }

extern "C" int sucre_fit_prepare(const sucre_store* store_host, void* workspace, void* stream) {
    clear_error();
    if (check_store(store_host, "sucre_fit_prepare")) return 1;
    SUCRE_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sucre_fit_prepare: workspace must be non-null, 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_warps = fit_grid() * kFitWarps;
    char* ws = (char*)workspace;
    SUCRE_CUDA(cudaMemsetAsync(ws + kWsTicket, 0, 16, st));
    partition_kernel<<<(n_warps + 1 + 255) / 256, 256, 0, st>>>((const long long*)store_host->blk_off, (const long long*)store_host->seg_off,
                                                                 store_host->n_tiles, n_warps, (int*)(ws + kWsPartition));
    return check_launch("partition_kernel");
}

extern "C" int sucre_fit_sums(int mode, const sucre_store* store_host, const float* params, float* J, float* J_moments,
                              int64_t n_obs, int t, double lr, double* sums, void* workspace, void* stream) {
    clear_error();
    if (check_store(store_host, "sucre_fit_sums")) return 1;
    SUCRE_REQUIRE(params && J && sums && workspace, "sucre_fit_sums: null pointer");
    SUCRE_REQUIRE(mode == kClosedForm || (mode == kParamJ && J_moments && n_obs > 0 && t >= 1), "sucre_fit_sums: bad mode/arguments");
    FitArgs a = base_args(store_host, workspace);
    a.params = const_cast<float*>(params);
    a.J = J;
    a.J_moments = J_moments;
    a.sums_out = sums;
    a.do_step = 0;
    if (mode == kParamJ) a.adam = adam_scalars(t, lr, n_obs);
    if (mode == kClosedForm) launch_fit<kClosedForm>(a, fit_grid(), (cudaStream_t)stream, false);
    else launch_fit<kParamJ>(a, fit_grid(), (cudaStream_t)stream, false);
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

static int fit_loop(int mode, const sucre_store* store_host, int64_t n_obs, float* params, float* adam_state, float* J,
                    float* J_moments, int first_step, int num_iter, double lr, float* history, void* workspace,
                    const uint64_t* peers_host, int rank, int world, uint32_t first_epoch, void* stream, const char* who) {
    clear_error();
    if (check_store(store_host, who)) return 1;
    SUCRE_REQUIRE(params && adam_state && J && workspace, "%s: null pointer", who);
    SUCRE_REQUIRE(mode == kClosedForm || (mode == kParamJ && J_moments), "%s: bad mode %d", who, mode);
    SUCRE_REQUIRE(n_obs > 0 && first_step >= 1 && num_iter >= 0, "%s: bad n_obs/first_step/num_iter", who);
    FitArgs a = base_args(store_host, workspace);
    a.params = params;
    a.moments = adam_state;
    a.J = J;
    a.J_moments = J_moments;
    a.do_step = 1;
    if (world > 1) {
        SUCRE_REQUIRE(peers_host && world <= SUCRE_MAX_PEERS && rank >= 0 && rank < world, "%s: bad peer arguments", who);
        a.rank = rank;
        a.world = world;
        for (int p = 0; p < world; ++p) {
            SUCRE_REQUIRE(peers_host[p] != 0 && (peers_host[p] & 15) == 0, "%s: peer buffer %d null or misaligned", who, p);
            a.peer[p] = peers_host[p];
        }
    }
    const int ctas = fit_grid();
    for (int it = 0; it < num_iter; ++it) {
        a.adam = adam_scalars(first_step + it, lr, n_obs);
        a.history_row = history ? history + (size_t)it * kSums : nullptr;
        a.epoch = first_epoch + (uint32_t)it;
        if (mode == kClosedForm) launch_fit<kClosedForm>(a, ctas, (cudaStream_t)stream, it > 0);
        else launch_fit<kParamJ>(a, ctas, (cudaStream_t)stream, it > 0);
    }
    return check_launch(who);
}

extern "C" int sucre_fit(int mode, const sucre_store* store_host, int64_t n_obs, float* params, float* adam_state, float* J,
                         float* J_moments, int first_step, int num_iter, double lr, float* history, void* workspace,
                         void* stream) {
    return fit_loop(mode, store_host, n_obs, params, adam_state, J, J_moments, first_step, num_iter, lr, history, workspace,
                    nullptr, 0, 1, 0, stream, "sucre_fit");
}

extern "C" int sucre_fit_sharded(int mode, const sucre_store* store_host, int64_t n_obs_global, float* params,
                                 float* adam_state, float* J, float* J_moments, int first_step, int num_iter, double lr,
                                 float* history, void* workspace, const uint64_t* peers_host, int rank, int world,
                                 uint32_t first_epoch, void* stream) {
    SUCRE_REQUIRE(first_epoch != 0, "sucre_fit_sharded: epochs start at 1 (0 is the cleared state of a peer buffer)");
    return fit_loop(mode, store_host, n_obs_global, params, adam_state, J, J_moments, first_step, num_iter, lr, history,
                    workspace, peers_host, rank, world, first_epoch, stream, "sucre_fit_sharded");
}

extern "C" int sucre_fit_write_J(const sucre_store* store_host, const float* params, const float* J_ref, float* J,
                                 void* workspace, void* stream) {
    clear_error();
    if (check_store(store_host, "sucre_fit_write_J")) return 1;
    SUCRE_REQUIRE(params && J && workspace, "sucre_fit_write_J: null pointer");
    FitArgs a = base_args(store_host, workspace);
    a.params = const_cast<float*>(params);
    a.J = const_cast<float*>(J_ref);
    a.J_out = J;
    launch_fit<kWriteJ>(a, fit_grid(), (cudaStream_t)stream, false);
    return check_launch("fit_kernel<write J>");
}
