// Stage 2 — per-pixel fit of the underwater image formation model for sm_100a.
//
// One warp owns a run of tiles (32 slots each), one lane one slot = one target pixel (the pixel behind a slot is the
// store's business: sucre_gather_permute deals pixels to slots by observation count; the loop never needs to know).
// The observation store is a stream of ELL ROWS (include/sucre_b200.h): row j of a tile holds, for every lane, the
// j-th observation of its pixel or an all-zero sentinel; a tile has as many rows as its fullest pixel has
// observations.  A row is one coalesced, bank-conflict-free 256-byte access (8-byte records {z, u8 r, g, b}), every
// lane of the warp walks the same rows, and there is nothing to decode: no headers, offsets or masks in the inner loop.
//
// Every Adam iteration reads every row exactly ONCE.  The reference makes two passes (update_J, then
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
// With u8 colour the statistics are kept in 255-scaled units: I = u8 / 255 (loader.py:157) is linear in the byte, so
// D' * 255 = byte - (255 B) h - (255 Jref) a needs no per-record division; the scale is divided out once per pixel
// (J) and once per CTA (sums).
//
// Balance: the rows (plus a fixed per-tile overhead) are split EVENLY over the resident warps by sucre_fit_prepare.
// A boundary may fall inside a tile; the warp that gets the tail of such a tile evaluates it first, parks the
// tile's partial statistics in shared memory, and the neighbouring warp of the same CTA that owns the tile's head
// adds them before finalising the pixel — so the per-warp work differs by one row at most, whatever the tile sizes,
// and the summation order stays fixed (bit-reproducible results run to run).
//
// Global sums and the loop: a thread's partial sums are fp32 over its handful of tiles, then double: warp shuffles, one
// row of ten doubles per CTA.  The grid (one CTA per SM) is resident and runs ALL iterations of a call: every CTA
// publishes its row as tagged 8-byte words, polls all rows, reduces them in a fixed order and applies
// torch.optim.Adam's update to its own shared-memory copy of the 9 scalars — 200 iterations are one launch, with no
// host round trip and no CTA waiting for another one's result.  For a target sharded over several GPUs the all-reduce
// of the rank totals runs inside the same kernel over NVLink peer memory (sucre_fit_sharded).
#include "common.cuh"

namespace sucre {

#ifndef SUCRE_FIT_THREADS
#define SUCRE_FIT_THREADS 512
#endif
constexpr int kFitThreads = SUCRE_FIT_THREADS;
constexpr int kFitWarps = kFitThreads / 32;
constexpr int kMaxFitCtas = 1024;
constexpr int kSums = 10;
constexpr int kStats = 27;     // per-pixel statistics: 9 per channel
constexpr float kLog2e = 1.4426950408889634f;
#ifndef SUCRE_FIT_TILE_COST
#define SUCRE_FIT_TILE_COST 3  // per-tile overhead (J load/store, finalisation, double accumulation) in row-equivalents
#endif
constexpr int kTileCost = SUCRE_FIT_TILE_COST;
#ifndef SUCRE_FIT_UNROLL
#define SUCRE_FIT_UNROLL 4
#endif
constexpr int kRowLoopUnroll = SUCRE_FIT_UNROLL;   // two-row steps per trip of the row loop: unrolled, the accumulators stop
                                                   // rotating through registers at every back-edge (98 -> 91.5 instructions per step)

// workspace layout (bytes)
constexpr int kLLWords = 2 * kSums;   // a row of ten doubles as tagged 8-byte words {tag : 32 | half a double : 32}
constexpr size_t kWsPartials = 0;                                                     // u64[2][kMaxFitCtas][kLLWords] (tag parity); the light kernels keep plain double rows here
constexpr size_t kWsPartRow = kWsPartials + sizeof(unsigned long long) * kLLWords * kMaxFitCtas * 2;  // long long[kMaxFitCtas*kFitWarps + 1]
constexpr size_t kWsPartTile = kWsPartRow + sizeof(long long) * (kMaxFitCtas * kFitWarps + 2);  // int[kMaxFitCtas*kFitWarps + 1]
constexpr size_t kWsTicket = kWsPartTile + sizeof(int) * (kMaxFitCtas * kFitWarps + 4);  // unsigned [4]: unused, status bits, tag base, unused; 16-aligned
constexpr int kMaxLoopIters = 1024;   // Adam iterations per launch of the persistent loop (longer runs are split)
constexpr size_t kWsAdamTab = kWsTicket + 16;                        // AdamScalars[kMaxLoopIters]
constexpr size_t kWsBytes = kWsAdamTab + 16 * kMaxLoopIters;
static_assert(kWsPartRow % 8 == 0 && kWsPartTile % 8 == 0 && kWsTicket % 16 == 0 && kWsAdamTab % 16 == 0, "workspace alignment");

__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

enum FitMode { kClosedForm = 0, kParamJ = 1, kWriteJ = 2 };

// Developer build only (-DSUCRE_FIT_TRACE, tools/fit_trace.py): globaltimer stamps of the phases of an iteration, per CTA,
// and of the end of every warp's sweep, for the first kTraceIters iterations of a launch.
#if defined(SUCRE_FIT_TRACE)
constexpr int kTraceIters = 64, kTraceCtas = 160, kTraceStamps = 8;
__device__ unsigned long long g_trace[kTraceIters][kTraceCtas][kTraceStamps];
__device__ unsigned long long g_trace_warp[kTraceIters][kTraceCtas][SUCRE_FIT_THREADS / 32];
#define SUCRE_TRACE(slot)                                                                                   \
    do {                                                                                                    \
        if (MODE != kWriteJ && threadIdx.x == 0 && it < kTraceIters && blockIdx.x < kTraceCtas) g_trace[it][blockIdx.x][slot] = globaltimer(); \
    } while (0)
#define SUCRE_TRACE_WARP()                                                                                  \
    do {                                                                                                    \
        if (MODE != kWriteJ && lane == 0 && it < kTraceIters && blockIdx.x < kTraceCtas) g_trace_warp[it][blockIdx.x][warp] = globaltimer(); \
    } while (0)
#else
#define SUCRE_TRACE(slot) do {} while (0)
#define SUCRE_TRACE_WARP() do {} while (0)
#endif

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
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

static_assert(sizeof(AdamScalars) == 16, "Adam table entry");

static AdamScalars adam_scalars(int t, double lr, long long n_obs) {
    const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
    AdamScalars s;
    s.neg_step_size = (float)(-(lr / bc1));
    s.bc2_sqrt = (float)sqrt(bc2);
    s.grad_scale = 2.0 / (3.0 * (double)n_obs);
    return s;
}

// ---- per-warp bulk-copy ring --------------------------------------------------------------------------------
// The rows of one warp are ONE contiguous byte range in HBM.  The warp streams that range through a private
// shared-memory ring with cp.async.bulk (TMA 1-D copies of kChunkBytes, kStages slots, one mbarrier per slot): HBM
// latency is covered by the copies in flight instead of by occupancy, and the arithmetic reads one record per lane
// and row from shared memory.  A slot is refilled as soon as the walk has left it.
#ifndef SUCRE_FIT_CHUNK_BYTES
#define SUCRE_FIT_CHUNK_BYTES 4096
#endif
#ifndef SUCRE_FIT_STAGES
#define SUCRE_FIT_STAGES 2
#endif
constexpr int kChunkBytes = SUCRE_FIT_CHUNK_BYTES;   // 16 rows of 8-byte records, 8 rows of 16-byte records
constexpr int kStages = SUCRE_FIT_STAGES;
constexpr int kRingBytes = kChunkBytes * kStages;    // 8 KB per warp; a power of two, so a row offset wraps with one AND
constexpr int kChunkBytesWriteJ = 1024;              // the memory-bound write-J sweep: 8 slots of 1 KB
constexpr size_t kParkBytes = (size_t)kFitWarps * kStats * 32 * sizeof(float);   // 54 KB: parked statistics of split tiles
constexpr size_t kFitSmem = (size_t)kFitWarps * kRingBytes + kParkBytes;
static_assert(kChunkBytes % 512 == 0, "a chunk must hold whole rows of both record sizes");
static_assert((kRingBytes & (kRingBytes - 1)) == 0 && kStages >= 2, "the ring size must be a power of two, at least two slots");
static_assert(kFitSmem <= 225 * 1024, "shared memory budget");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): two statistics per instruction ---------------
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

// one record as the arithmetic wants it: range z and the three colour values in the store's units
// (SUCRE_REC_Z_U8: the bytes as exact floats 0..255; SUCRE_REC_Z_F32: I in [0,1]).  For u8 colour, x holds the
// "magic" floats 2^23 + byte (byte | 0x4B000000, one PRMT each, ALU pipe) and kBias = -2^23 is added — exactly —
// where the value is first used, packed for red + green.
struct Obs {
    float z, x[3];
};
template <int REC> struct RecTraits;
template <> struct RecTraits<SUCRE_REC_Z_U8> {
    typedef uint2 raw;
    static constexpr int kBytes = 8;
    static constexpr float kScale = 255.0f;
    static constexpr float kBias = -8388608.0f;
    static __device__ __forceinline__ float range(const raw q) { return __uint_as_float(q.x); }
    static __device__ __forceinline__ Obs unpack(const raw q) {
        Obs o;
        o.z = __uint_as_float(q.x);
#pragma unroll
        for (int c = 0; c < 3; ++c) o.x[c] = __uint_as_float(__byte_perm(q.y, 0x4B000000u, 0x7650u + c));
        return o;
    }
};
template <> struct RecTraits<SUCRE_REC_Z_F32> {
    typedef float4 raw;
    static constexpr int kBytes = 16;
    static constexpr float kScale = 1.0f;
    static constexpr float kBias = 0.0f;
    static __device__ __forceinline__ float range(const raw q) { return q.x; }
    static __device__ __forceinline__ Obs unpack(const raw q) {
        Obs o;
        o.z = q.x, o.x[0] = q.y, o.x[1] = q.z, o.x[2] = q.w;
        return o;
    }
};

// Per-pixel statistics, nine per channel (k = 0..8: S1 .. S9 of the header comment), all driven by fma.rn.f32x2:
//   red + green  CHANNEL-packed: every quantity of the pair of channels lives in one 64-bit register — exponents,
//                a, g, h, D, D' and each of the nine sums — so one packed instruction serves both channels of a row;
//   blue         ROW-packed: the two rows of a step share the registers instead (even rows in the low half, odd
//                rows in the high half; the halves are added when the tile is finalised).
// Per step of two rows: 54 packed arithmetic instructions + 12 MUFU.EX2 for 6 channel-records.
template <int MODE, bool PRECISE, int REC>
struct PixelStats {
    u64 P[9];   // (red, green) of S1 .. S9
    u64 Q[9];   // blue: (even rows, odd rows) of S1 .. S9

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < 9; ++k) P[k] = Q[k] = 0ull;
    }

    // statistic k of channel c
    __device__ __forceinline__ float get(int c, int k) const {
        if (c == 0) return lo(P[k]);
        if (c == 1) return hi(P[k]);
        return lo(Q[k]) + hi(Q[k]);
    }

    // Constants of the sweep, in the store's units: exponent scales (kb, kg), -B, and -Jref per tile.
    struct Consts {
        u64 kb_rg, kg_rg, nB_rg;   // (kb_r, kb_g), (kg_r, kg_g), (-B_r, -B_g)
        float kb_b, kg_b, nB_b;
    };

    // The nine updates of one packed pair (D', a, g, h given; zz = the ranges).
    static __device__ __forceinline__ void accumulate(u64 (&S)[9], const u64 Dp, const u64 a, const u64 g, const u64 h, const u64 zz) {
        S[0] = fma2(Dp, a, S[0]);
        S[1] = fma2(a, a, S[1]);
        if (MODE != kWriteJ) {
            S[2] = fma2(Dp, h, S[2]);
            S[3] = fma2(a, h, S[3]);
#if !defined(SUCRE_EXPERIMENT_FEWSTATS)   // timing experiment only (wrong results): a third of the FMA work removed
            const u64 Dz = mul2(Dp, zz), az = mul2(a, zz);
            S[4] = fma2(Dz, a, S[4]);
            S[5] = fma2(az, a, S[5]);
            S[6] = fma2(Dz, g, S[6]);
            S[7] = fma2(az, g, S[7]);
#endif
            S[8] = fma2(Dp, Dp, S[8]);
        }
    }

    static __device__ __forceinline__ u64 exp2_pair(const u64 e) {
#if defined(SUCRE_EXPERIMENT_NOEXP)      // timing experiment only (wrong results): no MUFU at all
        return add2(e, pk(1.0f, 1.0f));
#elif defined(SUCRE_EXPERIMENT_HALFEXP)  // timing experiment only: half of the MUFUs
        return pk(fast_exp2(lo(e)), hi(e) + 1.0f);
#else
        return pk(PRECISE ? expf(lo(e)) : fast_exp2(lo(e)), PRECISE ? expf(hi(e)) : fast_exp2(hi(e)));
#endif
    }

    // One step: rows r0 and r1 of this lane's column.  Either may be a sentinel (z == 0, colour 0): a is multiplied by
    // w = (z != 0), and since z == 0 gives g = 1, h = 0, D = 0, the masked a makes D' = D - Jref a vanish too — every
    // statistic carries a factor D' or a, so a sentinel adds exactly nothing, without a branch.  For a real record
    // w = 1 and nothing changes.  (An odd last row is a step whose second row is an all-zero record.)
    __device__ __forceinline__ void add(const Obs r0, const Obs r1, const Consts& k, const u64 nJ_rg, const float nJ_b) {
        const u64 one = pk(1.0f, 1.0f), neg = pk(-1.0f, -1.0f);
        const u64 bias = pk(RecTraits<REC>::kBias, RecTraits<REC>::kBias);
        const float w0 = r0.z != 0.0f ? 1.0f : 0.0f, w1 = r1.z != 0.0f ? 1.0f : 0.0f;
        // red + green, row by row
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const Obs& r = i == 0 ? r0 : r1;
            const u64 zz = pk(r.z, r.z);
            const float w = i == 0 ? w0 : w1;
            const u64 a = mul2(exp2_pair(mul2(k.kb_rg, zz)), pk(w, w));
            const u64 g = exp2_pair(mul2(k.kg_rg, zz));
            const u64 h = fma2(g, neg, one);                               // 1 - g
            u64 x = pk(r.x[0], r.x[1]);
            if (RecTraits<REC>::kBias != 0.0f) x = add2(x, bias);
            const u64 D = fma2(k.nB_rg, h, x);                             // I - B (1 - g)
            const u64 Dp = fma2(nJ_rg, a, D);                              // shifted residual D - Jref a
            accumulate(P, Dp, a, g, h, zz);
        }
        // blue, both rows at once
        {
            const u64 zz = pk(r0.z, r1.z);
            const u64 a = mul2(exp2_pair(mul2(zz, pk(k.kb_b, k.kb_b))), pk(w0, w1));
            const u64 g = exp2_pair(mul2(zz, pk(k.kg_b, k.kg_b)));
            const u64 h = fma2(g, neg, one);
            u64 x = pk(r0.x[2], r1.x[2]);
            if (RecTraits<REC>::kBias != 0.0f) x = add2(x, bias);
            const u64 D = fma2(pk(k.nB_b, k.nB_b), h, x);
            const u64 Dp = fma2(pk(nJ_b, nJ_b), a, D);
            accumulate(Q, Dp, a, g, h, zz);
        }
    }

    // A single row (the odd last row of a tile): red + green packed as in add(), blue in scalar arithmetic on the
    // even-row halves of Q — about half the instructions of a step whose second row would be an all-zero record.
    __device__ __forceinline__ void add_one(const Obs r, const Consts& k, const u64 nJ_rg, const float nJ_b) {
        const u64 one = pk(1.0f, 1.0f), neg = pk(-1.0f, -1.0f);
        const float w = r.z != 0.0f ? 1.0f : 0.0f;
        {
            const u64 zz = pk(r.z, r.z);
            const u64 a = mul2(exp2_pair(mul2(k.kb_rg, zz)), pk(w, w));
            const u64 g = exp2_pair(mul2(k.kg_rg, zz));
            const u64 h = fma2(g, neg, one);
            u64 x = pk(r.x[0], r.x[1]);
            if (RecTraits<REC>::kBias != 0.0f) x = add2(x, pk(RecTraits<REC>::kBias, RecTraits<REC>::kBias));
            const u64 D = fma2(k.nB_rg, h, x);
            const u64 Dp = fma2(nJ_rg, a, D);
            accumulate(P, Dp, a, g, h, zz);
        }
        {
            const float z = r.z;
            const float a = (PRECISE ? expf(k.kb_b * z) : fast_exp2(k.kb_b * z)) * w;
            const float g = PRECISE ? expf(k.kg_b * z) : fast_exp2(k.kg_b * z);
            const float h = 1.0f - g;
            const float x = r.x[2] + RecTraits<REC>::kBias;
            const float Dp = fmaf(nJ_b, a, fmaf(k.nB_b, h, x));
            float q[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) q[i] = lo(Q[i]);
            q[0] = fmaf(Dp, a, q[0]);
            q[1] = fmaf(a, a, q[1]);
            if (MODE != kWriteJ) {
                const float Dz = Dp * z, az = a * z;
                q[2] = fmaf(Dp, h, q[2]);
                q[3] = fmaf(a, h, q[3]);
                q[4] = fmaf(Dz, a, q[4]);
                q[5] = fmaf(az, a, q[5]);
                q[6] = fmaf(Dz, g, q[6]);
                q[7] = fmaf(az, g, q[7]);
                q[8] = fmaf(Dp, Dp, q[8]);
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) Q[i] = pk(q[i], hi(Q[i]));
        }
    }

    // parked form: kStats floats per lane, [9 * c + k][lane]
    __device__ __forceinline__ void park(float* slot, int lane) const {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < 9; ++k) slot[(9 * c + k) * 32 + lane] = get(c, k);
    }
    __device__ __forceinline__ void add_parked(const float* slot, int lane) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            P[k] = add2(P[k], pk(slot[k * 32 + lane], slot[(9 + k) * 32 + lane]));
            Q[k] = add2(Q[k], pk(slot[(18 + k) * 32 + lane], 0.f));
        }
    }
};

struct FitArgs {
    const unsigned char* cells;
    const long long* row_off;
    int n_tiles;
    long long pixels;          // entries of the slot-ordered J arrays (J, J_moments)
    const int* pix;            // write-J: slot -> entry of J_out (nullptr: the identity, bounded by out_pixels)
    long long out_pixels;
    float* params;         // 9: B, beta, gamma (read when the launch starts; written by CTA 0 when it ends, if do_step)
    float* moments;        // 18: Adam state of the 9 scalars
    float* J;              // pixels*3: Jref (closed form, in/out), the J parameter (in/out), or Jref (write-J, may be null)
    float* J_out;          // pixels*3: write-J mode output
    float* J_moments;      // pixels*6: per pixel {m[3], v[3]} (J parameter mode)
    const long long* part_row;  // per global warp: first row of its stream; [n_warps] = rows of the store
    const int* part_tile;       // per global warp: first tile it owns (finalises); [n_warps] = n_tiles
    unsigned long long* rows;   // [2][kMaxFitCtas][kLLWords] tagged words: the CTAs' rows of partial sums
    unsigned* ticket;      // ticket[1] = status bits, ticket[2] = tag base (iterations run on this workspace so far)
    double* sums_out;      // if non-null CTA 0 stores the reduced sums here
    float* history;        // if non-null: num_iter rows of {params after the step [9], cost}
    int do_step;           // apply Adam to the 9 scalars (every CTA, on its own copy)
    int num_iter;          // iterations this launch runs (the grid stays resident and meets at a flag between them)
    const AdamScalars* adam_tab;   // num_iter entries (device)
    // pixel-band sharding over several GPUs: one-shot all-reduce of the 10 sums over NVLink peer memory, fused
    // into the loop (world == 1: single GPU, nothing exchanged)
    int rank, world;
    unsigned epoch;                           // tag of the first iteration; unique and increasing on every rank
    unsigned long long peer[SUCRE_MAX_PEERS]; // peer[p] = address of rank p's exchange buffer (PeerSlot[2][SUCRE_MAX_PEERS])
};

// What rank r leaves in every peer's buffer at [epoch & 1][r]: the 10 sums as 20 words {epoch : 32 | half of a double : 32}.
// An aligned 8-byte store is atomic, so a word whose tag is the current epoch carries valid data: no fence, no
// separate flag, one NVLink store latency per exchange (the layout of NCCL's LL protocol).
struct PeerSlot {
    unsigned long long w[kLLWords];
};
static_assert(sizeof(PeerSlot) == 160 && sizeof(PeerSlot) * 2 * SUCRE_MAX_PEERS == SUCRE_PEER_BUFFER_BYTES, "peer buffer layout");

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ld_volatile_pair(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// Warp-wide: the sum, over the CTAs' rows, of the double whose two tagged words start at `col_words` in every row.  All of
// a lane's loads are in flight together and are re-issued until each carries the tag: one L2 round trip after the last
// row lands, not one per row.  Rows are added in a fixed order.  (Not inlined: its registers must not weigh on the
// allocation of the sweep's inner loop.)
__device__ __noinline__ double gather_column(const unsigned long long* col_words, unsigned tag, int lane, int n_ctas) {
    constexpr int kDepth = 5;   // rows per lane and pass: 160 CTAs per pass
    double v = 0.0;
    for (int r0 = lane; r0 < n_ctas; r0 += 32 * kDepth) {
        unsigned long long lo[kDepth], hi[kDepth];
        bool ok;
        do {
            ok = true;
#pragma unroll
            for (int k = 0; k < kDepth; ++k) {
                const int r = r0 + 32 * k;
                if (r < n_ctas) {
                    ld_volatile_pair(col_words + (size_t)r * kLLWords, lo[k], hi[k]);
                    ok = ok && (unsigned)(lo[k] >> 32) == tag && (unsigned)(hi[k] >> 32) == tag;
                }
            }
        } while (!ok);
#pragma unroll
        for (int k = 0; k < kDepth; ++k)
            if (r0 + 32 * k < n_ctas) v += __longlong_as_double((long long)((hi[k] << 32) | (lo[k] & 0xffffffffull)));
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// One launch = num_iter Adam iterations (MODE kWriteJ: one sweep).  The grid (one CTA per SM, all resident) stays on the
// machine: after an iteration every CTA publishes its row of partial sums, waits until all rows are there, reduces them
// itself (with the other GPUs' rows if the target is sharded) and takes the Adam step on its own copy of the state,
// while the bulk copies of the next iteration's first rows are already in flight.
template <int MODE, int REC, bool PRECISE>
__global__ void __launch_bounds__(kFitThreads, 1)
fit_kernel(const __grid_constant__ FitArgs A) {
    typedef RecTraits<REC> RT;
    typedef typename RT::raw Raw;
    constexpr int kRowBytes = 32 * RT::kBytes;
    // The Adam sweep is bound by arithmetic: few, large copies keep the ring bookkeeping small.  The write-J sweep does a
    // quarter of the arithmetic and is bound by memory: it wants more copies in flight, so it cuts the same ring finer.
    constexpr int kChunk = MODE == kWriteJ ? kChunkBytesWriteJ : kChunkBytes;
    constexpr int kSlots = kRingBytes / kChunk;
    constexpr int CR = kChunk / kRowBytes;        // rows per chunk
    constexpr float kScale = RT::kScale;
    constexpr float kInv = 1.0f / kScale;
    static_assert(CR >= 2, "a chunk must hold at least two rows");

    extern __shared__ __align__(128) unsigned char fit_smem[];
    __shared__ __align__(8) unsigned long long bars[kFitWarps][kRingBytes / kChunkBytesWriteJ];
    __shared__ __align__(8) unsigned long long park_bar[kFitWarps];
    __shared__ double sm[kFitWarps][kSums];
    __shared__ double tot[kSums];
    __shared__ unsigned ll_half[SUCRE_MAX_PEERS][kLLWords];
    __shared__ float s_state[27];   // this CTA's copy of the 9 parameters and their Adam moments (m[9], v[9])
    __shared__ AdamScalars s_ad;
    __shared__ unsigned s_tag_base; // iterations run on this workspace before this launch (CTA 0 moves it on when the launch ends)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * kFitWarps + warp;
    const unsigned char* ring = fit_smem + (size_t)warp * kRingBytes;
    float* const park_all = reinterpret_cast<float*>(fit_smem + (size_t)kFitWarps * kRingBytes);
    const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(&bars[warp][0]);
    if (lane == 0) {
        for (int s = 0; s < kSlots; ++s) mbar_init(bar_s + 8 * s, 1);
        mbar_init(smem_u32(&park_bar[warp]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 32 && MODE != kWriteJ) s_tag_base = A.ticket[2];
    if (threadIdx.x < 27) {
        const bool moments = threadIdx.x >= 9 && A.moments != nullptr;
        s_state[threadIdx.x] = threadIdx.x < 9 ? A.params[threadIdx.x] : (moments ? A.moments[threadIdx.x - 9] : 0.f);
    }
    __syncthreads();  // the park barrier of a warp is waited on by its neighbour; s_state is read by everybody

    const long long B0 = A.part_row[gw], B1 = A.part_row[gw + 1];
    const int T0 = A.part_tile[gw], T1 = A.part_tile[gw + 1];
    const int n_rows_w = (int)(B1 - B0);
    const int n_chunks = (n_rows_w + CR - 1) / CR;
    const unsigned char* src = A.cells + B0 * kRowBytes;
    // work items: [the tail of tile T0-1, owned by the previous warp,] then the owned tiles T0 .. T1-1, the last of
    // which may continue in the next warp's stream
    const bool has_head = T0 > 0 && A.row_off[T0] > B0;
    const int t_first = T0 - (has_head ? 1 : 0);

    // ring bookkeeping, all warp-uniform (stream rows are relative to B0).  Slots and barrier phases carry over from
    // one iteration to the next: an iteration ends with every issued chunk consumed, so the next one starts at
    // whatever slot comes next.
    int next_issue = 0, issue_slot = 0;  // next chunk to copy and the slot it goes to
    int wait_slot = 0;                   // slot of the next chunk to wait for
    uint32_t wait_parity = 0;
    int avail = 0;                       // rows that have landed
    int release_at = CR;                 // the oldest slot is recyclable once `pos` reaches this
    int pos = 0, roff = 0;               // rows consumed; byte offset of row `pos` in the ring
    auto issue = [&]() {  // lane 0: arm the slot's barrier and start the copy of chunk `next_issue`
        const int first = next_issue * CR;
        const uint32_t bytes = (uint32_t)min(CR, n_rows_w - first) * (uint32_t)kRowBytes;
        mbar_expect_tx(bar_s + 8 * issue_slot, bytes);
        bulk_load(ring_s + issue_slot * (uint32_t)kChunk, src + (size_t)first * kRowBytes, bytes, bar_s + 8 * issue_slot);
    };
    auto advance_issue = [&]() {
        if (lane == 0) issue();
        ++next_issue;
        issue_slot = issue_slot + 1 == kSlots ? 0 : issue_slot + 1;
    };
    auto acquire = [&](int upto) {  // rows [0, upto) of the warp's stream have landed
        while (upto > avail) {
            mbar_wait(bar_s + 8 * wait_slot, wait_parity);
            avail += CR;
            if (++wait_slot == kSlots) {
                wait_slot = 0;
                wait_parity ^= 1u;
            }
        }
    };
    auto release = [&]() {  // every lane is done with the oldest slot(s): refill
        __syncwarp();
        while (pos >= release_at) {
            if (next_issue < n_chunks) advance_issue();
            release_at += CR;
        }
    };
    auto restart_stream = [&]() {  // the cells never change: the first copies of an iteration can start before its parameters exist
        __syncwarp();
        next_issue = 0, avail = 0, release_at = CR, pos = 0;
        roff = issue_slot * kChunk;
        for (int c = 0; c < min(kSlots, n_chunks); ++c) advance_issue();
    };
    restart_stream();

    const int n_iter = MODE == kWriteJ ? 1 : A.num_iter;
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
        typename PixelStats<MODE, PRECISE, REC>::Consts kc;
        {
            float Bs[3], kb[3], kg[3];   // Bs: B in the store's units
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float B = s_state[c], beta = s_state[3 + c], gamma = s_state[6 + c];
                Bs[c] = B * kScale;
                kb[c] = PRECISE ? -beta : -beta * kLog2e;     // e^{-beta z} = 2^{kb z}
                kg[c] = PRECISE ? -gamma : -gamma * kLog2e;
            }
            kc.kb_rg = pk(kb[0], kb[1]), kc.kg_rg = pk(kg[0], kg[1]), kc.nB_rg = pk(-Bs[0], -Bs[1]);
            kc.kb_b = kb[2], kc.kg_b = kg[2], kc.nB_b = -Bs[2];
        }
        // The scalars of this iteration's Adam step wait in shared memory (registers are what the sweep is short of); only
        // the J-parameter mode needs them inside the sweep.
        if (MODE != kWriteJ && threadIdx.x == 96) s_ad = A.adam_tab ? A.adam_tab[it] : AdamScalars{0.f, 1.f, 0.0};
        const AdamScalars ad = (MODE == kParamJ && A.adam_tab) ? A.adam_tab[it] : AdamScalars{0.f, 1.f, 0.0};
        SUCRE_TRACE(0);

        // per-thread partial sums of the ten global sums: fp32 over this warp's tiles (a few dozen per-pixel values,
        // each itself an fp32 sum), promoted to double for everything that follows (warp tree, CTA row, rows of all CTAs, peers)
        float acc[kSums];
#pragma unroll
        for (int i = 0; i < kSums; ++i) acc[i] = 0.f;

        // tile extents relative to the warp's stream and pixel indices are 32-bit (check_store bounds both): the tile
        // loop lives at the register limit
        int t = t_first;
        int rend_next = t < T1 ? (int)(A.row_off[t + 1] - B0) : 0;
        int p_next = t * kTile + lane;
        float Jnext[3] = {0.f, 0.f, 0.f};
        if (A.J && t < T1 && p_next < A.pixels) {
            const float* Jp = A.J + (size_t)3 * p_next;
            Jnext[0] = Jp[0], Jnext[1] = Jp[1], Jnext[2] = Jp[2];
        }

#pragma unroll 1
        for (; t < T1; ++t) {
            const bool head = t < T0;
            const int rend = rend_next;
            const int p = p_next;
            float Jref[3] = {Jnext[0], Jnext[1], Jnext[2]};
            if (t + 1 < T1) {  // the next tile's extent and reference J: their latency hides behind this tile's rows
                rend_next = (int)(A.row_off[t + 2] - B0);
                p_next = p + kTile;
                if (A.J && p_next < A.pixels) {
                    const float* Jp = A.J + (size_t)3 * p_next;
                    Jnext[0] = Jp[0], Jnext[1] = Jp[1], Jnext[2] = Jp[2];
                }
            }
            if (MODE == kWriteJ) {
#pragma unroll
                for (int c = 0; c < 3; ++c) Jref[c] = Jref[c] == Jref[c] ? Jref[c] : 0.f;  // a NaN reference is no reference
            }
            const bool cut = rend > n_rows_w;                     // the tile's last rows are in the next warp's stream
            const int rb = cut ? n_rows_w : rend;                 // stream row where this item ends
            PixelStats<MODE, PRECISE, REC> st;
            st.clear();
            unsigned seen_bits = 0;   // OR of the z bit patterns of the lane's records: non-zero iff it has an observation
            {
                const u64 nJ_rg = pk(-Jref[0] * kScale, -Jref[1] * kScale);
                const float nJ_b = -Jref[2] * kScale;
                const unsigned char* lane_ring = ring + lane * RT::kBytes;
                int r = pos;
                // Two rows per step for instruction-level parallelism (their exp / residual chains are independent until
                // the accumulators); the body is branch-free (sentinels are masked arithmetically).  The ring is dealt
                // with outside the inner loop: it runs up to what has landed, then the consumed slots are refilled and
                // the next chunk is awaited.
                while (true) {
                    const int lim = min(rb, avail);
#pragma unroll kRowLoopUnroll
                    for (; r + 2 <= lim; r += 2) {
                        const int o1 = (roff + kRowBytes) & (kRingBytes - 1);
                        const Raw q0 = *reinterpret_cast<const Raw*>(lane_ring + roff);
                        const Raw q1 = *reinterpret_cast<const Raw*>(lane_ring + o1);
                        roff = (o1 + kRowBytes) & (kRingBytes - 1);
                        seen_bits |= __float_as_uint(RT::range(q0));
                        st.add(RT::unpack(q0), RT::unpack(q1), kc, nJ_rg, nJ_b);
                    }
                    pos = r;
                    if (pos >= release_at) release();
                    if (r + 2 > rb) break;
                    acquire(r + 2);
                }
                if (r < rb) {  // odd row count: one last row, paired with an all-zero record
                    if (r + 1 > avail) acquire(r + 1);
                    const Raw q0 = *reinterpret_cast<const Raw*>(lane_ring + roff);
                    roff = (roff + kRowBytes) & (kRingBytes - 1);
                    seen_bits |= __float_as_uint(RT::range(q0));
#if defined(SUCRE_FIT_ODD_AS_PAIR)
                    st.add(RT::unpack(q0), RT::unpack(Raw{}), kc, nJ_rg, nJ_b);
#else
                    st.add_one(RT::unpack(q0), kc, nJ_rg, nJ_b);
#endif
                    pos = r + 1;
                    if (pos >= release_at) release();
                }
            }
            const bool seen = seen_bits != 0;
            if (head) {  // hand the partial statistics of the previous warp's last tile over
                st.park(park_all + (size_t)warp * kStats * 32, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&park_bar[warp]));
                continue;
            }
            if (cut) {   // the next warp evaluated the rest of this tile first thing (one arrival per iteration: phase = it & 1)
                mbar_wait(smem_u32(&park_bar[warp + 1]), (uint32_t)it & 1u);
                st.add_parked(park_all + (size_t)(warp + 1) * kStats * 32, lane);
            }
            if (MODE == kWriteJ) {
                const long long dst = A.pix ? (long long)A.pix[p] : (p < A.out_pixels ? (long long)p : -1);
                if (dst >= 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        A.J_out[(size_t)3 * dst + c] = seen ? Jref[c] + (st.get(c, 0) / st.get(c, 1)) * kInv : __int_as_float(0x7fc00000);
                }
            } else if (seen) {  // lanes whose pixel has no observation in any kept view contribute nothing and keep their J
                float Jout[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // all in the store's units: S1, S3, S5, S7 carry one factor kScale, S9 two
                    const float S1 = st.get(c, 0), S3 = st.get(c, 2), S5 = st.get(c, 4), S7 = st.get(c, 6), S9 = st.get(c, 8);
                    float delta = 0.f, rh = S3, rza = S5, rzg = S7, rr = S9;
                    if (MODE == kClosedForm) {
                        delta = S1 / st.get(c, 1);
                        rh = fmaf(-delta, st.get(c, 3), S3);
                        rza = fmaf(-delta, st.get(c, 5), S5);
                        rzg = fmaf(-delta, st.get(c, 7), S7);
                        rr = fmaf(-delta, S1, S9);
                    }
                    const float Js = fmaf(Jref[c], kScale, delta);   // J in the store's units
                    Jout[c] = Js * kInv;
                    acc[c] += rh;                                // sum r (1 - e^{-gamma z})    x kScale
                    acc[3 + c] = fmaf(Js, rza, acc[3 + c]);      // sum r J z e^{-beta z}       x kScale^2
                    acc[6 + c] += rzg;                           // sum r z e^{-gamma z}        x kScale; times B at the end of the sweep
                    acc[9] += rr;                                // sum r^2                     x kScale^2
                    if (MODE == kParamJ) {
                        // dL/dJ = -(2 / 3N) sum r a; the pixel's own Adam step with the pre-step B, beta, gamma (sucre.py:144-148)
                        float* mv = A.J_moments + (size_t)6 * p;
                        float m = mv[c], v = mv[3 + c];
                        Jout[c] = adam_update(Jref[c], (float)(-ad.grad_scale) * (S1 * kInv), m, v, ad.neg_step_size, ad.bc2_sqrt);
                        mv[c] = m;
                        mv[3 + c] = v;
                    }
                }
                A.J[(size_t)3 * p + 0] = Jout[0];
                A.J[(size_t)3 * p + 1] = Jout[1];
                A.J[(size_t)3 * p + 2] = Jout[2];
            }
        }
        if (MODE == kWriteJ) return;
        SUCRE_TRACE_WARP();

        // warp tree -> one slot per warp -> one row per CTA (the store's units are divided out here, in double)
#pragma unroll
        for (int i = 0; i < kSums; ++i) {
            double v = (double)acc[i];
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (lane == 0) sm[warp][i] = v;
        }
        __syncthreads();
        // A CTA's row is published as 20 tagged 8-byte words {tag : 32 | half a double : 32} (two tag parities: nobody is
        // two iterations ahead).  An aligned 8-byte store is atomic, so a word that carries this iteration's tag IS the
        // data: no fence, no counter, no second round trip to L2 (the layout of the exchange between GPUs below).
        // Every CTA polls all rows, reduces them ITSELF in the same fixed order and takes the same Adam step on its own
        // copy of the parameters and moments: no CTA waits for another one to publish the result, and the next sweep
        // starts straight from shared memory.  CTA 0 alone talks to the outside (history, sums, peers, final state).
        SUCRE_TRACE(1);
        const unsigned tag = s_tag_base + (unsigned)it + 1u;
        unsigned long long* const rows = A.rows + (size_t)(tag & 1u) * kMaxFitCtas * kLLWords;
        if (threadIdx.x < kSums) {
            double part[4] = {0.0, 0.0, 0.0, 0.0};   // four interleaved chains, not one 16-deep chain of dependent double additions
#pragma unroll
            for (int wi = 0; wi < kFitWarps; ++wi) part[wi & 3] += sm[wi][threadIdx.x];
            double unit = threadIdx.x < 3 ? 1.0 / (double)kScale : 1.0 / ((double)kScale * (double)kScale);
            if (threadIdx.x >= 6 && threadIdx.x < 9) unit = (double)s_state[threadIdx.x - 6] / (double)kScale;   // sum r B z e^{-gamma z}
            const unsigned long long bits = (unsigned long long)__double_as_longlong(((part[0] + part[1]) + (part[2] + part[3])) * unit);
            unsigned long long* w = rows + (size_t)blockIdx.x * kLLWords + 2 * threadIdx.x;
            st_relaxed_gpu(w, ((unsigned long long)tag << 32) | (bits & 0xffffffffull));
            st_relaxed_gpu(w + 1, ((unsigned long long)tag << 32) | (bits >> 32));
        }
        SUCRE_TRACE(2);
        // The next iteration's first rows are requested now: in flight while the sums are reduced, but behind this CTA's row
        // (148 SMs x 16 warps x 8 KB of copies ahead of it would hold the row, and with it every other CTA, back).
        if (it + 1 < n_iter) restart_stream();
        for (int col = warp; col < kSums; col += kFitWarps) {
            const double v = gather_column(rows + 2 * col, tag, lane, (int)gridDim.x);
            if (lane == 0) tot[col] = v;
        }
        SUCRE_TRACE(3);
        __syncthreads();
        SUCRE_TRACE(4);
        if (A.world > 1) {
            // One-shot all-reduce over NVLink.  CTA 0 of every rank stores the rank's 10 sums, as 20 tagged 8-byte words,
            // into every rank's buffer (slot [epoch & 1][rank], its own included); every CTA polls the words of all ranks
            // in its own GPU's buffer until they carry this epoch and adds the rows in rank order — the same order
            // everywhere, so all CTAs of all ranks take the identical Adam step without any host or NCCL round trip.
            // Two parities: a rank can be at most one iteration ahead of the slowest reader of its previous message.
            // The wait is bounded: after SUCRE_PEER_TIMEOUT_NS a silent peer raises status bit 0 and the loop carries on.
            const unsigned epoch = A.epoch + (unsigned)it;
            const unsigned par = epoch & 1u;
            if (threadIdx.x < A.world * kLLWords) {
                const int p = threadIdx.x / kLLWords, j = threadIdx.x % kLLWords;
                if (blockIdx.x == 0) {
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(tot[j >> 1]);
                    const unsigned half = (j & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
                    PeerSlot* dst = reinterpret_cast<PeerSlot*>(A.peer[p]) + par * SUCRE_MAX_PEERS + A.rank;
                    st_relaxed_sys(&dst->w[j], ((unsigned long long)epoch << 32) | half);
                }
                const PeerSlot* mine = reinterpret_cast<const PeerSlot*>(A.peer[A.rank]) + par * SUCRE_MAX_PEERS + p;
                const unsigned long long t0 = globaltimer();
                unsigned long long word;
                while ((unsigned)((word = ld_relaxed_sys(&mine->w[j])) >> 32) != epoch) {
                    if (globaltimer() - t0 > SUCRE_PEER_TIMEOUT_NS) {
                        atomicOr(A.ticket + 1, 1u);
                        break;
                    }
                }
                ll_half[p][j] = (unsigned)word;
            }
            __syncthreads();
            if (threadIdx.x < kSums) {
                double v = 0.0;
                for (int p = 0; p < A.world; ++p)
                    v += __longlong_as_double((long long)(((unsigned long long)ll_half[p][2 * threadIdx.x + 1] << 32) | ll_half[p][2 * threadIdx.x]));
                tot[threadIdx.x] = v;
            }
            __syncthreads();
        }
        const bool writer = blockIdx.x == 0;
        float* history_row = (writer && A.history) ? A.history + (size_t)it * kSums : nullptr;
        if (writer && threadIdx.x < kSums && A.sums_out) A.sums_out[threadIdx.x] = tot[threadIdx.x];
        if (A.do_step && threadIdx.x < 9) {
            const int i = threadIdx.x;
            // B: -2 r (1-g); beta: +2 r J z a; gamma: -2 r B z g   (all times 1 / 3N)
            const AdamScalars as = s_ad;
            const float g = (float)((i >= 3 && i < 6 ? as.grad_scale : -as.grad_scale) * tot[i]);
            float m = s_state[9 + i], v = s_state[18 + i];
            const float pnew = adam_update(s_state[i], g, m, v, as.neg_step_size, as.bc2_sqrt);
            s_state[i] = pnew, s_state[9 + i] = m, s_state[18 + i] = v;
            if (history_row) history_row[i] = pnew;
            if (writer && it + 1 == n_iter) {   // the state leaves the kernel once
                A.params[i] = pnew;
                A.moments[i] = m;
                A.moments[9 + i] = v;
            }
        }
        if (threadIdx.x == 9 && history_row) history_row[9] = (float)tot[9];
        __syncthreads();   // s_state is read at the top of the next iteration
        SUCRE_TRACE(5);
    }
    // every CTA read the base before the first rendezvous of this launch, so it can move on now
    if (MODE != kWriteJ && blockIdx.x == 0 && threadIdx.x == 0) A.ticket[2] = s_tag_base + (unsigned)n_iter;
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

// Static partition of the store over the resident warps.  Cost of tile t = kTileCost + rows(t), laid out on a
// virtual axis (tile t starts at row_off[t] + kTileCost * t).  Two levels: the CTAs split the axis at the tile
// boundaries nearest to its n_ctas-quantiles; inside a CTA, warp k starts at the k/kFitWarps-quantile of the CTA's
// range.  A warp boundary inside the rows of a tile SPLITS the tile: the warp before owns it (and its head rows), the
// warp after starts with its tail rows.  A tile is split at most once (a second boundary inside the same rows moves
// up to the end of the tile) and never between two CTAs, so the two halves meet in shared memory.
//   part_tile[w] = first tile warp w owns, part_row[w] = first row of its stream; [n_warps] = (n_tiles, rows).
__global__ void partition_kernel(const long long* __restrict__ row_off, int n_tiles, int n_ctas, long long* __restrict__ part_row,
                                 int* __restrict__ part_tile) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_warps = n_ctas * kFitWarps;
    if (w > n_warps) return;
    const long long R = row_off[n_tiles];
    const long long V = R + (long long)kTileCost * n_tiles;
    auto vstart = [&](int t) { return row_off[t] + (long long)kTileCost * t; };
    auto tile_of = [&](long long x) {  // largest t in [0, n_tiles) with vstart(t) <= x
        int lo = 0, hi = n_tiles - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (vstart(mid) <= x) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    auto cta_boundary = [&](int c) {  // first tile of CTA c: the tile boundary nearest to the quantile
        if (c >= n_ctas) return n_tiles;
        const long long x = V * c / n_ctas;
        const int t = tile_of(x);
        const long long off = x - vstart(t), cost = kTileCost + row_off[t + 1] - row_off[t];
        return 2 * off < cost ? t : t + 1;
    };
    const int c = w / kFitWarps, k = w % kFitWarps;
    const int tb = cta_boundary(c);
    if (k == 0 || tb == n_tiles) {
        part_tile[w] = tb;
        part_row[w] = row_off[tb];
        return;
    }
    const long long Cc = vstart(tb), Cn = vstart(cta_boundary(c + 1));
    const long long ideal = Cc + (Cn - Cc) * k / kFitWarps;
    if (ideal >= V) {
        part_tile[w] = n_tiles;
        part_row[w] = R;
        return;
    }
    const int t = tile_of(ideal);
    const long long off = ideal - vstart(t) - kTileCost;  // rows of tile t before the boundary (<= 0: in its overhead part)
    if (off <= 0) {
        part_tile[w] = t;
        part_row[w] = row_off[t];
        return;
    }
    const long long prev = Cc + (Cn - Cc) * (k - 1) / kFitWarps;
    const bool up = prev - vstart(t) - kTileCost > 0;     // the previous boundary already cut this tile
    part_tile[w] = t + 1;
    part_row[w] = up ? row_off[t + 1] : row_off[t] + off;
}

static bool precise_exp() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SUCRE_PRECISE_EXP");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <int MODE, int REC, bool PRECISE>
static int occupancy() {
    cudaFuncSetAttribute(fit_kernel<MODE, REC, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFitSmem);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fit_kernel<MODE, REC, PRECISE>, kFitThreads, kFitSmem) != cudaSuccess)
        per_sm = 0;
    return per_sm;
}

template <int MODE>
static int occupancy_mode() {
    return min(min(occupancy<MODE, SUCRE_REC_Z_U8, false>(), occupancy<MODE, SUCRE_REC_Z_U8, true>()),
               min(occupancy<MODE, SUCRE_REC_Z_F32, false>(), occupancy<MODE, SUCRE_REC_Z_F32, true>()));
}

// persistent grid shared by every mode (the warp partition is computed for it): one CTA per SM.  Function attributes
// and the SM count are per device, so both are set up once per device ordinal.
static int fit_grid() {
    constexpr int kMaxDevices = 64;
    static int ctas[kMaxDevices] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    if (ctas[dev] == 0) {
        int per_sm = min(occupancy_mode<kClosedForm>(), min(occupancy_mode<kParamJ>(), occupancy_mode<kWriteJ>()));
        if (per_sm <= 0) per_sm = 1;
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        ctas[dev] = min(kMaxFitCtas, sms * min(per_sm, 1));
    }
    return ctas[dev];
}

template <int MODE>
static void launch_fit(const FitArgs& a, int rec, int ctas, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kFitThreads);
    cfg.dynamicSmemBytes = kFitSmem;
    cfg.stream = st;
    if (rec == SUCRE_REC_Z_U8) {
        if (precise_exp()) cudaLaunchKernelEx(&cfg, fit_kernel<MODE, SUCRE_REC_Z_U8, true>, a);
        else cudaLaunchKernelEx(&cfg, fit_kernel<MODE, SUCRE_REC_Z_U8, false>, a);
    } else {
        if (precise_exp()) cudaLaunchKernelEx(&cfg, fit_kernel<MODE, SUCRE_REC_Z_F32, true>, a);
        else cudaLaunchKernelEx(&cfg, fit_kernel<MODE, SUCRE_REC_Z_F32, false>, a);
    }
}

static int check_store(const sucre_store* s, const char* who) {
    SUCRE_REQUIRE(s != nullptr, "%s: null store", who);
    SUCRE_REQUIRE(s->row_off, "%s: null row_off in store", who);
    SUCRE_REQUIRE(s->n_rows >= 0 && (s->cells || s->n_rows == 0), "%s: null cells in a store of %lld rows", who, (long long)s->n_rows);
    SUCRE_REQUIRE(s->n_tiles > 0 && s->pixels > 0 && (s->pix || s->pixels <= (int64_t)s->n_tiles * kTile), "%s: bad store sizes", who);
    SUCRE_REQUIRE((reinterpret_cast<uintptr_t>(s->cells) & 15) == 0, "%s: cells must be 16-byte aligned", who);
    SUCRE_REQUIRE(s->record_format == SUCRE_REC_Z_U8 || s->record_format == SUCRE_REC_Z_F32,
                  "%s: this entry point reads {z, I} stores (SUCRE_REC_Z_U8 / SUCRE_REC_Z_F32), got format %d", who, s->record_format);
    return 0;
}

static FitArgs base_args(const sucre_store* s, void* workspace) {
    FitArgs a{};
    a.cells = reinterpret_cast<const unsigned char*>(s->cells);
    a.row_off = (const long long*)s->row_off;
    a.n_tiles = s->n_tiles;
    a.pixels = s->pix ? (long long)s->n_tiles * kTile : s->pixels;   // slot-ordered arrays: every slot of a permuted store is addressable
    a.pix = s->pix;
    a.out_pixels = s->pixels;
    char* ws = (char*)workspace;
    a.rows = (unsigned long long*)(ws + kWsPartials);
    a.part_row = (const long long*)(ws + kWsPartRow);
    a.part_tile = (const int*)(ws + kWsPartTile);
    a.ticket = (unsigned*)(ws + kWsTicket);
    a.adam_tab = (const AdamScalars*)(ws + kWsAdamTab);
    a.num_iter = 1;
    a.rank = 0;
    a.world = 1;
    return a;
}

}  // namespace sucre

using namespace sucre;

extern "C" size_t sucre_fit_workspace_bytes(void) { return kWsBytes; }

#if defined(SUCRE_FIT_TRACE)
extern "C" int sucre_debug_fit_trace(unsigned long long* cta_stamps, unsigned long long* warp_stamps) {
    if (cudaMemcpyFromSymbol(cta_stamps, g_trace, sizeof(g_trace)) != cudaSuccess) return 1;
    if (cudaMemcpyFromSymbol(warp_stamps, g_trace_warp, sizeof(g_trace_warp)) != cudaSuccess) return 1;
    return 0;
}
#endif

extern "C" int sucre_fit_prepare(const sucre_store* store_host, void* workspace, void* stream) {
    clear_error();
    if (check_store(store_host, "sucre_fit_prepare")) return 1;
    SUCRE_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "sucre_fit_prepare: workspace must be non-null, 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_ctas = fit_grid();
    const int n_warps = n_ctas * kFitWarps;
    char* ws = (char*)workspace;
    // rows and tag base start cleared: tags count from 1, so a cleared word is never taken for data
    SUCRE_CUDA(cudaMemsetAsync(ws + kWsPartials, 0, kWsPartRow - kWsPartials, st));
    SUCRE_CUDA(cudaMemsetAsync(ws + kWsTicket, 0, 16, st));
    partition_kernel<<<(n_warps + 1 + 255) / 256, 256, 0, st>>>((const long long*)store_host->row_off, store_host->n_tiles, n_ctas,
                                                                 (long long*)(ws + kWsPartRow), (int*)(ws + kWsPartTile));
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
    const AdamScalars one = mode == kParamJ ? adam_scalars(t, lr, n_obs) : AdamScalars{0.f, 1.f, 0.0};
    SUCRE_CUDA(cudaMemcpyAsync((char*)workspace + kWsAdamTab, &one, sizeof one, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (mode == kClosedForm) launch_fit<kClosedForm>(a, store_host->record_format, fit_grid(), (cudaStream_t)stream);
    else launch_fit<kParamJ>(a, store_host->record_format, fit_grid(), (cudaStream_t)stream);
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
    cudaStream_t st = (cudaStream_t)stream;
    // the whole loop is ONE launch per kMaxLoopIters iterations: the per-iteration Adam scalars (python-float arithmetic of
    // torch.optim.Adam, evaluated here in double) go to the workspace, the iteration flag is cleared, the grid stays resident
    for (int done = 0; done < num_iter; done += kMaxLoopIters) {
        const int n = min(kMaxLoopIters, num_iter - done);
        AdamScalars tab[kMaxLoopIters];
        for (int it = 0; it < n; ++it) tab[it] = adam_scalars(first_step + done + it, lr, n_obs);
        SUCRE_CUDA(cudaMemcpyAsync((char*)workspace + kWsAdamTab, tab, sizeof(AdamScalars) * n, cudaMemcpyHostToDevice, st));  // pageable source: staged before the call returns
        a.num_iter = n;
        a.history = history ? history + (size_t)done * kSums : nullptr;
        a.epoch = first_epoch + (uint32_t)done;
        if (mode == kClosedForm) launch_fit<kClosedForm>(a, store_host->record_format, ctas, st);
        else launch_fit<kParamJ>(a, store_host->record_format, ctas, st);
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

__global__ void status_kernel(unsigned* ticket, uint32_t* out) {
    *out = ticket[1];
    ticket[1] = 0u;
}

extern "C" int sucre_fit_status(void* workspace, uint32_t* status, void* stream) {
    clear_error();
    SUCRE_REQUIRE(workspace && status, "sucre_fit_status: null pointer");
    status_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned*)((char*)workspace + kWsTicket), status);
    return check_launch("status_kernel");
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
    launch_fit<kWriteJ>(a, store_host->record_format, fit_grid(), (cudaStream_t)stream);
    return check_launch("fit_kernel<write J>");
}
