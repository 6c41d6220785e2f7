// Stage 2 with the artificial-light model (--light-model, sucre.py:44-46, 54-61) for sm_100a.
//
//   lP = R cP + t,  lp = lP.xy / lP.z,  l = exp(-lp^T Sigma^-1 lp / 2),  z = ||cP|| + ||lP||
//   I_hat = l (J e^{-beta z} + B (1 - e^{-gamma z}))
//
// R, t = se3.exp(cam2light) and Sigma^-1 = (sigma^T sigma)^-1 are tiny host-side torch evaluations (the reference's
// own expressions); the kernels take them as 15 of the 24 "derived" parameters and return, besides the nine sums
// of the plain model, the sums that carry dL/dSigma^-1, dL/dR and dL/dt — the host pushes those through matrix_exp
// and the 2x2 inverse.  Two kernels per iteration instead of one sweep: the light gradients need the final
// residual of every observation times per-observation geometry, which does not factor through per-pixel
// statistics the way the plain model does.  One warp per tile, lanes own pixels and walk their ELL column with direct,
// row-coalesced 128-bit global loads: this optional mode is correctness-first.
#include "common.cuh"

namespace sucre {

constexpr int kLightThreads = 256;
constexpr int kLightWarps = kLightThreads / 32;
constexpr int kLightSums = 25;
constexpr int kLightMaxCtas = 800;  // kLightSums * kLightMaxCtas doubles fit the fit workspace's partial-sum area

struct LightParams {
    float B[3], beta[3], gamma[3], R[9], t[3], S[3];
};

__device__ __forceinline__ LightParams load_light(const float* __restrict__ p) {
    LightParams q;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        q.B[i] = p[i];
        q.beta[i] = p[3 + i];
        q.gamma[i] = p[6 + i];
        q.t[i] = p[18 + i];
        q.S[i] = p[21 + i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) q.R[i] = p[9 + i];
    return q;
}

// per-observation light geometry
struct LightGeom {
    float lP[3], nl, x, y, l, z;
};

__device__ __forceinline__ LightGeom light_geom(const LightParams& q, const float4 c) {
    LightGeom g;
#pragma unroll
    for (int i = 0; i < 3; ++i) g.lP[i] = q.R[3 * i] * c.x + q.R[3 * i + 1] * c.y + q.R[3 * i + 2] * c.z + q.t[i];
    g.nl = sqrtf(g.lP[0] * g.lP[0] + g.lP[1] * g.lP[1] + g.lP[2] * g.lP[2]);
    g.x = g.lP[0] / g.lP[2];
    g.y = g.lP[1] / g.lP[2];
    const float quad = q.S[0] * g.x * g.x + 2.0f * q.S[1] * g.x * g.y + q.S[2] * g.y * g.y;
    g.l = expf(-0.5f * quad);
    g.z = c.w + g.nl;  // ||cP|| (stored) + ||lP||
    return g;
}

// Walks the observations of this lane's pixel in one tile of a SUCRE_REC_P_* store (ELL rows: the lane's column, top
// to bottom, until the first sentinel): f({cP_x, cP_y, cP_z, ||cP||}, {I_r, I_g, I_b, 0}).  Returns their number.
// The loads run two rows ahead of the arithmetic (a column's records are 512 / 1024 bytes apart, a warp's row is one
// coalesced line set), so the few warps a register-heavy caller leaves per SM still keep HBM busy.
template <class F>
__device__ __forceinline__ int walk_tile(const sucre_store& S, int tile, int lane, F&& f) {
    const long long r0 = S.row_off[tile];
    const int n = (int)(S.row_off[tile + 1] - r0);
    const float4* cells = reinterpret_cast<const float4*>(S.cells);
    const float4 none = make_float4(0.f, 0.f, 0.f, 0.f);
    int seen = 0;
    if (S.record_format == SUCRE_REC_P_U8) {
        const float4* col = cells + r0 * kTile + lane;
        float4 q0 = n > 0 ? __ldg(col) : none, q1 = n > 1 ? __ldg(col + kTile) : none;
        for (int j = 0; j < n; ++j) {
            const float4 q = q0;
            q0 = q1;
            q1 = j + 2 < n ? __ldg(col + (size_t)(j + 2) * kTile) : none;
            if (q.z == 0.0f) break;  // sentinel: a real observation has cP_z = source depth > 0
            // sucre.py:53 cP.norm(dim=0): sequential squares, no fma; loader.py:157: u8 / 255
            const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)));
            const uint32_t rgb = __float_as_uint(q.w);
            f(make_float4(q.x, q.y, q.z, nrm),
              make_float4(__fdiv_rn((float)(rgb & 0xffu), 255.0f), __fdiv_rn((float)((rgb >> 8) & 0xffu), 255.0f),
                          __fdiv_rn((float)((rgb >> 16) & 0xffu), 255.0f), 0.f));
            ++seen;
        }
    } else {
        const float4* col = cells + 2 * (r0 * kTile + lane);
        float4 c0 = n > 0 ? __ldg(col) : none, i0 = n > 0 ? __ldg(col + 1) : none;
        for (int j = 0; j < n; ++j) {
            const float4 c = c0, I4 = i0;
            if (j + 1 < n) c0 = __ldg(col + (size_t)(j + 1) * 2 * kTile), i0 = __ldg(col + (size_t)(j + 1) * 2 * kTile + 1);
            if (c.z == 0.0f) break;
            f(c, I4);
            ++seen;
        }
    }
    return seen;
}

// the entry of the pixel-ordered J arrays that lane `lane` of tile `tile` owns, -1 for none
__device__ __forceinline__ long long pixel_of(const sucre_store& S, int tile, int lane) {
    const long long q = (long long)tile * kTile + lane;
    if (S.pix) return (long long)__ldg(S.pix + q);
    return q < S.pixels ? q : -1;
}

// closed-form J with the light terms (sucre.py:66-77): absorption = l e^{-beta z}, backscatter = l B (1 - e^{-gamma z})
__global__ void __launch_bounds__(kLightThreads)
light_J_kernel(const __grid_constant__ sucre_store S, const float* __restrict__ params, float* __restrict__ J) {
    const LightParams q = load_light(params);
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kLightWarps + (threadIdx.x >> 5);
    if (tile >= S.n_tiles) return;
    float num[3] = {0.f, 0.f, 0.f}, den[3] = {0.f, 0.f, 0.f};
    const int seen = walk_tile(S, tile, lane, [&](const float4 c, const float4 I4) {
        const LightGeom g = light_geom(q, c);
        const float I[3] = {I4.x, I4.y, I4.z};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float absorption = g.l * expf(-q.beta[ch] * g.z);
            const float backscatter = g.l * q.B[ch] * (1.0f - expf(-q.gamma[ch] * g.z));
            num[ch] += (I[ch] - backscatter) * absorption;
            den[ch] += absorption * absorption;
        }
    });
    const long long p = pixel_of(S, tile, lane);
    if (p >= 0) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) J[3 * p + ch] = seen ? num[ch] / den[ch] : __int_as_float(0x7fc00000);
    }
}

__device__ __forceinline__ float adam_update1(float p, float g, float& m, float& v, float neg_step_size, float bc2_sqrt) {
    m = m + (float)(1.0 - 0.9) * (g - m);
    v = v * (float)0.999 + (float)(1.0 - 0.999) * (g * g);
    const float denom = sqrtf(v) / bc2_sqrt + (float)1e-8;
    return p + neg_step_size * (m / denom);
}

// residual pass: J given per pixel -> 25 sums (see include/sucre_b200.h); PARAM_J: Adam step of J fused
template <bool PARAM_J>
__global__ void __launch_bounds__(kLightThreads, 2)
light_sums_kernel(const __grid_constant__ sucre_store S, const float* __restrict__ params, float* __restrict__ J,
                  float* __restrict__ J_moments, float grad_scale, float neg_step_size, float bc2_sqrt,
                  double* __restrict__ partials) {
    const LightParams q = load_light(params);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // A thread's running sums stay fp32 over the handful of tiles it owns (a few hundred terms, the same rounding
    // scale as one tile's) and become double for everything summed across threads.
    float s[kLightSums];
#pragma unroll
    for (int i = 0; i < kLightSums; ++i) s[i] = 0.f;

    for (int tile = blockIdx.x * kLightWarps + warp; tile < S.n_tiles; tile += gridDim.x * kLightWarps) {
        const long long p = pixel_of(S, tile, lane);
        float Jp[3] = {0.f, 0.f, 0.f};
        if (p >= 0) {
            Jp[0] = J[3 * p], Jp[1] = J[3 * p + 1], Jp[2] = J[3 * p + 2];
        }
        float gJ[3] = {0.f, 0.f, 0.f};
        const int seen = walk_tile(S, tile, lane, [&](const float4 c, const float4 I4) {
            const LightGeom g = light_geom(q, c);
            const float I[3] = {I4.x, I4.y, I4.z};
            float g_l = 0.f, g_z = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float a = expf(-q.beta[ch] * g.z), e = expf(-q.gamma[ch] * g.z);
                const float M = Jp[ch] * a + q.B[ch] * (1.0f - e);
                const float r = I[ch] - g.l * M;  // sucre.py:81
                s[ch] += r * g.l * (1.0f - e);
                s[3 + ch] += r * g.l * Jp[ch] * g.z * a;
                s[6 + ch] += r * g.l * q.B[ch] * g.z * e;
                s[9] += r * r;
                g_l += r * M;
                g_z += r * g.l * (-q.beta[ch] * Jp[ch] * a + q.gamma[ch] * q.B[ch] * e);
                if (PARAM_J) gJ[ch] += r * g.l * a;
            }
            // dL/dSigma^-1 carrier: g_l * l * lp lp^T
            const float gll = g_l * g.l;
            s[10] += gll * g.x * g.x;
            s[11] += gll * g.x * g.y;
            s[12] += gll * g.y * g.y;
            // g_lP = g_l dl/dlP + g_z lP/||lP||;  dl/dlp = -l Sigma^-1 lp, lp = lP.xy / lP.z
            const float dlx = -g.l * (q.S[0] * g.x + q.S[1] * g.y), dly = -g.l * (q.S[1] * g.x + q.S[2] * g.y);
            const float iz = 1.0f / g.lP[2], zn = g_z / g.nl;
            const float gP[3] = {g_l * dlx * iz + zn * g.lP[0], g_l * dly * iz + zn * g.lP[1],
                                 -g_l * (dlx * g.x + dly * g.y) * iz + zn * g.lP[2]};
            const float cP[3] = {c.x, c.y, c.z};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) s[13 + 3 * i + j] += gP[i] * cP[j];
                s[22 + i] += gP[i];
            }
        });
        if (seen) {
            if (PARAM_J) {
                float* mv = J_moments + 6 * p;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float m = mv[ch], v = mv[3 + ch];
                    J[3 * p + ch] = adam_update1(Jp[ch], -grad_scale * gJ[ch], m, v, neg_step_size, bc2_sqrt);
                    mv[ch] = m;
                    mv[3 + ch] = v;
                }
            }
        }
    }
    __shared__ double sm[kLightWarps][kLightSums];
    for (int i = 0; i < kLightSums; ++i) {
        double v = (double)s[i];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) sm[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLightSums) {
        double v = 0.0;
        for (int wi = 0; wi < kLightWarps; ++wi) v += sm[wi][threadIdx.x];
        partials[(size_t)blockIdx.x * kLightSums + threadIdx.x] = v;
    }
}

// fixed-order reduction of the per-CTA rows: warp w sums column w
__global__ void __launch_bounds__(kLightSums * 32)
light_reduce_kernel(const double* __restrict__ partials, int n_rows, double* __restrict__ sums) {
    const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
    double v = 0.0;
    for (int r = lane; r < n_rows; r += 32) v += partials[(size_t)r * kLightSums + col];
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) sums[col] = v;
}

static int check_light_store(const sucre_store* s, const char* who) {
    SUCRE_REQUIRE(s != nullptr, "%s: null store", who);
    SUCRE_REQUIRE(s->cells && s->row_off, "%s: null pointer in store", who);
    SUCRE_REQUIRE(s->n_tiles > 0 && s->pixels > 0 && (s->pix || s->pixels <= (int64_t)s->n_tiles * kTile), "%s: bad store sizes", who);
    SUCRE_REQUIRE(s->record_format == SUCRE_REC_P_U8 || s->record_format == SUCRE_REC_P_F32,
                  "%s: the light model needs stores with the camera-frame point (SUCRE_REC_P_U8 / SUCRE_REC_P_F32), got format %d",
                  who, s->record_format);
    return 0;
}

}  // namespace sucre

using namespace sucre;

extern "C" int sucre_light_J(const sucre_store* store_host, const float* params24, float* J, void* stream) {
    clear_error();
    if (check_light_store(store_host, "sucre_light_J")) return 1;
    SUCRE_REQUIRE(params24 && J, "sucre_light_J: null pointer");
    light_J_kernel<<<(store_host->n_tiles + kLightWarps - 1) / kLightWarps, kLightThreads, 0, (cudaStream_t)stream>>>(*store_host, params24, J);
    return check_launch("light_J_kernel");
}

extern "C" int sucre_light_sums(int mode, const sucre_store* store_host, const float* params24, float* J, float* J_moments,
                                int64_t n_obs, int t, double lr, double* sums, void* workspace, void* stream) {
    clear_error();
    if (check_light_store(store_host, "sucre_light_sums")) return 1;
    SUCRE_REQUIRE(params24 && J && sums && workspace, "sucre_light_sums: null pointer");
    SUCRE_REQUIRE(mode == SUCRE_FIT_CLOSED_FORM || (mode == SUCRE_FIT_PARAM_J && J_moments && n_obs > 0 && t >= 1),
                  "sucre_light_sums: bad mode/arguments");
    cudaStream_t st = (cudaStream_t)stream;
    // two CTAs are resident per SM (launch bounds); two waves of them, each warp striding over the tiles
    const int ctas = min(min(kLightMaxCtas, 4 * num_sms()), (store_host->n_tiles + kLightWarps - 1) / kLightWarps);
    double* partials = (double*)workspace;
    if (mode == SUCRE_FIT_PARAM_J) {
        const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
        light_sums_kernel<true><<<ctas, kLightThreads, 0, st>>>(*store_host, params24, J, J_moments, (float)(2.0 / (3.0 * (double)n_obs)),
                                                               (float)(-(lr / bc1)), (float)sqrt(bc2), partials);
    } else {
        light_sums_kernel<false><<<ctas, kLightThreads, 0, st>>>(*store_host, params24, J, nullptr, 0.f, 0.f, 1.f, partials);
    }
    light_reduce_kernel<<<1, kLightSums * 32, 0, st>>>(partials, ctas, sums);
    return check_launch("light_sums_kernel");
}
