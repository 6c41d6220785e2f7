// Stage 1 — fused multi-view correspondence gather for sm_100a.
//
// What the reference does with ~80 whole-image ATen kernels, a reverse index map and four stream
// compactions per (target, view) pair (sfm.py:115-138, 154-175; 306 B of intermediates per pixel-view)
// is done here per target pixel, in registers:
//   match   one thread owns PIX target pixels, keeps their world points, and walks a chunk of 32 source
//           views whose constants sit in shared memory; the backward projection is evaluated only at the
//           source pixel the forward projection lands on (one 2-byte gather), and the per-(tile, view)
//           result is a 32-bit ballot mask.
//   plan    per-view counts -> min_cover decision -> per-tile record/block/segment counts -> exclusive scan.
//   sample  one warp per tile re-projects only the matched pixels, fetches depth + colour of the source
//           pixel and writes {z, I} records into the tile-major segmented stream (include/sucre_b200.h),
//           lane-major within each segment of up to 15 views.
//
// Arithmetic contract (SURVEY.md §8a', pinned by tests against the reference's own outputs): every fp32
// operation below is a single correctly rounded IEEE op written with an explicit intrinsic, in the order the
// reference's ATen calls perform them; this file is additionally compiled with -fmad=false.
#include "common.cuh"

namespace sucre {

// y = M (3x3 row-major) applied to x the way torch.mm(3x3, 3xn) rounds it on the reference's CPU path:
// per row fma(m2, x2, fma(m1, x1, m0*x0)).
__device__ __forceinline__ void mat3(const float* M, float x0, float x1, float x2, float& y0, float& y1, float& y2) {
    y0 = __fmaf_rn(M[2], x2, __fmaf_rn(M[1], x1, __fmul_rn(M[0], x0)));
    y1 = __fmaf_rn(M[5], x2, __fmaf_rn(M[4], x1, __fmul_rn(M[3], x0)));
    y2 = __fmaf_rn(M[8], x2, __fmaf_rn(M[7], x1, __fmul_rn(M[6], x0)));
}

// sfm.py:90-93 unproject_depth with depth = u16 / 1000 (loader.py:167): cP = Kinv @ (d * (u+.5, v+.5, 1))
__device__ __forceinline__ void unproject(const float* Kinv, int u, int v, float d, float& c0, float& c1, float& c2) {
    const float x0 = __fmul_rn(d, __fadd_rn((float)u, 0.5f));
    const float x1 = __fmul_rn(d, __fadd_rn((float)v, 0.5f));
    mat3(Kinv, x0, x1, d, c0, c1, c2);
}

// sfm.py:49-55 Pose.transform: (R @ P) + t, the add rounded separately
__device__ __forceinline__ void rigid(const float* R, const float* t, float x0, float x1, float x2, float& y0, float& y1, float& y2) {
    mat3(R, x0, x1, x2, y0, y1, y2);
    y0 = __fadd_rn(y0, t[0]);
    y1 = __fadd_rn(y1, t[1]);
    y2 = __fadd_rn(y2, t[2]);
}

// sfm.py:103-107 project_to_view followed by sfm.py:116-117: .long() truncates toward zero, then
// 0 <= u < W, 0 <= v < H.  trunc(x) >= 0 <=> x > -1, so (-1,0) is accepted as index 0 exactly like the
// reference; NaN / inf / huge fail the comparisons (the reference's INT64_MIN fails `0 <=`).
__device__ __forceinline__ bool project(const float* Ri, const float* ti, const float* K, int W, int H,
                                        float w0, float w1, float w2, int& u, int& v) {
    float c0, c1, c2, p0, p1, p2;
    rigid(Ri, ti, w0, w1, w2, c0, c1, c2);
    mat3(K, c0, c1, c2, p0, p1, p2);
    const float px = __fdiv_rn(p0, p2), py = __fdiv_rn(p1, p2);
    const bool in = px > -1.0f && px < (float)W && py > -1.0f && py < (float)H;
    u = __float2int_rz(px);
    v = __float2int_rz(py);
    return in;
}

constexpr int kViewWords = sizeof(sucre_view) / 4;  // 52
constexpr int kSegViewsMax = 15;                    // per-lane record counts of a segment are read as small ints
constexpr int kSegHeaderCells = SUCRE_SEGMENT_HEADER_CELLS;
constexpr int kChunk = 32;                          // source views per CTA: lane j keeps the mask of view j
constexpr int kWarps = 8;
#ifndef SUCRE_MATCH_PIX
#define SUCRE_MATCH_PIX 2  // target pixels per thread of gather_match_kernel
#endif

template <int PIX>
__global__ void __launch_bounds__(kWarps * 32)
gather_match_kernel(const __grid_constant__ sucre_view T, const sucre_view* __restrict__ views, int n_views,
                    uint32_t* __restrict__ masks, int first_tile, int n_tiles) {
    __shared__ __align__(16) sucre_view sv[kChunk];  // 6.5 KB
    const int vbase = blockIdx.y * kChunk;
    const int nv = min(kChunk, n_views - vbase);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(views + vbase);
        uint32_t* dst = reinterpret_cast<uint32_t*>(sv);
        for (int i = threadIdx.x; i < nv * kViewWords; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile0 = (blockIdx.x * kWarps + warp) * PIX;  // local tile index; global tile = first_tile + local
    const int P = T.width * T.height;

    float w[PIX][3];
    int u1[PIX], v1[PIX];
    bool valid[PIX];
#pragma unroll
    for (int k = 0; k < PIX; ++k) {
        const int p = (first_tile + tile0 + k) * kTile + lane;
        const bool inside = p < P && tile0 + k < n_tiles;
        const float d1 = __fdiv_rn((float)(inside ? __ldg(T.depth + p) : (uint16_t)0), 1000.0f);
        valid[k] = d1 > 0.0f;  // sfm.py:96
        v1[k] = p / T.width;
        u1[k] = p - v1[k] * T.width;
        float c0, c1, c2;
        unproject(T.Kinv, u1[k], v1[k], d1, c0, c1, c2);
        rigid(T.R, T.t, c0, c1, c2, w[k][0], w[k][1], w[k][2]);
    }

    uint32_t mine[PIX];
#pragma unroll
    for (int k = 0; k < PIX; ++k) mine[k] = 0;

    for (int s = 0; s < nv; ++s) {
        const sucre_view& S = sv[s];
#pragma unroll
        for (int k = 0; k < PIX; ++k) {
            int u2, v2;
            bool m = project(S.Ri, S.ti, S.K, S.width, S.height, w[k][0], w[k][1], w[k][2], u2, v2) && valid[k];
            if (m) {
                // backward leg, only at the source pixel the forward leg landed on (sfm.py:124, 154-159)
                const float d2 = __fdiv_rn((float)__ldg(S.depth + (size_t)v2 * S.width + u2), 1000.0f);
                float c0, c1, c2, b0, b1, b2;
                unproject(S.Kinv, u2, v2, d2, c0, c1, c2);
                rigid(S.R, S.t, c0, c1, c2, b0, b1, b2);
                int ub, vb;
                const bool back = project(T.Ri, T.ti, T.K, T.width, T.height, b0, b1, b2, ub, vb);
                m = d2 > 0.0f && back && ub == u1[k] && vb == v1[k];  // sfm.py:96 on S, sfm.py:173
            }
            const uint32_t ballot = __ballot_sync(kFull, m);
            if (lane == s) mine[k] = ballot;
        }
    }
#pragma unroll
    for (int k = 0; k < PIX; ++k)
        if (tile0 + k < n_tiles && lane < nv) masks[(size_t)(tile0 + k) * n_views + vbase + lane] = mine[k];
}

// ---- plan ------------------------------------------------------------------------------------------------
// matches per view: lane <-> view (coalesced rows of masks), warps stride over tiles
__global__ void __launch_bounds__(256)
count_views_kernel(const uint32_t* __restrict__ masks, int n_tiles, int n_views, unsigned long long* __restrict__ view_count) {
    __shared__ unsigned long long part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int view = blockIdx.x * 32 + lane;
    unsigned long long acc = 0;
    if (view < n_views)
        for (int tile = blockIdx.y * 8 + warp; tile < n_tiles; tile += gridDim.y * 8)
            acc += __popc(__ldg(masks + (size_t)tile * n_views + view));
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && view < n_views) {
        for (int i = 1; i < 8; ++i) acc += part[i][lane];
        if (acc) atomicAdd(view_count + view, acc);
    }
}

__global__ void kept_kernel(const long long* __restrict__ view_count, int n_views, double pixels, double min_cover,
                            uint8_t* __restrict__ view_kept) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    // sfm.py:136: len(matches) / (width * height) > min_cover, python floats = IEEE double
    if (v < n_views) view_kept[v] = ((double)view_count[v] / pixels > min_cover) ? 1 : 0;
}

// records, non-empty blocks and segments per tile over kept views: one warp per tile
__global__ void __launch_bounds__(256)
tile_count_kernel(const uint32_t* __restrict__ masks, const uint8_t* __restrict__ view_kept, int n_tiles, int n_views,
                  int seg_views, long long* __restrict__ rec_cnt, long long* __restrict__ blk_cnt,
                  long long* __restrict__ seg_cnt) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    int rec = 0, blk = 0;
    for (int view = lane; view < n_views; view += 32) {
        const uint32_t m = view_kept[view] ? __ldg(masks + (size_t)tile * n_views + view) : 0u;
        rec += __popc(m);
        blk += m != 0;
    }
    for (int o = 16; o; o >>= 1) {
        rec += __shfl_xor_sync(kFull, rec, o);
        blk += __shfl_xor_sync(kFull, blk, o);
    }
    if (lane == 0) {
        rec_cnt[tile] = rec;
        blk_cnt[tile] = blk;
        seg_cnt[tile] = (blk + seg_views - 1) / seg_views;
    }
}

// in-place exclusive scan of three count arrays (n entries -> n+1 offsets), one CTA: rounds of 1024 coalesced
// elements, warp-shuffle scans, running carries
__global__ void __launch_bounds__(1024)
scan_kernel(long long* __restrict__ a, long long* __restrict__ b, long long* __restrict__ c, int n,
            long long* __restrict__ totals) {
    __shared__ long long wsum[3][32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long carry[3] = {0, 0, 0};
    long long* arr[3] = {a, b, c};
    for (int base = 0; base < n; base += 1024) {
        const int i = base + t;
        long long v[3], incl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = i < n ? arr[k][i] : 0;
            incl[k] = v[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long up = __shfl_up_sync(kFull, incl[k], o);
                if (lane >= o) incl[k] += up;
            }
            if (lane == 31) wsum[k][warp] = incl[k];
        }
        __syncthreads();
        if (warp < 3) {  // warp k scans the 32 warp totals of array k
            long long w = wsum[warp][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long up = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += up;
            }
            wsum[warp][lane] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const long long before = carry[k] + (warp ? wsum[k][warp - 1] : 0);
            if (i < n) arr[k][i] = before + incl[k] - v[k];
            carry[k] += wsum[k][31];
        }
        __syncthreads();
    }
    if (t == 0) {
        a[n] = carry[0];
        b[n] = carry[1];
        c[n] = carry[2];
        totals[0] = carry[0];
        totals[1] = carry[1];
        totals[2] = carry[2];
    }
}

// One (target pixel, source view) observation in two steps, so that two of them can overlap their memory
// latency: issue() projects the pixel into the view and starts the gathers at the source pixel it lands on,
// finish() turns the fetched depth / colour into the record.
struct Probe {
    int u2, v2, fmt;
    unsigned d16;
    float raw0, raw1, raw2;
    float Kinv[9];

    __device__ __forceinline__ void issue(const sucre_view* S, float w0, float w1, float w2) {  // S: warp-uniform => broadcast loads
        float Ri[9], ti[3], K[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            Ri[i] = __ldg(&S->Ri[i]);
            K[i] = __ldg(&S->K[i]);
            Kinv[i] = __ldg(&S->Kinv[i]);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) ti[i] = __ldg(&S->ti[i]);
        const int Ws = __ldg(&S->width), Hs = __ldg(&S->height);
        const uint16_t* depth = reinterpret_cast<const uint16_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&S->depth)));
        const void* rgb = reinterpret_cast<const void*>(__ldg(reinterpret_cast<const unsigned long long*>(&S->rgb)));
        fmt = __ldg(&S->rgb_format);
        project(Ri, ti, K, Ws, Hs, w0, w1, w2, u2, v2);
        const size_t q = (size_t)v2 * Ws + u2;
        d16 = __ldg(depth + q);                                                // sfm.py:137
        if (fmt == SUCRE_RGB_F32) {  // resampled on the host in float (--image-scale), loader.py:158-163
            const float* px = reinterpret_cast<const float*>(rgb) + 3 * q;
            raw0 = __ldg(px + 0), raw1 = __ldg(px + 1), raw2 = __ldg(px + 2);
        } else {
            const uint8_t* px = reinterpret_cast<const uint8_t*>(rgb) + 3 * q;
            raw0 = (float)__ldg(px + 0), raw1 = (float)__ldg(px + 1), raw2 = (float)__ldg(px + 2);
        }
    }

    __device__ __forceinline__ void finish(float4* out, uint32_t* src_out, int record_cells) const {
        const float d2 = __fdiv_rn((float)d16, 1000.0f);
        float c0, c1, c2;
        unproject(Kinv, u2, v2, d2, c0, c1, c2);                               // loader.py:113
        // sucre.py:53 cP.norm(dim=0): sequential squares, no fma
        const float z = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2)));
        const bool f32 = fmt == SUCRE_RGB_F32;                                 // else loader.py:157, 87: u8 / 255
        const float I0 = f32 ? raw0 : __fdiv_rn(raw0, 255.0f), I1 = f32 ? raw1 : __fdiv_rn(raw1, 255.0f),
                    I2 = f32 ? raw2 : __fdiv_rn(raw2, 255.0f);
        if (record_cells == 1) {
            out[0] = make_float4(z, I0, I1, I2);
        } else {  // light model: the camera-frame point itself is needed (sucre.py:57)
            out[0] = make_float4(c0, c1, c2, z);
            out[1] = make_float4(I0, I1, I2, 0.f);
        }
        if (src_out) *src_out = (uint32_t)u2 | ((uint32_t)v2 << 16);
    }
};

// ---- sample ----------------------------------------------------------------------------------------------
// One warp per tile.  Phase A compacts the tile's non-empty kept blocks (lane mask + view index) into
// blk_mask / blk_view.  Phase B walks them in segments of seg_views blocks: per-lane record counts -> header
// cells, exclusive scan over lanes -> each lane's first cell, then every matched (pixel, view) is re-projected,
// its source depth + colour fetched, and the record stored in the lane's run (lane-major within the segment).
__global__ void __launch_bounds__(256)
gather_sample_kernel(const __grid_constant__ sucre_view T, const sucre_view* __restrict__ views, int n_views,
                     const uint32_t* __restrict__ masks, const uint8_t* __restrict__ view_kept,
                     const long long* __restrict__ rec_off, const long long* __restrict__ blk_off,
                     const long long* __restrict__ seg_off, int first_tile, int n_tiles, int seg_views, int record_cells,
                     float4* __restrict__ cells, uint32_t* __restrict__ blk_mask, int32_t* __restrict__ blk_view,
                     uint32_t* __restrict__ cell_src) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const uint32_t lt = (1u << lane) - 1u;
    const int P = T.width * T.height;
    const int p = (first_tile + tile) * kTile + lane;
    float w0, w1, w2;
    {
        const float d1 = __fdiv_rn((float)(p < P ? __ldg(T.depth + p) : (uint16_t)0), 1000.0f);
        const int v1 = p / T.width, u1 = p - v1 * T.width;
        float c0, c1, c2;
        unproject(T.Kinv, u1, v1, d1, c0, c1, c2);
        rigid(T.R, T.t, c0, c1, c2, w0, w1, w2);
    }
    // phase A
    const long long blk0 = blk_off[tile];
    int nb = 0;
    for (int base = 0; base < n_views; base += 32) {
        const int view = base + lane;
        uint32_t m = 0;
        if (view < n_views && view_kept[view]) m = __ldg(masks + (size_t)tile * n_views + view);
        const unsigned nz = __ballot_sync(kFull, m != 0);
        if (m != 0) {
            const long long at = blk0 + nb + __popc(nz & lt);
            blk_mask[at] = m;
            blk_view[at] = view;
        }
        nb += __popc(nz);
    }
    __syncwarp();
    // phase B
    long long cell = record_cells * rec_off[tile] + kSegHeaderCells * seg_off[tile];
    for (int s0 = 0; s0 < nb; s0 += seg_views) {
        const int ns = min(seg_views, nb - s0);
        uint32_t bm_l = 0;
        int bv_l = 0;
        if (lane < ns) {
            bm_l = __ldcg(blk_mask + blk0 + s0 + lane);
            bv_l = __ldcg(blk_view + blk0 + s0 + lane);
        }
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < kSegViewsMax; ++j) cnt += (__shfl_sync(kFull, bm_l, j) >> lane) & 1u;  // bm_l = 0 for j >= ns
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += up;
        }
        const int n = __shfl_sync(kFull, incl, 31);
        reinterpret_cast<uint8_t*>(cells + cell)[lane] = (uint8_t)cnt;  // header: 32 lane counts
        long long at = cell + kSegHeaderCells + (long long)record_cells * (incl - cnt);
        // two blocks per step: the gathers of the second are in flight while the first is finished
        for (int j = 0; j < ns; j += 2) {
            const uint32_t bm0 = __shfl_sync(kFull, bm_l, j), bm1 = __shfl_sync(kFull, bm_l, j + 1);  // bm_l = 0 beyond ns
            const int s0v = __shfl_sync(kFull, bv_l, j), s1v = __shfl_sync(kFull, bv_l, j + 1);
            const bool a0 = (bm0 >> lane) & 1u, a1 = (bm1 >> lane) & 1u;
            Probe p0, p1;
            if (a0) p0.issue(views + s0v, w0, w1, w2);
            if (a1) p1.issue(views + s1v, w0, w1, w2);
            if (a0) {
                p0.finish(cells + at, cell_src ? cell_src + at : nullptr, record_cells);
                at += record_cells;
            }
            if (a1) {
                p1.finish(cells + at, cell_src ? cell_src + at : nullptr, record_cells);
                at += record_cells;
            }
        }
        cell += kSegHeaderCells + (long long)record_cells * n;
    }
}

static int check_view_host(const sucre_view* v, const char* who) {
    SUCRE_REQUIRE(v != nullptr, "%s: null view", who);
    SUCRE_REQUIRE(v->width > 0 && v->height > 0 && v->width <= 32767 && v->height <= 32767,
                  "%s: image size %dx%d outside [1, 32767] (the reference stores pixel indices as int16, loader.py:71-74)",
                  who, v->width, v->height);
    SUCRE_REQUIRE((long long)v->width * v->height <= 0x7fffffffLL - 64, "%s: too many pixels", who);
    SUCRE_REQUIRE(v->depth != nullptr, "%s: null depth pointer", who);
    return 0;
}

}  // namespace sucre

using namespace sucre;

static int check_tile_range(const sucre_view* t, int first_tile, int n_tiles, const char* who) {
    const int total = (t->width * t->height + kTile - 1) / kTile;
    SUCRE_REQUIRE(first_tile >= 0 && n_tiles > 0 && first_tile + n_tiles <= total,
                  "%s: tiles [%d, %d) outside the target's %d tiles", who, first_tile, first_tile + n_tiles, total);
    return 0;
}

extern "C" int sucre_gather_match(const sucre_view* target_host, const sucre_view* views, int n_views, int first_tile,
                                  int n_tiles, uint32_t* masks, void* stream) {
    clear_error();
    if (check_view_host(target_host, "sucre_gather_match(target)")) return 1;
    SUCRE_REQUIRE(views && masks, "sucre_gather_match: null pointer");
    SUCRE_REQUIRE(n_views > 0, "sucre_gather_match: n_views = %d", n_views);
    if (check_tile_range(target_host, first_tile, n_tiles, "sucre_gather_match")) return 1;
    constexpr int PIX = SUCRE_MATCH_PIX;
    dim3 grid((n_tiles + kWarps * PIX - 1) / (kWarps * PIX), (n_views + kChunk - 1) / kChunk);
    gather_match_kernel<PIX><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(*target_host, views, n_views, masks, first_tile, n_tiles);
    return check_launch("gather_match_kernel");
}

extern "C" int sucre_gather_count(const uint32_t* masks, int n_tiles, int n_views, int64_t* view_count, void* stream) {
    clear_error();
    SUCRE_REQUIRE(masks && view_count, "sucre_gather_count: null pointer");
    SUCRE_REQUIRE(n_tiles > 0 && n_views > 0, "sucre_gather_count: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    SUCRE_CUDA(cudaMemsetAsync(view_count, 0, sizeof(int64_t) * n_views, st));
    dim3 cgrid((n_views + 31) / 32, min(64, (n_tiles + 7) / 8));
    count_views_kernel<<<cgrid, 256, 0, st>>>(masks, n_tiles, n_views, (unsigned long long*)view_count);
    return check_launch("count_views_kernel");
}

extern "C" int sucre_gather_plan(const uint32_t* masks, int n_tiles, int n_views, const int64_t* view_count,
                                 int64_t target_pixels, double min_cover, int seg_views, uint8_t* view_kept, int64_t* rec_off,
                                 int64_t* blk_off, int64_t* seg_off, int64_t* totals, void* stream) {
    clear_error();
    SUCRE_REQUIRE(masks && view_count && view_kept && rec_off && blk_off && seg_off && totals, "sucre_gather_plan: null pointer");
    SUCRE_REQUIRE(n_tiles > 0 && n_views > 0 && target_pixels > 0, "sucre_gather_plan: bad sizes");
    SUCRE_REQUIRE(seg_views >= 1 && seg_views <= kSegViewsMax, "sucre_gather_plan: seg_views %d outside [1, %d]", seg_views, kSegViewsMax);
    cudaStream_t st = (cudaStream_t)stream;
    kept_kernel<<<(n_views + 127) / 128, 128, 0, st>>>((const long long*)view_count, n_views, (double)target_pixels, min_cover, view_kept);
    tile_count_kernel<<<(n_tiles + 7) / 8, 256, 0, st>>>(masks, view_kept, n_tiles, n_views, seg_views, (long long*)rec_off,
                                                         (long long*)blk_off, (long long*)seg_off);
    scan_kernel<<<1, 1024, 0, st>>>((long long*)rec_off, (long long*)blk_off, (long long*)seg_off, n_tiles, (long long*)totals);
    return check_launch("sucre_gather_plan kernels");
}

extern "C" int sucre_gather_sample(const sucre_view* target_host, const sucre_view* views, int n_views, int first_tile,
                                   int n_tiles, const uint32_t* masks, const uint8_t* view_kept, const int64_t* rec_off,
                                   const int64_t* blk_off, const int64_t* seg_off, int seg_views, int record_cells,
                                   float* cells, uint32_t* blk_mask, int32_t* blk_view, uint32_t* cell_src, void* stream) {
    clear_error();
    if (check_view_host(target_host, "sucre_gather_sample(target)")) return 1;
    SUCRE_REQUIRE(views && masks && view_kept && rec_off && blk_off && seg_off && cells && blk_mask && blk_view,
                  "sucre_gather_sample: null pointer");
    if (check_tile_range(target_host, first_tile, n_tiles, "sucre_gather_sample")) return 1;
    SUCRE_REQUIRE((reinterpret_cast<uintptr_t>(cells) & 15) == 0, "sucre_gather_sample: cells must be 16-byte aligned");
    SUCRE_REQUIRE(seg_views >= 1 && seg_views <= kSegViewsMax && (record_cells == 1 || record_cells == 2),
                  "sucre_gather_sample: seg_views %d / record_cells %d not supported", seg_views, record_cells);
    gather_sample_kernel<<<(n_tiles + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        *target_host, views, n_views, masks, view_kept, (const long long*)rec_off, (const long long*)blk_off,
        (const long long*)seg_off, first_tile, n_tiles, seg_views, record_cells, reinterpret_cast<float4*>(cells), blk_mask,
        blk_view, cell_src);
    return check_launch("gather_sample_kernel");
}
