// Stage 1 — fused multi-view correspondence gather for sm_100a.
//
// What the reference does with ~80 whole-image ATen kernels, a reverse index map and four stream
// compactions per (target, view) pair (sfm.py:115-138, 154-175; 306 B of intermediates per pixel-view)
// is done here per target pixel, in registers:
//   match   one thread owns PIX target pixels, keeps their world points, and walks a chunk of 32 source
//           views whose constants sit in shared memory.  A conservative per-(warp, view) frustum test first
//           drops the views in which none of the warp's pixels can land (lane j tests view j).  For the rest
//           the backward projection is evaluated only at the source pixel the forward projection lands on
//           (one 2-byte gather), and the per-(tile, view) result is a 32-bit ballot mask.
//   plan    per-view counts -> min_cover decision -> per-tile record/block/row counts -> exclusive scan.
//   sample  one warp per tile re-projects only the matched pixels, fetches depth + colour of the source
//           pixel and writes the records into the tile's ELL rows (include/sucre_b200.h): row j, lane i = the
//           j-th observation of pixel i, zero sentinels below the shorter columns.
//
// Arithmetic contract (SURVEY.md §8a', pinned by tests against the reference's own outputs): every fp32
// operation below is a single correctly rounded IEEE op written with an explicit intrinsic, in the order the
// reference's ATen calls perform them; this file is additionally compiled with -fmad=false.
#include <cfloat>
#include <climits>

#include "common.cuh"

namespace sucre {

// y = M (3x3 row-major) applied to x the way torch.mm(3x3, 3xn) rounds it on the reference's CPU path:
// per row fma(m2, x2, fma(m1, x1, m0*x0)).
// SPARSE: M has the PINHOLE pattern [[a,0,b],[0,c,d],[0,0,1]] exactly (SUCRE_VIEW_*_SPARSE, checked by the host on
// the values).  fma(0, x1, m0*x0) == m0*x0, fma(m4, x1, 0*x0) == m4*x1 and fma(1, x2, 0) == x2 for finite x, so the
// skipped operations cannot change a rounded result; for a non-finite x both forms yield a non-finite (rejected)
// projection.
template <bool SPARSE>
__device__ __forceinline__ void mat3(const float* M, float x0, float x1, float x2, float& y0, float& y1, float& y2) {
    if (SPARSE) {
        y0 = __fmaf_rn(M[2], x2, __fmul_rn(M[0], x0));
        y1 = __fmaf_rn(M[5], x2, __fmul_rn(M[4], x1));
        y2 = x2;
    } else {
        y0 = __fmaf_rn(M[2], x2, __fmaf_rn(M[1], x1, __fmul_rn(M[0], x0)));
        y1 = __fmaf_rn(M[5], x2, __fmaf_rn(M[4], x1, __fmul_rn(M[3], x0)));
        y2 = __fmaf_rn(M[8], x2, __fmaf_rn(M[7], x1, __fmul_rn(M[6], x0)));
    }
}

// sfm.py:90-93 unproject_depth with depth = u16 / 1000 (loader.py:167): cP = Kinv @ (d * (u+.5, v+.5, 1))
template <bool SPARSE>
__device__ __forceinline__ void unproject(const float* Kinv, int u, int v, float d, float& c0, float& c1, float& c2) {
    const float x0 = __fmul_rn(d, __fadd_rn((float)u, 0.5f));
    const float x1 = __fmul_rn(d, __fadd_rn((float)v, 0.5f));
    mat3<SPARSE>(Kinv, x0, x1, d, c0, c1, c2);
}

// sfm.py:49-55 Pose.transform: (R @ P) + t, the add rounded separately
__device__ __forceinline__ void rigid(const float* R, const float* t, float x0, float x1, float x2, float& y0, float& y1, float& y2) {
    mat3<false>(R, x0, x1, x2, y0, y1, y2);
    y0 = __fadd_rn(y0, t[0]);
    y1 = __fadd_rn(y1, t[1]);
    y2 = __fadd_rn(y2, t[2]);
}

// sfm.py:103-107 project_to_view followed by sfm.py:116-117: .long() truncates toward zero, then
// 0 <= u < W, 0 <= v < H.  trunc(x) >= 0 <=> x > -1, so (-1,0) is accepted as index 0 exactly like the
// reference; NaN / inf / huge fail the comparisons (the reference's INT64_MIN fails `0 <=`).
template <bool SPARSE>
__device__ __forceinline__ bool project(const float* Ri, const float* ti, const float* K, int W, int H,
                                        float w0, float w1, float w2, int& u, int& v) {
    float c0, c1, c2, p0, p1, p2;
    rigid(Ri, ti, w0, w1, w2, c0, c1, c2);
    mat3<SPARSE>(K, c0, c1, c2, p0, p1, p2);
    const float px = __fdiv_rn(p0, p2), py = __fdiv_rn(p1, p2);
    const bool in = px > -1.0f && px < (float)W && py > -1.0f && py < (float)H;
    u = __float2int_rz(px);
    v = __float2int_rz(py);
    return in;
}

constexpr int kViewWords = sizeof(sucre_view) / 4;  // 52
constexpr int kChunk = 32;                          // source views per CTA: lane j keeps the mask of view j
constexpr int kWarps = 8;
#ifndef SUCRE_MATCH_PIX
#define SUCRE_MATCH_PIX 2  // target pixels per thread of gather_match_kernel
#endif

// the reference's two-way test of one (target pixel, source view) pair: forward leg T -> S, backward leg only at the
// source pixel the forward leg landed on (sfm.py:124, 154-159, 171-175)
template <bool SPARSE>
__device__ __forceinline__ bool two_way(const sucre_view& T, const sucre_view& S, float w0, float w1, float w2, int u1, int v1,
                                        bool valid, bool& fwd_in) {
    int u2, v2;
    bool m = project<SPARSE>(S.Ri, S.ti, S.K, S.width, S.height, w0, w1, w2, u2, v2) && valid;
    fwd_in = m;
    if (m) {
        const float d2 = __fdiv_rn((float)__ldg(S.depth + (size_t)v2 * S.width + u2), 1000.0f);
        float c0, c1, c2, b0, b1, b2;
        unproject<SPARSE>(S.Kinv, u2, v2, d2, c0, c1, c2);
        rigid(S.R, S.t, c0, c1, c2, b0, b1, b2);
        int ub, vb;
        const bool back = project<SPARSE>(T.Ri, T.ti, T.K, T.width, T.height, b0, b1, b2, ub, vb);
        m = d2 > 0.0f && back && ub == u1 && vb == v1;  // sfm.py:96 on S, sfm.py:173
    }
    return m;
}

// Conservative frustum test of one source view against the world-space slab that contains every valid target
// pixel of a warp (corners wc[8]: extreme pixel centres x {smallest, largest} depth).  In exact arithmetic the slab
// is convex and, when all eight corners are in front of the source camera, projects inside the hull of their
// projections.  The kernels' fp32 projection of a pixel differs from the exact one by at most
// rel * (fx + |px - cx|) with rel = E / z (E bounds the absolute error of the camera-frame coordinates), so a
// view is dropped only if the hull misses the image by more than that on one side.  Returns true = keep.
__device__ __forceinline__ bool may_land(const sucre_view& S, const float (*wc)[3], float w1max, float tT1) {
    if (!(S.flags & SUCRE_VIEW_K_SPARSE)) return true;
    const float fx = S.K[0], cx = S.K[2], fy = S.K[4], cy = S.K[5];
    float xmin = FLT_MAX, xmax = -FLT_MAX, ymin = FLT_MAX, ymax = -FLT_MAX, zmin = FLT_MAX;
    bool finite = true;  // fminf / fmaxf drop NaNs silently: track them
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float c0, c1, c2;
        rigid(S.Ri, S.ti, wc[c][0], wc[c][1], wc[c][2], c0, c1, c2);
        const float px = __fdiv_rn(__fmaf_rn(cx, c2, __fmul_rn(fx, c0)), c2);
        const float py = __fdiv_rn(__fmaf_rn(cy, c2, __fmul_rn(fy, c1)), c2);
        xmin = fminf(xmin, px), xmax = fmaxf(xmax, px);
        ymin = fminf(ymin, py), ymax = fmaxf(ymax, py);
        zmin = fminf(zmin, c2);
        finite = finite && px == px && py == py && c2 == c2;
    }
    const float E = 2e-6f * (2.0f * w1max + tT1 + fabsf(S.ti[0]) + fabsf(S.ti[1]) + fabsf(S.ti[2]));
    const float rel = E / zmin;
    if (!(finite && zmin > 0.0f && rel <= 1e-3f)) return true;  // a corner at / behind the camera plane (or NaN): no claim
    const float W = (float)S.width, H = (float)S.height;
    const float mx = 1.0f + 4.0f * rel * (fabsf(fx) + fmaxf(fabsf(xmin), fabsf(xmax)) + W);
    const float my = 1.0f + 4.0f * rel * (fabsf(fy) + fmaxf(fabsf(ymin), fabsf(ymax)) + H);
    const bool out = (xmax + mx < -1.0f) || (xmin - mx > W) || (ymax + my < -1.0f) || (ymin - my > H);
    return !out;  // NaN bounds compare false: keep
}

template <int PIX>
__global__ void __launch_bounds__(kWarps * 32)
gather_match_kernel(const __grid_constant__ sucre_view T, const sucre_view* __restrict__ views, int n_views,
                    uint32_t* __restrict__ masks, const sucre_band band, unsigned long long* __restrict__ stats, int cull) {
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
    const int n_tiles = band.n_tiles;
    const int tile0 = (blockIdx.x * kWarps + warp) * PIX;  // local tile index; global tile = band_tile(band, local)
    if (tile0 >= n_tiles) return;
    const int P = T.width * T.height;
    const bool t_sparse = (T.flags & SUCRE_VIEW_K_SPARSE) != 0;

    float w[PIX][3];
    int u1[PIX], v1[PIX];
    bool valid[PIX];
    // extent of the warp's valid pixels: pixel-index box and depth range (positive floats order like their bit patterns)
    int ulo = INT_MAX, uhi = -1, vlo = INT_MAX, vhi = -1;
    unsigned dlo = 0x7f800000u, dhi = 0u;
#pragma unroll
    for (int k = 0; k < PIX; ++k) {
        const bool here = tile0 + k < n_tiles;
        const int p = (here ? band_tile(band, tile0 + k) : 0) * kTile + lane;
        const bool inside = here && p < P;
        const float d1 = __fdiv_rn((float)(inside ? __ldg(T.depth + p) : (uint16_t)0), 1000.0f);
        valid[k] = d1 > 0.0f;  // sfm.py:96
        v1[k] = p / T.width;
        u1[k] = p - v1[k] * T.width;
        float c0, c1, c2;
        unproject<false>(T.Kinv, u1[k], v1[k], d1, c0, c1, c2);
        rigid(T.R, T.t, c0, c1, c2, w[k][0], w[k][1], w[k][2]);
        if (valid[k]) {
            ulo = min(ulo, u1[k]), uhi = max(uhi, u1[k]);
            vlo = min(vlo, v1[k]), vhi = max(vhi, v1[k]);
            dlo = min(dlo, __float_as_uint(d1)), dhi = max(dhi, __float_as_uint(d1));
        }
    }

    const unsigned chunk_mask = nv == 32 ? kFull : ((1u << nv) - 1u);
    unsigned alive = chunk_mask;
    if (cull) {
        ulo = __reduce_min_sync(kFull, ulo), uhi = __reduce_max_sync(kFull, uhi);
        vlo = __reduce_min_sync(kFull, vlo), vhi = __reduce_max_sync(kFull, vhi);
        dlo = __reduce_min_sync(kFull, dlo), dhi = __reduce_max_sync(kFull, dhi);
        bool keep = false;
        if (uhi >= 0) {  // the warp has a valid pixel (warp-uniform)
            float wc[8][3], w1max = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float c0, c1, c2;
                unproject<false>(T.Kinv, (c & 1) ? uhi : ulo, (c & 2) ? vhi : vlo, __uint_as_float((c & 4) ? dhi : dlo), c0, c1, c2);
                rigid(T.R, T.t, c0, c1, c2, wc[c][0], wc[c][1], wc[c][2]);
                w1max = fmaxf(w1max, fabsf(wc[c][0]) + fabsf(wc[c][1]) + fabsf(wc[c][2]));
            }
            const float tT1 = fabsf(T.t[0]) + fabsf(T.t[1]) + fabsf(T.t[2]);
            keep = lane < nv ? may_land(sv[lane], wc, w1max, tT1) : false;
        }
        alive = __ballot_sync(kFull, keep) & chunk_mask;
    }

    uint32_t mine[PIX];
#pragma unroll
    for (int k = 0; k < PIX; ++k) mine[k] = 0;
    unsigned n_in = 0;

    for (unsigned todo = alive; todo; todo &= todo - 1) {
        const int s = __ffs(todo) - 1;
        const sucre_view& S = sv[s];
        const bool sparse = t_sparse && (S.flags & (SUCRE_VIEW_K_SPARSE | SUCRE_VIEW_KINV_SPARSE)) == (SUCRE_VIEW_K_SPARSE | SUCRE_VIEW_KINV_SPARSE);
#pragma unroll
        for (int k = 0; k < PIX; ++k) {
            bool fwd;
            const bool m = sparse ? two_way<true>(T, S, w[k][0], w[k][1], w[k][2], u1[k], v1[k], valid[k], fwd)
                                  : two_way<false>(T, S, w[k][0], w[k][1], w[k][2], u1[k], v1[k], valid[k], fwd);
            const uint32_t ballot = __ballot_sync(kFull, m);
            if (stats) n_in += __popc(__ballot_sync(kFull, fwd));
            if (lane == s) mine[k] = ballot;
        }
    }
    int tiles_here = 0;
#pragma unroll
    for (int k = 0; k < PIX; ++k)
        if (tile0 + k < n_tiles) {
            ++tiles_here;
            if (lane < nv) masks[(size_t)(tile0 + k) * n_views + vbase + lane] = mine[k];
        }
    if (stats && lane == 0) {
        const unsigned culled = __popc(chunk_mask & ~alive) * tiles_here;
        if (culled) atomicAdd(stats + 0, (unsigned long long)culled);
        if (n_in) atomicAdd(stats + 1, (unsigned long long)n_in);
    }
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

// ---- permute -----------------------------------------------------------------------------------------------
// One CTA per group of SUCRE_GROUP_TILES local tiles (1024 slots).  Thread t starts as natural slot t of the group (warp =
// natural tile): it counts its pixel's matches over all views, the CTA sorts (count descending, natural order among equals)
// and thread t ends as slot t of the PERMUTED group: it records its pixel in pix and the warp rebuilds the mask words of
// its permuted tile from the natural ones, staged in shared memory 256 views at a time.  Pixels with columns of equal
// height end up in the same tile, so the ELL rows of the store hold few sentinels.
constexpr int kGroupTiles = SUCRE_GROUP_TILES;
constexpr int kGroupSlots = kGroupTiles * kTile;
constexpr int kPermViews = 256;   // views staged per pass: 32 rows x 257 words
static_assert(kGroupSlots == 1024, "one thread per slot of a group");

__global__ void __launch_bounds__(kGroupSlots)
permute_kernel(const uint32_t* __restrict__ masks, int n_views, const sucre_band band, long long P, int32_t* __restrict__ pix,
               uint32_t* __restrict__ pmasks) {
    __shared__ uint32_t key[kGroupSlots];
    __shared__ int32_t nat_pix[kGroupSlots];
    __shared__ uint32_t rows[kGroupTiles][kPermViews + 1];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int tile0 = blockIdx.x * kGroupTiles;
    const int k = tile0 + warp;   // natural local tile of this warp, then its permuted tile
    const bool tile_ok = k < band.n_tiles;
    long long p = tile_ok ? (long long)band_tile(band, k) * kTile + lane : -1;
    if (p >= P) p = -1;
    int cnt = 0;
    if (tile_ok) {
        for (int base = 0; base < n_views; base += 32) {
            const uint32_t m = base + lane < n_views ? __ldg(masks + (size_t)k * n_views + base + lane) : 0u;
            for (unsigned nz = __ballot_sync(kFull, m != 0); nz; nz &= nz - 1)
                cnt += (__shfl_sync(kFull, m, __ffs(nz) - 1) >> lane) & 1u;
        }
    }
    nat_pix[t] = (int32_t)p;
    // ascending sort of {0x1fffff - count : 21 | natural slot : 10} (< 2^31); slots without a pixel go last
    key[t] = p < 0 ? 0xffffffffu : ((0x1fffffu - (uint32_t)cnt) << 10) | (uint32_t)t;
    __syncthreads();
    for (int size = 2; size <= kGroupSlots; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int other = t ^ stride;
            if (other > t) {
                const uint32_t a = key[t], b = key[other];
                const bool up = (t & size) == 0;
                if ((a > b) == up) {
                    key[t] = b;
                    key[other] = a;
                }
            }
            __syncthreads();
        }
    }
    const uint32_t mine = key[t];
    const int src = mine == 0xffffffffu ? -1 : (int)(mine & 1023u);   // natural slot whose pixel this slot takes
    if (tile_ok) pix[(size_t)k * kTile + lane] = src < 0 ? -1 : nat_pix[src];
    const int src_row = src < 0 ? 0 : src >> 5, src_lane = src & 31;
    for (int v0 = 0; v0 < n_views; v0 += kPermViews) {
        const int nv = min(kPermViews, n_views - v0);
        __syncthreads();   // the previous pass is done with `rows`
        for (int r = warp; r < kGroupTiles; r += kGroupSlots / 32)   // (one row per warp) coalesced over the views
            for (int c = lane; c < nv; c += 32)
                rows[r][c] = tile0 + r < band.n_tiles ? __ldg(masks + (size_t)(tile0 + r) * n_views + v0 + c) : 0u;
        __syncthreads();
        if (tile_ok) {
            for (int base = 0; base < nv; base += 32) {
                uint32_t out = 0;
                const int m = min(32, nv - base);
                for (int j = 0; j < m; ++j) {
                    const unsigned b = __ballot_sync(kFull, src >= 0 && ((rows[src_row][base + j] >> src_lane) & 1u));
                    if (lane == j) out = b;
                }
                if (lane < m) pmasks[(size_t)k * n_views + v0 + base + lane] = out;
            }
        }
    }
}

// sfm.py:136: len(matches) / (width * height) > min_cover, python floats = IEEE double
__device__ __forceinline__ bool view_passes(long long count, double pixels, double min_cover) {
    return (double)count / pixels > min_cover;
}

// records, non-empty blocks and ELL rows per tile over kept views: one warp per tile.  Also publishes the
// min_cover decision (view_kept) — every warp evaluates it for the views it touches, CTA 0 stores it.
__global__ void __launch_bounds__(256)
tile_count_kernel(const uint32_t* __restrict__ masks, const long long* __restrict__ view_count, double pixels, double min_cover,
                  int n_tiles, int n_views, uint8_t* __restrict__ view_kept, long long* __restrict__ rec_cnt,
                  long long* __restrict__ blk_cnt, long long* __restrict__ row_cnt) {
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0)
        for (int v = threadIdx.x; v < n_views; v += blockDim.x) view_kept[v] = view_passes(view_count[v], pixels, min_cover) ? 1 : 0;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    int rec = 0, blk = 0, mine = 0;  // mine: observations of this lane's pixel
    for (int base = 0; base < n_views; base += 32) {
        const int view = base + lane;
        uint32_t m = 0;
        if (view < n_views && view_passes(view_count[view], pixels, min_cover)) m = __ldg(masks + (size_t)tile * n_views + view);
        rec += __popc(m);
        blk += m != 0;
        for (unsigned nz = __ballot_sync(kFull, m != 0); nz; nz &= nz - 1)
            mine += (__shfl_sync(kFull, m, __ffs(nz) - 1) >> lane) & 1u;
    }
    rec = __reduce_add_sync(kFull, rec);
    blk = __reduce_add_sync(kFull, blk);
    const int rows = __reduce_max_sync(kFull, mine);
    if (lane == 0) {
        rec_cnt[tile] = rec;
        blk_cnt[tile] = blk;
        row_cnt[tile] = rows;
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
    long long nxt[3];   // the next round's elements are requested a round ahead: their latency hides behind this round
#pragma unroll
    for (int k = 0; k < 3; ++k) nxt[k] = t < n ? arr[k][t] : 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + t;
        long long v[3], incl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = nxt[k];
            nxt[k] = i + 1024 < n ? arr[k][i + 1024] : 0;
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
    bool sparse;
    unsigned d16;
    uint32_t rgb8;              // SUCRE_RGB_U8: r | g << 8 | b << 16
    float raw0, raw1, raw2;     // SUCRE_RGB_F32
    float Kinv[9];

    __device__ __forceinline__ void issue(const sucre_view* S, float w0, float w1, float w2) {  // S: warp-uniform => broadcast loads
        float Ri[9], ti[3], K[9];
        const int flags = __ldg(&S->flags);
        sparse = (flags & (SUCRE_VIEW_K_SPARSE | SUCRE_VIEW_KINV_SPARSE)) == (SUCRE_VIEW_K_SPARSE | SUCRE_VIEW_KINV_SPARSE);
#pragma unroll
        for (int i = 0; i < 9; ++i) Ri[i] = __ldg(&S->Ri[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) ti[i] = __ldg(&S->ti[i]);
        const int Ws = __ldg(&S->width), Hs = __ldg(&S->height);
        if (sparse) {
            K[0] = __ldg(&S->K[0]), K[2] = __ldg(&S->K[2]), K[4] = __ldg(&S->K[4]), K[5] = __ldg(&S->K[5]);
            Kinv[0] = __ldg(&S->Kinv[0]), Kinv[2] = __ldg(&S->Kinv[2]), Kinv[4] = __ldg(&S->Kinv[4]), Kinv[5] = __ldg(&S->Kinv[5]);
            project<true>(Ri, ti, K, Ws, Hs, w0, w1, w2, u2, v2);
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                K[i] = __ldg(&S->K[i]);
                Kinv[i] = __ldg(&S->Kinv[i]);
            }
            project<false>(Ri, ti, K, Ws, Hs, w0, w1, w2, u2, v2);
        }
        const uint16_t* depth = reinterpret_cast<const uint16_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&S->depth)));
        const void* rgb = reinterpret_cast<const void*>(__ldg(reinterpret_cast<const unsigned long long*>(&S->rgb)));
        fmt = __ldg(&S->rgb_format);
        const size_t q = (size_t)v2 * Ws + u2;
        d16 = __ldg(depth + q);                                                // sfm.py:137
        if (fmt == SUCRE_RGB_F32) {  // resampled on the host in float (--image-scale), loader.py:158-163
            const float* px = reinterpret_cast<const float*>(rgb) + 3 * q;
            raw0 = __ldg(px + 0), raw1 = __ldg(px + 1), raw2 = __ldg(px + 2);
        } else {
            const uint8_t* px = reinterpret_cast<const uint8_t*>(rgb) + 3 * q;
            rgb8 = (uint32_t)__ldg(px + 0) | ((uint32_t)__ldg(px + 1) << 8) | ((uint32_t)__ldg(px + 2) << 16);
        }
    }

    // writes the record at slot `at` of a store of the given format (see include/sucre_b200.h)
    __device__ __forceinline__ void finish(void* cells, long long at, int format, uint32_t* src_out) const {
        const float d2 = __fdiv_rn((float)d16, 1000.0f);
        float c0, c1, c2;
        if (sparse) unproject<true>(Kinv, u2, v2, d2, c0, c1, c2);             // loader.py:113
        else unproject<false>(Kinv, u2, v2, d2, c0, c1, c2);
        // sucre.py:53 cP.norm(dim=0): sequential squares, no fma
        const float z = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(c0, c0), __fmul_rn(c1, c1)), __fmul_rn(c2, c2)));
        if (format == SUCRE_REC_Z_U8) {
            reinterpret_cast<uint2*>(cells)[at] = make_uint2(__float_as_uint(z), rgb8);
        } else if (format == SUCRE_REC_P_U8) {
            reinterpret_cast<float4*>(cells)[at] = make_float4(c0, c1, c2, __uint_as_float(rgb8));
        } else {
            float I0, I1, I2;
            if (fmt == SUCRE_RGB_F32) {
                I0 = raw0, I1 = raw1, I2 = raw2;
            } else {  // loader.py:157, 87: u8 / 255
                I0 = __fdiv_rn((float)(rgb8 & 0xffu), 255.0f);
                I1 = __fdiv_rn((float)((rgb8 >> 8) & 0xffu), 255.0f);
                I2 = __fdiv_rn((float)((rgb8 >> 16) & 0xffu), 255.0f);
            }
            if (format == SUCRE_REC_Z_F32) {
                reinterpret_cast<float4*>(cells)[at] = make_float4(z, I0, I1, I2);
            } else {  // SUCRE_REC_P_F32: row-major pairs of cells, slot `at` = cells 2*at, 2*at+1
                reinterpret_cast<float4*>(cells)[2 * at] = make_float4(c0, c1, c2, z);
                reinterpret_cast<float4*>(cells)[2 * at + 1] = make_float4(I0, I1, I2, 0.f);
            }
        }
        if (src_out) src_out[at] = (uint32_t)u2 | ((uint32_t)v2 << 16);
    }
};

__device__ __forceinline__ void write_sentinel(void* cells, long long at, int format, uint32_t* src_out) {
    if (format == SUCRE_REC_Z_U8) {
        reinterpret_cast<uint2*>(cells)[at] = make_uint2(0u, 0u);
    } else if (format == SUCRE_REC_P_F32) {
        reinterpret_cast<float4*>(cells)[2 * at] = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(cells)[2 * at + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        reinterpret_cast<float4*>(cells)[at] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (src_out) src_out[at] = 0xffffffffu;
}

// ---- sample ----------------------------------------------------------------------------------------------
// One warp per tile walks the tile's non-empty kept blocks in pairing-list order (lane <-> view over 32-view chunks
// of the mask row); every matched (pixel, view) is re-projected, its source depth + colour fetched, and the record
// stored at row (lane's running count), column lane — lanes of a warp write the same or neighbouring rows, so the
// stores coalesce.  The block list (lane mask + view) is written alongside for export / parity checks.
__global__ void __launch_bounds__(256)
gather_sample_kernel(const __grid_constant__ sucre_view T, const sucre_view* __restrict__ views, int n_views,
                     const uint32_t* __restrict__ masks, const uint8_t* __restrict__ view_kept,
                     const long long* __restrict__ row_off, const long long* __restrict__ blk_off, const sucre_band band,
                     const int32_t* __restrict__ pix, int format, void* __restrict__ cells, uint32_t* __restrict__ blk_mask, int32_t* __restrict__ blk_view,
                     uint32_t* __restrict__ cell_src) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= band.n_tiles) return;
    const uint32_t lt = (1u << lane) - 1u;
    const int P = T.width * T.height;
    int p = pix ? __ldg(pix + (size_t)tile * kTile + lane) : band_tile(band, tile) * kTile + lane;
    if (p < 0) p = P;   // a slot without a pixel: depth 0, never matched
    float w0, w1, w2;
    {
        const float d1 = __fdiv_rn((float)(p < P ? __ldg(T.depth + p) : (uint16_t)0), 1000.0f);
        const int v1 = p / T.width, u1 = p - v1 * T.width;
        float c0, c1, c2;
        unproject<false>(T.Kinv, u1, v1, d1, c0, c1, c2);
        rigid(T.R, T.t, c0, c1, c2, w0, w1, w2);
    }
    const long long row0 = row_off[tile];
    const int n_rows = (int)(row_off[tile + 1] - row0);
    long long blk_at = blk_off[tile];
    long long at = row0 * kTile + lane;  // slot of this lane's next record
    int mine = 0;
    for (int base = 0; base < n_views; base += 32) {
        const int view = base + lane;
        uint32_t m = 0;
        if (view < n_views && view_kept[view]) m = __ldg(masks + (size_t)tile * n_views + view);
        unsigned nz = __ballot_sync(kFull, m != 0);
        if (m != 0) {
            const long long b = blk_at + __popc(nz & lt);
            blk_mask[b] = m;
            blk_view[b] = view;
        }
        blk_at += __popc(nz);
        // two blocks per step: the gathers of the second are in flight while the first is finished
        while (nz) {
            const int j0 = __ffs(nz) - 1;
            nz &= nz - 1;
            const int j1 = nz ? __ffs(nz) - 1 : j0;
            const bool two = nz != 0;
            nz &= nz - 1;
            const uint32_t bm0 = __shfl_sync(kFull, m, j0), bm1 = __shfl_sync(kFull, m, j1);
            const bool a0 = (bm0 >> lane) & 1u, a1 = two && ((bm1 >> lane) & 1u);
            Probe p0, p1;
            if (a0) p0.issue(views + base + j0, w0, w1, w2);
            if (a1) p1.issue(views + base + j1, w0, w1, w2);
            if (a0) {
                p0.finish(cells, at, format, cell_src);
                at += kTile;
                ++mine;
            }
            if (a1) {
                p1.finish(cells, at, format, cell_src);
                at += kTile;
                ++mine;
            }
        }
    }
    for (; mine < n_rows; ++mine, at += kTile) write_sentinel(cells, at, format, cell_src);
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

static bool match_cull_enabled() {  // SUCRE_MATCH_CULL=0 switches the frustum pre-test off (A/B timing; results are identical)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SUCRE_MATCH_CULL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

}  // namespace sucre

using namespace sucre;

static int check_band(const sucre_view* t, const sucre_band* b, const char* who) {
    const int total = (t->width * t->height + kTile - 1) / kTile;
    SUCRE_REQUIRE(b != nullptr, "%s: null band", who);
    SUCRE_REQUIRE(b->first_tile >= 0 && b->n_tiles > 0 && b->chunk_tiles > 0 && b->stride_tiles >= 0,
                  "%s: bad band {%d, %d, %d, %d}", who, b->first_tile, b->n_tiles, b->chunk_tiles, b->stride_tiles);
    SUCRE_REQUIRE(b->n_tiles <= b->chunk_tiles || b->stride_tiles >= b->chunk_tiles, "%s: the chunks of a band must not overlap", who);
    SUCRE_REQUIRE(band_tile(*b, b->n_tiles - 1) < total, "%s: band {%d, %d, %d, %d} reaches tile %d of the target's %d tiles", who,
                  b->first_tile, b->n_tiles, b->chunk_tiles, b->stride_tiles, band_tile(*b, b->n_tiles - 1), total);
    return 0;
}

// one float of the band's J -> its place in every destination image
struct ScatterPtrs {
    unsigned long long p[SUCRE_MAX_PEERS];
};
__global__ void __launch_bounds__(256)
scatter_J_kernel(const float* __restrict__ J_band, const sucre_band band, const int32_t* __restrict__ pix, long long floats_local,
                 long long floats_total, const ScatterPtrs dst, int n_dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= floats_local) return;
    const long long px = i / 3;
    const long long gp = pix ? (long long)pix[px] : (long long)band_tile(band, (int)(px / kTile)) * kTile + px % kTile;
    const long long g = gp * 3 + i % 3;
    if (gp < 0 || g >= floats_total) return;  // a slot without a pixel / beyond the last pixel of the image (partial last tile)
    const float v = J_band[i];
    for (int d = 0; d < n_dst; ++d) reinterpret_cast<float*>(dst.p[d])[g] = v;
}

extern "C" int sucre_record_bytes(int record_format) {
    switch (record_format) {
        case SUCRE_REC_Z_U8: return 8;
        case SUCRE_REC_Z_F32: return 16;
        case SUCRE_REC_P_U8: return 16;
        case SUCRE_REC_P_F32: return 32;
        default: return 0;
    }
}

extern "C" int sucre_gather_match(const sucre_view* target_host, const sucre_view* views, int n_views, const sucre_band* band_host,
                                  uint32_t* masks, int64_t* stats, void* stream) {
    clear_error();
    if (check_view_host(target_host, "sucre_gather_match(target)")) return 1;
    SUCRE_REQUIRE(views && masks, "sucre_gather_match: null pointer");
    SUCRE_REQUIRE(n_views > 0, "sucre_gather_match: n_views = %d", n_views);
    if (check_band(target_host, band_host, "sucre_gather_match")) return 1;
    constexpr int PIX = SUCRE_MATCH_PIX;
    const int n_tiles = band_host->n_tiles;
    dim3 grid((n_tiles + kWarps * PIX - 1) / (kWarps * PIX), (n_views + kChunk - 1) / kChunk);
    gather_match_kernel<PIX><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(*target_host, views, n_views, masks, *band_host,
                                                                             (unsigned long long*)stats, match_cull_enabled() ? 1 : 0);
    return check_launch("gather_match_kernel");
}

extern "C" int sucre_gather_permute(const uint32_t* masks, int n_views, const sucre_band* band_host, int64_t target_pixels,
                                    int32_t* pix, uint32_t* pmasks, void* stream) {
    clear_error();
    SUCRE_REQUIRE(masks && band_host && pix && pmasks, "sucre_gather_permute: null pointer");
    SUCRE_REQUIRE(masks != pmasks, "sucre_gather_permute: pmasks must not alias masks");
    SUCRE_REQUIRE(n_views > 0 && n_views < (1 << 21) && target_pixels > 0 && band_host->n_tiles > 0 && band_host->chunk_tiles > 0,
                  "sucre_gather_permute: bad sizes");
    SUCRE_REQUIRE((long long)band_tile(*band_host, band_host->n_tiles - 1) * kTile < target_pixels, "sucre_gather_permute: band outside the image");
    const int groups = (band_host->n_tiles + kGroupTiles - 1) / kGroupTiles;
    permute_kernel<<<groups, kGroupSlots, 0, (cudaStream_t)stream>>>(masks, n_views, *band_host, (long long)target_pixels, pix, pmasks);
    return check_launch("permute_kernel");
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
                                 int64_t target_pixels, double min_cover, uint8_t* view_kept, int64_t* rec_off, int64_t* blk_off,
                                 int64_t* row_off, int64_t* totals, void* stream) {
    clear_error();
    SUCRE_REQUIRE(masks && view_count && view_kept && rec_off && blk_off && row_off && totals, "sucre_gather_plan: null pointer");
    SUCRE_REQUIRE(n_tiles > 0 && n_views > 0 && target_pixels > 0, "sucre_gather_plan: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    tile_count_kernel<<<(n_tiles + 7) / 8, 256, 0, st>>>(masks, (const long long*)view_count, (double)target_pixels, min_cover, n_tiles,
                                                         n_views, view_kept, (long long*)rec_off, (long long*)blk_off, (long long*)row_off);
    scan_kernel<<<1, 1024, 0, st>>>((long long*)rec_off, (long long*)blk_off, (long long*)row_off, n_tiles, (long long*)totals);
    return check_launch("sucre_gather_plan kernels");
}

extern "C" int sucre_gather_sample(const sucre_view* target_host, const sucre_view* views, int n_views, const sucre_band* band_host,
                                   const int32_t* pix, const uint32_t* masks, const uint8_t* view_kept, const int64_t* row_off,
                                   const int64_t* blk_off, int record_format, void* cells, uint32_t* blk_mask, int32_t* blk_view,
                                   uint32_t* cell_src, void* stream) {
    clear_error();
    if (check_view_host(target_host, "sucre_gather_sample(target)")) return 1;
    SUCRE_REQUIRE(views && masks && view_kept && row_off && blk_off && cells && blk_mask && blk_view,
                  "sucre_gather_sample: null pointer");
    if (check_band(target_host, band_host, "sucre_gather_sample")) return 1;
    const int n_tiles = band_host->n_tiles;
    SUCRE_REQUIRE((reinterpret_cast<uintptr_t>(cells) & 15) == 0, "sucre_gather_sample: cells must be 16-byte aligned");
    SUCRE_REQUIRE(sucre_record_bytes(record_format) != 0, "sucre_gather_sample: unknown record format %d", record_format);
    gather_sample_kernel<<<(n_tiles + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        *target_host, views, n_views, masks, view_kept, (const long long*)row_off, (const long long*)blk_off, *band_host, pix,
        record_format, cells, blk_mask, blk_view, cell_src);
    return check_launch("gather_sample_kernel");
}

extern "C" int sucre_band_scatter_J(const float* J_band, const sucre_band* band_host, const int32_t* pix, int64_t target_pixels,
                                    const uint64_t* dst_ptrs_host, int n_dst, void* stream) {
    clear_error();
    SUCRE_REQUIRE(J_band && band_host && dst_ptrs_host, "sucre_band_scatter_J: null pointer");
    SUCRE_REQUIRE(n_dst >= 1 && n_dst <= SUCRE_MAX_PEERS && target_pixels > 0 && band_host->n_tiles > 0 && band_host->chunk_tiles > 0,
                  "sucre_band_scatter_J: bad arguments");
    SUCRE_REQUIRE((long long)band_tile(*band_host, band_host->n_tiles - 1) * kTile < target_pixels, "sucre_band_scatter_J: band outside the image");
    ScatterPtrs ptrs{};   // passed by value: no device-side table to manage
    for (int i = 0; i < n_dst; ++i) {
        SUCRE_REQUIRE(dst_ptrs_host[i] != 0, "sucre_band_scatter_J: null destination %d", i);
        ptrs.p[i] = dst_ptrs_host[i];
    }
    const long long floats_local = (long long)band_host->n_tiles * kTile * 3;
    scatter_J_kernel<<<(unsigned)((floats_local + 255) / 256), 256, 0, (cudaStream_t)stream>>>(J_band, *band_host, pix, floats_local,
                                                                                             target_pixels * 3, ptrs, n_dst);
    return check_launch("scatter_J_kernel");
}
