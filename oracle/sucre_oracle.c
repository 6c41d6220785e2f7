/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of SUCRe's hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (sucre_b200/) never does.  Parity is PINNED: tests/test_oracle_golden.py checks every
 * function below bit-for-bit (indices, d, cP, z, I) or to 1e-5 (fit trajectories) against outputs of the
 * unmodified reference captured by oracle/gen_golden.py (tests/golden/ *.npz).
 *
 * Each function restates, per target pixel / per observation, what the reference computes with whole-tensor
 * ATen ops.  fp32 roundings follow the probe-validated contract of SURVEY.md §8a': every `mm(3x3, 3xn)` is
 * fma(a2,x2, fma(a1,x1, a0*x0)) per row, the translation is a separately rounded add, divisions are IEEE.
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -shared -fPIC (see oracle/Makefile) — never -ffast-math.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Per-view constants, produced on the host with the reference's own torch expressions:
 * K (sfm.py:204-208), Kinv = K.inverse() (sfm.py:92), R,t = cam->world pose (sfm.py:219-222),
 * Ri = R.T, ti = -R.T @ t (sfm.py:47).  All row-major 3x3 / 3-vectors. */
typedef struct {
    float K[9], Kinv[9], R[9], t[3], Ri[9], ti[3];
    int32_t width, height;
} oracle_view;

static inline void mat3(const float* M, const float* x, float* y) {
    for (int i = 0; i < 3; ++i)
        y[i] = fmaf(M[3 * i + 2], x[2], fmaf(M[3 * i + 1], x[1], M[3 * i] * x[0]));
}

/* sfm.py:90-93 unproject_depth: cP = Kinv @ (d * (u+.5, v+.5, 1)) ; depth = u16/1000 (loader.py:167) */
static inline void unproject(const oracle_view* V, int u, int v, float d, float* cP) {
    float X[3] = {d * ((float)u + 0.5f), d * ((float)v + 0.5f), d};
    mat3(V->Kinv, X, cP);
}

/* sfm.py:49-55 Pose.transform: R @ P + t (two roundings) */
static inline void to_world(const oracle_view* V, const float* cP, float* wP) {
    float r[3];
    mat3(V->R, cP, r);
    for (int i = 0; i < 3; ++i) wP[i] = r[i] + V->t[i];
}

/* sfm.py:103-107 project_to_view + sfm.py:116-117 truncation and bounds test.
 * Returns 1 and the integer pixel if it lands in the image of V. */
static inline int project(const oracle_view* V, const float* wP, int* u, int* v) {
    float r[3], c[3], p[3];
    mat3(V->Ri, wP, r);
    for (int i = 0; i < 3; ++i) c[i] = r[i] + V->ti[i];
    mat3(V->K, c, p);
    float px = p[0] / p[2], py = p[1] / p[2];
    /* .long() truncates toward zero; NaN/inf/huge become INT64_MIN and fail `0 <= u2` (probe).
     * trunc(x) >= 0  <=>  x > -1 : pixels projecting into (-1,0) are accepted as index 0, as in the reference. */
    if (!(px > -1.0f && px < (float)V->width && py > -1.0f && py < (float)V->height)) return 0;
    *u = (int)px;
    *v = (int)py;
    return 1;
}

/* Dense two-way match of target T against source S (sfm.py:121-125, 154-159, 171-175, restated per pixel:
 * the backward projection is evaluated only at the source pixel the forward projection lands on).
 * idx[p] = u2 | v2 << 16 for matched target pixels (row-major p = v1*W_T + u1), -1 otherwise.
 * Returns the number of matches; *n_inbounds counts forward projections that land inside S. */
int64_t oracle_match_pair(const uint16_t* depthT, const oracle_view* T, const uint16_t* depthS,
                          const oracle_view* S, int32_t* idx, int64_t* n_inbounds) {
    const int W = T->width, H = T->height;
    int64_t n = 0, nin = 0;
#pragma omp parallel for schedule(static) reduction(+ : n, nin)
    for (int v1 = 0; v1 < H; ++v1) {
        for (int u1 = 0; u1 < W; ++u1) {
            const int64_t p = (int64_t)v1 * W + u1;
            idx[p] = -1;
            const float d1 = (float)depthT[p] / 1000.0f;
            if (!(d1 > 0.0f)) continue; /* sfm.py:96 */
            float cP[3], wP[3];
            unproject(T, u1, v1, d1, cP);
            to_world(T, cP, wP);
            int u2, v2;
            if (!project(S, wP, &u2, &v2)) continue;
            ++nin;
            const float d2 = (float)depthS[(int64_t)v2 * S->width + u2] / 1000.0f;
            if (!(d2 > 0.0f)) continue; /* S pixel has no back-projection => map entry stays -1 (sfm.py:155) */
            float cP2[3], wP2[3];
            unproject(S, u2, v2, d2, cP2);
            to_world(S, cP2, wP2);
            int ub, vb;
            if (!project(T, wP2, &ub, &vb)) continue;
            if (ub != u1 || vb != v1) continue; /* sfm.py:173 */
            idx[p] = u2 | (v2 << 16);
            ++n;
        }
    }
    if (n_inbounds) *n_inbounds = nin;
    return n;
}

/* Compacts idx (row-major over T, the order torch.where yields, sfm.py:96) and samples the observation payload:
 * u1,v1,u2,v2 (int16, loader.py:71-74), d = depth_S[v2,u2] (sfm.py:137), I = rgb_S[v2,u2]/255 (loader.py:87,157),
 * cP = Kinv_S @ (d*(u2+.5,v2+.5,1)) (loader.py:113), z = ||cP|| sequential, no fma (sucre.py:53).
 * Arrays are length n (cP and I are (3,n) planar like the reference's tensors). Returns n. */
int64_t oracle_sample_pair(const int32_t* idx, int W_T, int H_T, const uint16_t* depthS, const uint8_t* rgbS,
                           const oracle_view* S, int16_t* u1, int16_t* v1, int16_t* u2, int16_t* v2, float* d,
                           float* cP, float* z, float* I, int64_t n) {
    int64_t k = 0;
    for (int64_t p = 0; p < (int64_t)W_T * H_T; ++p) {
        if (idx[p] < 0) continue;
        if (k >= n) return -1;
        const int uu = idx[p] & 0xffff, vv = idx[p] >> 16;
        const int64_t q = (int64_t)vv * S->width + uu;
        u1[k] = (int16_t)(p % W_T);
        v1[k] = (int16_t)(p / W_T);
        u2[k] = (int16_t)uu;
        v2[k] = (int16_t)vv;
        const float dd = (float)depthS[q] / 1000.0f;
        d[k] = dd;
        float c[3];
        unproject(S, uu, vv, dd, c);
        for (int i = 0; i < 3; ++i) cP[i * n + k] = c[i];
        z[k] = sqrtf(((c[0] * c[0]) + (c[1] * c[1])) + (c[2] * c[2]));
        if (rgbS)
            for (int i = 0; i < 3; ++i) I[i * n + k] = (float)rgbS[3 * q + i] / 255.0f;
        ++k;
    }
    return k;
}

/* ------------------------------------------------------------------------------------------------------
 * Fit (sucre.py:52-82 model, 124-157 adam).  Observations are given per kept view (name-sorted like the
 * HDF5 groups): view k owns obs [view_off[k], view_off[k+1]); pix = v1*W+u1 is unique within a view, so
 * the per-view scatter-adds of update_J (sucre.py:73-76) are race-free under `omp parallel for`.
 * I is (N,3) interleaved.  Per-observation arithmetic is fp32 like the reference's; the global sums (cost,
 * gradients) are accumulated in double (the reference's fp32 tree sums differ from this by ~1e-7 relative).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
    float m[9], v[9];
} adam9;

/* torch/optim/adam.py _single_tensor_adam, non-capturable CPU branch: fp32 tensors, python-float scalars. */
static inline float adam_update(float p, float g, float* m, float* v, int t, double lr) {
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    *m = *m + (float)(1.0 - b1) * (g - *m);              /* exp_avg.lerp_(grad, 1-beta1), weight < 0.5 branch */
    *v = (*v * (float)b2) + (float)(1.0 - b2) * (g * g); /* mul_(beta2).addcmul_(grad, grad, value=1-beta2) */
    const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
    const double step_size = lr / bc1, bc2_sqrt = sqrt(bc2);
    const float denom = (sqrtf(*v) / (float)bc2_sqrt) + (float)eps;
    return p + (float)(-step_size) * (*m / denom);       /* param.addcdiv_(exp_avg, denom, value=-step_size) */
}

/* mode 0: --use-closed-form (J recomputed from the pre-step parameters each iteration, sucre.py:66-77,141;
 *         once more after the loop, sucre.py:156).  J is output only.
 * mode 1: default CLI mode, J is an Adam parameter (sucre.py:47-50); J holds the initial image on entry
 *         (NaN where target depth <= 0) and the optimised image on exit.
 * params = B[3], beta[3], gamma[3] in/out.  history (num_iter x 9) = parameters after each step;
 * cost (num_iter) = sum of squared residuals before each step (sucre.py:144-146,150). */
int oracle_fit(int mode, int n_views, const int64_t* view_off, const int32_t* pix, const float* z, const float* I,
               int64_t P, float* params, float* J, int num_iter, double lr, float* history, double* cost) {
    const int64_t N = view_off[n_views];
    float* B = params;
    float* beta = params + 3;
    float* gamma = params + 6;
    adam9 st;
    memset(&st, 0, sizeof st);
    float *num = NULL, *den = NULL, *Jm = NULL, *Jv = NULL, *Jg = NULL;
    if (mode == 0) {
        num = (float*)malloc(sizeof(float) * 3 * P);
        den = (float*)malloc(sizeof(float) * 3 * P);
    } else {
        Jm = (float*)calloc(3 * P, sizeof(float));
        Jv = (float*)calloc(3 * P, sizeof(float));
        Jg = (float*)malloc(sizeof(float) * 3 * P);
    }
    for (int it = 0; it <= num_iter; ++it) {
        if (mode == 0) { /* update_J */
            memset(num, 0, sizeof(float) * 3 * P);
            memset(den, 0, sizeof(float) * 3 * P);
            for (int k = 0; k < n_views; ++k) {
#pragma omp parallel for schedule(static)
                for (int64_t o = view_off[k]; o < view_off[k + 1]; ++o) {
                    for (int c = 0; c < 3; ++c) {
                        const float a = expf(-beta[c] * z[o]);
                        const float bs = B[c] * (1.0f - expf(-gamma[c] * z[o]));
                        num[3 * (int64_t)pix[o] + c] += (I[3 * o + c] - bs) * a;
                        den[3 * (int64_t)pix[o] + c] += a * a;
                    }
                }
            }
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < 3 * P; ++i) J[i] = num[i] / den[i]; /* 0/0 = NaN for unobserved pixels */
        }
        if (it == num_iter) break;
        double g[9] = {0}, sq = 0.0;
        if (mode == 1) memset(Jg, 0, sizeof(float) * 3 * P);
        for (int k = 0; k < n_views; ++k) {
            double gB[3] = {0}, gb[3] = {0}, gg[3] = {0}, s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : gB[:3], gb[:3], gg[:3], s)
            for (int64_t o = view_off[k]; o < view_off[k + 1]; ++o) {
                for (int c = 0; c < 3; ++c) {
                    const float a = expf(-beta[c] * z[o]);
                    const float e = expf(-gamma[c] * z[o]);
                    const float Jc = J[3 * (int64_t)pix[o] + c];
                    const float r = I[3 * o + c] - (Jc * a + B[c] * (1.0f - e)); /* sucre.py:81 */
                    s += (double)r * r;
                    gB[c] += (double)(r * (1.0f - e));
                    gb[c] += (double)(r * Jc * z[o] * a);
                    gg[c] += (double)(r * B[c] * z[o] * e);
                    if (mode == 1) Jg[3 * (int64_t)pix[o] + c] += r * a;
                }
            }
            for (int c = 0; c < 3; ++c) {
                g[c] += gB[c];
                g[3 + c] += gb[c];
                g[6 + c] += gg[c];
            }
            sq += s;
        }
        cost[it] = sq;
        const double sc = 2.0 / (3.0 * (double)N); /* d/dtheta of sum r^2 / N / 3 (sucre.py:145) */
        for (int c = 0; c < 3; ++c) {
            const float dB = (float)(-sc * g[c]), db = (float)(sc * g[3 + c]), dg = (float)(-sc * g[6 + c]);
            B[c] = adam_update(B[c], dB, &st.m[c], &st.v[c], it + 1, lr);
            beta[c] = adam_update(beta[c], db, &st.m[3 + c], &st.v[3 + c], it + 1, lr);
            gamma[c] = adam_update(gamma[c], dg, &st.m[6 + c], &st.v[6 + c], it + 1, lr);
        }
        if (mode == 1) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < 3 * P; ++i)
                J[i] = adam_update(J[i], (float)(-sc) * Jg[i], &Jm[i], &Jv[i], it + 1, lr);
        }
        memcpy(history + 9 * it, params, sizeof(float) * 9);
    }
    free(num); free(den); free(Jm); free(Jv); free(Jg);
    return 0;
}

int oracle_abi_version(void) { return 1; }
