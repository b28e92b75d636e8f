/* r3dp_oracle.c — plain-C (scalar, single-threaded) restatement of the Real3D-Portrait volumetric render path.
 *
 * TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into oracle/_build/libr3dp_oracle.so and loaded (ctypes) by tests/ to cross-check
 * oracle/real3d_oracle.py, the reference fixtures and the CUDA path.  Never linked into or called by the product.
 *
 * Written from SURVEY.md Appendix A (A.1-A.7), independent of torch: grid_sample, softplus, cumprod, searchsorted, sort are all restated
 * here.  Every function cites the reference file:line it follows (paths relative to the reference tree).  All arithmetic is fp32. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define C_FEAT 32
#define C_HID 64
#define C_OUT 33

/* A.1 RaySampler.forward — modules/eg3ds/volumetric_rendering/ray_sampler.py:24-63.  c2w[N][16], K[N][9] -> o,d [N][res*res][3] */
void orc_gen_rays(const float* c2w, const float* K, int N, int res, float* ro, float* rd) {
    for (int n = 0; n < N; ++n) {
        const float* m = c2w + n * 16; const float* k = K + n * 9;
        const float fx = k[0], sk = k[1], cx = k[2], fy = k[4], cy = k[5];
        for (int i = 0; i < res; ++i)
            for (int j = 0; j < res; ++j) {
                const float xc = (float)j * (1.0f / res) + 0.5f / res, yc = (float)i * (1.0f / res) + 0.5f / res;   /* :43 */
                const float xl = (xc - cx + cy * sk / fy - sk * yc / fy) / fx, yl = (yc - cy) / fy;                /* :51-52 */
                float w[3];
                for (int a = 0; a < 3; ++a) w[a] = m[a * 4 + 0] * xl + m[a * 4 + 1] * yl + m[a * 4 + 2] + m[a * 4 + 3];
                float dx = w[0] - m[3], dy = w[1] - m[7], dz = w[2] - m[11];
                float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
                if (nrm < 1e-12f) nrm = 1e-12f;                                                                      /* F.normalize */
                float* o = ro + ((size_t)n * res * res + (size_t)i * res + j) * 3; float* d = rd + ((size_t)n * res * res + (size_t)i * res + j) * 3;
                o[0] = m[3]; o[1] = m[7]; o[2] = m[11]; d[0] = dx / nrm; d[1] = dy / nrm; d[2] = dz / nrm;
            }
    }
}

/* A.2 get_ray_limits_box — math_utils.py:46-98 */
static void ray_box(const float* o, const float* d, float box, float* t0, float* t1) {
    const float lo = -box / 2, hi = box / 2;
    float inv[3], tn[3], tf[3];
    for (int a = 0; a < 3; ++a) { inv[a] = 1.0f / d[a]; tn[a] = ((inv[a] < 0 ? hi : lo) - o[a]) * inv[a]; tf[a] = ((inv[a] < 0 ? lo : hi) - o[a]) * inv[a]; }
    int valid = 1;
    float tmin = tn[0], tmax = tf[0];
    if (tmin > tf[1] || tn[1] > tmax) valid = 0;
    tmin = (tmin != tmin || tn[1] != tn[1]) ? NAN : fmaxf(tmin, tn[1]); tmax = (tmax != tmax || tf[1] != tf[1]) ? NAN : fminf(tmax, tf[1]);
    if (tmin > tf[2] || tn[2] > tmax) valid = 0;
    tmin = (tmin != tmin || tn[2] != tn[2]) ? NAN : fmaxf(tmin, tn[2]); tmax = (tmax != tmax || tf[2] != tf[2]) ? NAN : fminf(tmax, tf[2]);
    *t0 = valid ? tmin : -1.0f; *t1 = valid ? tmax : -2.0f;
}

/* A.4 sample_from_planes for ONE point — renderer.py:49-75 (grid_sample bilinear, zeros, align_corners=False); planes[3][C][H][W]; out += */
static float texel(const float* plane, int c, int H, int W, int y, int x) { return (x >= 0 && x < W && y >= 0 && y < H) ? plane[((size_t)c * H + y) * W + x] : 0.0f; }
static void sample_point(const float* planes, int H, int W, float box_warp, const float* xyz, float* feat /*[3][C]*/) {
    static const int UV[3][2] = {{0, 1}, {0, 2}, {2, 0}};                        /* plane0 (x,y), plane1 (x,z), plane2 (z,x): renderer.py:30-63 */
    for (int p = 0; p < 3; ++p) {
        const float gu = (2.0f / box_warp) * xyz[UV[p][0]], gv = (2.0f / box_warp) * xyz[UV[p][1]];
        const float px = ((gu + 1.0f) * W - 1.0f) / 2.0f, py = ((gv + 1.0f) * H - 1.0f) / 2.0f;
        const float fx0 = floorf(px), fy0 = floorf(py);
        const float wx1 = px - fx0, wy1 = py - fy0, wx0 = (fx0 + 1.0f) - px, wy0 = (fy0 + 1.0f) - py;
        const int inr = (fabsf(px) < 1e9f) && (fabsf(py) < 1e9f);
        const int x0 = inr ? (int)fx0 : -10, y0 = inr ? (int)fy0 : -10;
        const float* pl = planes + (size_t)p * C_FEAT * H * W;
        for (int c = 0; c < C_FEAT; ++c)
            feat[p * C_FEAT + c] = texel(pl, c, H, W, y0, x0) * (wx0 * wy0) + texel(pl, c, H, W, y0, x0 + 1) * (wx1 * wy0) +
                                   texel(pl, c, H, W, y0 + 1, x0) * (wx0 * wy1) + texel(pl, c, H, W, y0 + 1, x0 + 1) * (wx1 * wy1);
    }
}
void orc_sample_planes(const float* planes, int N, int H, int W, const float* coords, int P, float box_warp, float* out /*[N][3][P][C]*/) {
    float f[3 * C_FEAT];
    for (int n = 0; n < N; ++n)
        for (int s = 0; s < P; ++s) {
            sample_point(planes + (size_t)n * 3 * C_FEAT * H * W, H, W, box_warp, coords + ((size_t)n * P + s) * 3, f);
            for (int p = 0; p < 3; ++p) memcpy(out + (((size_t)n * 3 + p) * P + s) * C_FEAT, f + p * C_FEAT, sizeof(float) * C_FEAT);
        }
}

/* A.5 OSGDecoder.forward — modules/img2plane/triplane.py:133-146, networks_stylegan2.py:113-127 */
static float softplusf(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
static void decode(const float* f3 /*[3][C]*/, const float* w1, const float* b1, const float* w2, const float* b2, float* out /*[33]: sigma, rgb[32]*/) {
    float x[C_FEAT], h[C_HID];
    for (int c = 0; c < C_FEAT; ++c) x[c] = (f3[c] + f3[C_FEAT + c] + f3[2 * C_FEAT + c]) / 3.0f;       /* mean over planes :136 */
    const float g1 = 1.0f / sqrtf((float)C_FEAT), g2 = 1.0f / sqrtf((float)C_HID);
    for (int j = 0; j < C_HID; ++j) { float a = b1[j]; for (int c = 0; c < C_FEAT; ++c) a += x[c] * (w1[j * C_FEAT + c] * g1); h[j] = softplusf(a); }
    for (int o = 0; o < C_OUT; ++o) {
        float a = b2[o];
        for (int j = 0; j < C_HID; ++j) a += h[j] * (w2[o * C_HID + j] * g2);
        out[o] = o == 0 ? a : (1.0f / (1.0f + expf(-a))) * 1.002f - 0.001f;                                 /* :144 */
    }
}

/* A.6 MipRayMarcher2.run_forward for one ray — ray_marcher.py:26-57.  dep[S], val[S][33] -> rgb[32], wsum, depth (unclamped), w[S-1] */
static void march(const float* dep, const float* val, int S, int white_back, float* rgb, float* wsum, float* depth, float* w) {
    float T = 1.0f, ws = 0.0f, ds = 0.0f;
    for (int c = 0; c < C_OUT - 1; ++c) rgb[c] = 0.0f;
    for (int k = 0; k + 1 < S; ++k) {
        const float delta = dep[k + 1] - dep[k];
        const float smid = softplusf((val[k * C_OUT] + val[(k + 1) * C_OUT]) / 2.0f - 1.0f);              /* :33 */
        const float alpha = 1.0f - expf(-(smid * delta));
        const float wk = alpha * T;
        T *= (1.0f - alpha + 1e-10f);
        for (int c = 0; c < C_OUT - 1; ++c) rgb[c] += wk * ((val[k * C_OUT + 1 + c] + val[(k + 1) * C_OUT + 1 + c]) / 2.0f);
        ws += wk; ds += wk * ((dep[k] + dep[k + 1]) / 2.0f);
        if (w) w[k] = wk;
    }
    for (int c = 0; c < C_OUT - 1; ++c) { if (white_back) rgb[c] = rgb[c] + 1.0f - ws; rgb[c] = rgb[c] * 2.0f - 1.0f; }
    *wsum = ws; *depth = ds / ws;
}

/* A.7 sample_importance + sample_pdf for one ray — renderer.py:234-297 */
static void importance(const float* z, const float* w, int S, const float* u, int Ni, float* zf) {
    float a[512], cdf[512];
    float total = 0.0f;
    for (int i = 0; i < S - 1; ++i) {                       /* max_pool1d(2,1,pad 1) then avg_pool1d(2,1), + 0.01 */
        const float m0 = fmaxf(i - 1 >= 0 ? w[i - 1] : -INFINITY, w[i]);
        const float m1 = fmaxf(w[i], i + 1 < S - 1 ? w[i + 1] : -INFINITY);
        a[i] = 0.5f * (m0 + m1) + 0.01f;
    }
    const int nb = S - 3;                                   /* weights[:, 1:-1] */
    for (int i = 0; i < nb; ++i) total += a[i + 1] + 1e-5f;
    cdf[0] = 0.0f;
    for (int i = 0; i < nb; ++i) cdf[i + 1] = cdf[i] + (a[i + 1] + 1e-5f) / total;
    for (int j = 0; j < Ni; ++j) {
        int idx = 0;
        while (idx < nb + 1 && cdf[idx] <= u[j]) ++idx;    /* searchsorted(right=True) */
        const int lo = idx - 1 < 0 ? 0 : idx - 1, hi = idx > nb ? nb : idx;
        float den = cdf[hi] - cdf[lo];
        if (den < 1e-5f) den = 1.0f;
        const float blo = 0.5f * (z[lo] + z[lo + 1]), bhi = 0.5f * (z[hi] + z[hi + 1]);
        zf[j] = blo + (u[j] - cdf[lo]) / den * (bhi - blo);
    }
}

/* ImportanceRenderer.forward with 'auto' limits — renderer.py:118-167.  planes [N][3][32][H][W]; rays [N][M][3]; u_coarse [N][M][S]; u_fine [N*M][Ni]
 * mlp = {w1[64*32], b1[64], w2[33*64], b2[33]} raw state_dict tensors.  Outputs rgb [N][M][32], depth [N][M], wsum [N][M], valid [N][M] (0/1). */
void orc_render(const float* planes, int N, int H, int W, const float* ro, const float* rd, int M, int S, int Ni, float box_warp, int white_back,
                const float* u_coarse, const float* u_fine, const float* w1, const float* b1, const float* w2, const float* b2,
                float* rgb, float* depth, float* wsum, unsigned char* valid) {
    const int ST = S + Ni;
    float* t0 = (float*)malloc(sizeof(float) * N * M); float* t1 = (float*)malloc(sizeof(float) * N * M);
    float smin = INFINITY, smax = -INFINITY; int any = 0;
    for (int i = 0; i < N * M; ++i) {
        ray_box(ro + (size_t)i * 3, rd + (size_t)i * 3, box_warp, &t0[i], &t1[i]);
        valid[i] = t1[i] > t0[i];
        if (valid[i]) { any = 1; smin = fminf(smin, t0[i]); smax = fmaxf(smax, t0[i]); }
    }
    if (any) for (int i = 0; i < N * M; ++i) if (!valid[i]) { t0[i] = smin; t1[i] = smax; }                 /* :125-126 (far end from ray_start, sic) */
    float* dep = (float*)malloc(sizeof(float) * ST); float* val = (float*)malloc(sizeof(float) * ST * C_OUT);
    float* sd = (float*)malloc(sizeof(float) * ST); float* sv = (float*)malloc(sizeof(float) * ST * C_OUT);
    float* w = (float*)malloc(sizeof(float) * ST); int* ord = (int*)malloc(sizeof(int) * ST);
    float dmin = INFINITY, dmax = -INFINITY, f[3 * C_FEAT], xyz[3];
    for (int n = 0; n < N; ++n)
        for (int m = 0; m < M; ++m) {
            const size_t r = (size_t)n * M + m;
            const float* o = ro + r * 3; const float* d = rd + r * 3;
            const float* pl = planes + (size_t)n * 3 * C_FEAT * H * W;
            for (int k = 0; k < S; ++k) {                                                                   /* A.3: renderer.py:223-226 */
                dep[k] = t0[r] + ((float)k / (float)(S - 1)) * (t1[r] - t0[r]) + u_coarse[r * S + k] * ((t1[r] - t0[r]) / (float)(S - 1));
                for (int a = 0; a < 3; ++a) xyz[a] = o[a] + dep[k] * d[a];
                sample_point(pl, H, W, box_warp, xyz, f);
                decode(f, w1, b1, w2, b2, val + k * C_OUT);
            }
            int cnt = S;
            const float* fd = dep; const float* fv = val;
            if (Ni > 0) {
                float ws_, dp_, tmp[C_OUT];
                march(dep, val, S, white_back, tmp, &ws_, &dp_, w);
                importance(dep, w, S, u_fine + r * Ni, Ni, dep + S);
                for (int j = 0; j < Ni; ++j) {
                    for (int a = 0; a < 3; ++a) xyz[a] = o[a] + dep[S + j] * d[a];
                    sample_point(pl, H, W, box_warp, xyz, f);
                    decode(f, w1, b1, w2, b2, val + (S + j) * C_OUT);
                }
                for (int i = 0; i < ST; ++i) {                                                              /* stable sort by depth: renderer.py:197-207 */
                    int rank = 0;
                    for (int j = 0; j < ST; ++j) rank += (dep[j] < dep[i]) || (dep[j] == dep[i] && j < i);
                    ord[rank] = i;
                }
                for (int i = 0; i < ST; ++i) { sd[i] = dep[ord[i]]; memcpy(sv + i * C_OUT, val + ord[i] * C_OUT, sizeof(float) * C_OUT); }
                cnt = ST; fd = sd; fv = sv;
            }
            for (int k = 0; k < cnt; ++k) { dmin = fminf(dmin, fd[k]); dmax = fmaxf(dmax, fd[k]); }
            march(fd, fv, cnt, white_back, rgb + r * (C_OUT - 1), &wsum[r], &depth[r], NULL);
        }
    for (int i = 0; i < N * M; ++i) {                                                                       /* ray_marcher.py:49-50 */
        float dd = depth[i];
        if (dd != dd) dd = INFINITY;
        depth[i] = fminf(fmaxf(dd, dmin), dmax);
    }
    free(t0); free(t1); free(dep); free(val); free(sd); free(sv); free(w); free(ord);
}
