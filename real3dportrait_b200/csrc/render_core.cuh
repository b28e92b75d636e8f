// Device building blocks shared by the fused renderer (render.cu) and the stand-alone sample ops (sample.cu):
// ray generation, box limits, channels-last tri-plane gather, the OSG decoder MLP.
#pragma once
#include "common.cuh"

namespace r3dp {

constexpr int kC = 32;        // tri-plane feature channels == decoder inputs
constexpr int kHidden = 64;   // OSGDecoder hidden width
constexpr int kOut = 33;      // 1 density + 32 colour features
constexpr int kRow = 33;      // smem row stride (floats) per sample: odd => conflict-free thread-per-row access
constexpr int kW2Pad = 36;    // W2^T rows padded 33 -> 36 floats (9 x float4)

// ---- decoder weights staged in shared memory ------------------------------------------------------------------
// w1c : [8 chunks][32 c][8 hidden]  (W1 * 1/sqrt(32)), so one hidden-chunk/channel is two broadcast LDS.128
// w2t : [64 hidden][36]             (W2 * 1/sqrt(64)) transposed, row j = weights of hidden j to the 33 outputs
struct MlpSmem {
    float w1c[8 * kC * 8];
    float b1[kHidden];
    float w2t[kHidden * kW2Pad];
    float b2[kW2Pad];
};

__device__ __forceinline__ void load_mlp_smem(MlpSmem& s, const r3dp_mlp_t& m, int tid, int nthreads) {
    const float g1 = 0.17677669529663687f;  // 1/sqrt(32)  (FullyConnectedLayer.weight_gain, networks_stylegan2.py:113)
    const float g2 = 0.125f;                // 1/sqrt(64)
    for (int i = tid; i < kHidden * kC; i += nthreads) {
        int j = i / kC, c = i % kC;
        s.w1c[((j >> 3) * kC + c) * 8 + (j & 7)] = __ldg(m.w1 + i) * g1;
    }
    for (int i = tid; i < kHidden * kW2Pad; i += nthreads) {
        int j = i / kW2Pad, o = i % kW2Pad;
        s.w2t[i] = (o < kOut) ? __ldg(m.w2 + o * kHidden + j) * g2 : 0.0f;
    }
    for (int i = tid; i < kHidden; i += nthreads) s.b1[i] = __ldg(m.b1 + i);
    for (int i = tid; i < kW2Pad; i += nthreads) s.b2[i] = (i < kOut) ? __ldg(m.b2 + i) : 0.0f;
}

// ---- rays (ray_sampler.py:43-61) -------------------------------------------------------------------------------
struct Ray {
    float ox, oy, oz, dx, dy, dz;
};

__device__ __forceinline__ Ray make_ray(const float* __restrict__ c2w, const float* __restrict__ K, int res, int m) {
    const float fx = K[0], sk = K[1], cx = K[2], fy = K[4], cy = K[5];
    const int i = m / res, j = m - i * res;
    const float inv = 1.0f / (float)res, half = 0.5f / (float)res;
    const float x_cam = __fadd_rn(__fmul_rn((float)j, inv), half);
    const float y_cam = __fadd_rn(__fmul_rn((float)i, inv), half);
    float xl = __fadd_rn(__fsub_rn(x_cam, cx), __fdiv_rn(__fmul_rn(cy, sk), fy));
    xl = __fdiv_rn(__fsub_rn(xl, __fdiv_rn(__fmul_rn(sk, y_cam), fy)), fx);
    const float yl = __fdiv_rn(__fsub_rn(y_cam, cy), fy);
    Ray r;
    r.ox = c2w[3]; r.oy = c2w[7]; r.oz = c2w[11];
    // world = c2w * (xl, yl, 1, 1)
    float wx = c2w[0] * xl + c2w[1] * yl + c2w[2] + c2w[3];
    float wy = c2w[4] * xl + c2w[5] * yl + c2w[6] + c2w[7];
    float wz = c2w[8] * xl + c2w[9] * yl + c2w[10] + c2w[11];
    float dx = wx - r.ox, dy = wy - r.oy, dz = wz - r.oz;
    float nrm = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);   // F.normalize eps
    r.dx = __fdiv_rn(dx, nrm); r.dy = __fdiv_rn(dy, nrm); r.dz = __fdiv_rn(dz, nrm);
    return r;
}

// ---- box limits (math_utils.py:46-98) --------------------------------------------------------------------------
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }

__device__ __forceinline__ void ray_box(const Ray& r, float box, float& t0, float& t1) {
    const float lo = -0.5f * box, hi = 0.5f * box;
    const float ix = __fdiv_rn(1.0f, r.dx), iy = __fdiv_rn(1.0f, r.dy), iz = __fdiv_rn(1.0f, r.dz);
    bool valid = true;
    float tmin = __fmul_rn((ix < 0 ? hi : lo) - r.ox, ix), tmax = __fmul_rn((ix < 0 ? lo : hi) - r.ox, ix);
    float tymin = __fmul_rn((iy < 0 ? hi : lo) - r.oy, iy), tymax = __fmul_rn((iy < 0 ? lo : hi) - r.oy, iy);
    if (tmin > tymax || tymin > tmax) valid = false;
    tmin = max_nan(tmin, tymin); tmax = min_nan(tmax, tymax);
    float tzmin = __fmul_rn((iz < 0 ? hi : lo) - r.oz, iz), tzmax = __fmul_rn((iz < 0 ? lo : hi) - r.oz, iz);
    if (tmin > tzmax || tzmin > tmax) valid = false;
    tmin = max_nan(tmin, tzmin); tmax = min_nan(tmax, tzmax);
    t0 = valid ? tmin : -1.0f;
    t1 = valid ? tmax : -2.0f;
}

// ---- tri-plane gather (renderer.py:49-75), channels-last planes ------------------------------------------------
// One call = one sample point handled by an 8-lane group; lane `cq` (0..7) owns channels 4cq..4cq+3.
// Returns the SUM over the three planes (caller divides by 3 for the decoder's mean) or, through `per_plane`,
// the three per-plane results.
struct PlaneView {
    const float* base;   // planes_cl + n*3*H*W*C
    int H, W;
    float scale;         // 2 / box_warp
};

__device__ __forceinline__ float4 bilinear_cl(const float* __restrict__ plane, int H, int W, float gu, float gv, int cq) {
    // F.grid_sample(bilinear, zeros, align_corners=False): pixel = ((g+1)*size-1)/2
    const float px = ((gu + 1.0f) * (float)W - 1.0f) * 0.5f;
    const float py = ((gv + 1.0f) * (float)H - 1.0f) * 0.5f;
    const float fx0 = floorf(px), fy0 = floorf(py);
    const float wx1 = px - fx0, wy1 = py - fy0;
    const float wx0 = (fx0 + 1.0f) - px, wy0 = (fy0 + 1.0f) - py;
    // clamp the float before the int cast so far-out-of-range / non-finite coords cannot overflow
    const int x0 = (int)fminf(fmaxf(fx0, -2.0f), (float)W + 1.0f), y0 = (int)fminf(fmaxf(fy0, -2.0f), (float)H + 1.0f);
    const bool okx0 = (unsigned)x0 < (unsigned)W, okx1 = (unsigned)(x0 + 1) < (unsigned)W;
    const bool oky0 = (unsigned)y0 < (unsigned)H, oky1 = (unsigned)(y0 + 1) < (unsigned)H;
    const bool finite = (fabsf(px) <= 3.0e38f) && (fabsf(py) <= 3.0e38f);    // false for NaN and +-inf
    const float* p00 = plane + ((size_t)y0 * W + x0) * kC + cq * 4;
    float4 a = make_float4(0, 0, 0, 0), b = a, c = a, d = a;
    if (finite && oky0 && okx0) a = ldg_nc_f4(p00);
    if (finite && oky0 && okx1) b = ldg_nc_f4(p00 + kC);
    if (finite && oky1 && okx0) c = ldg_nc_f4(p00 + (size_t)W * kC);
    if (finite && oky1 && okx1) d = ldg_nc_f4(p00 + (size_t)W * kC + kC);
    // non-finite coordinates contribute nothing (their weights would be NaN)
    const float w00 = finite ? wx0 * wy0 : 0.f, w10 = finite ? wx1 * wy0 : 0.f, w01 = finite ? wx0 * wy1 : 0.f, w11 = finite ? wx1 * wy1 : 0.f;
    float4 r;
    r.x = a.x * w00 + b.x * w10 + c.x * w01 + d.x * w11;
    r.y = a.y * w00 + b.y * w10 + c.y * w01 + d.y * w11;
    r.z = a.z * w00 + b.z * w10 + c.z * w01 + d.z * w11;
    r.w = a.w * w00 + b.w * w10 + c.w * w01 + d.w * w11;
    return r;
}

// Descriptor of one bilinear tap quad for the fused renderer: out[0] = float-punned offset (in floats, from the frame's plane base)
// of a 2x2 texel block that lies fully inside the plane, out[1..4] = weights of its (x,y), (x+1,y), (x,y+1), (x+1,y+1) texels.
// grid_sample's zero padding is folded into the weights: a tap outside the plane gets weight 0 and the block is shifted inside,
// so the gather needs no bounds checks (requires H, W >= 2).  The products are the same ones grid_sample forms.
__device__ __forceinline__ void axis_taps(float p, int size, int& base, float& wa, float& wb) {
    const float f0 = floorf(p);
    const float w1 = p - f0, w0 = (f0 + 1.0f) - p;
    const int i0 = (int)fminf(fmaxf(f0, -2.0f), (float)size + 1.0f);
    base = min(max(i0, 0), size - 2);
    const bool finite = (p == p);
    wa = 0.f; wb = 0.f;
    if (finite) {
        if (i0 == base) { wa = w0; wb = w1; }                 // both taps inside
        else if (i0 == -1) { wa = w1; }                       // left tap outside: the right tap is texel 0 = base
        else if (i0 == size - 1) { wb = w0; }                 // right tap outside: the left tap is texel size-1 = base+1
    }
}
__device__ __forceinline__ void tap_desc(float gu, float gv, int H, int W, int plane, float* out) {
    const float px = ((gu + 1.0f) * (float)W - 1.0f) * 0.5f;      // align_corners=False
    const float py = ((gv + 1.0f) * (float)H - 1.0f) * 0.5f;
    int bx, by; float wxa, wxb, wya, wyb;
    axis_taps(px, W, bx, wxa, wxb);
    axis_taps(py, H, by, wya, wyb);
    out[0] = __int_as_float(((plane * H + by) * W + bx) * kC);
    out[1] = wxa * wya; out[2] = wxb * wya; out[3] = wxa * wyb; out[4] = wxb * wyb;
}

// plane 0 <- (x,y), plane 1 <- (x,z), plane 2 <- (z,x)   (generate_planes + project_onto_planes, renderer.py:30-63)
__device__ __forceinline__ void gather3(const PlaneView& pv, float x, float y, float z, int cq, float4& f0, float4& f1, float4& f2) {
    const float gx = pv.scale * x, gy = pv.scale * y, gz = pv.scale * z;
    const size_t psz = (size_t)pv.H * pv.W * kC;
    f0 = bilinear_cl(pv.base, pv.H, pv.W, gx, gy, cq);
    f1 = bilinear_cl(pv.base + psz, pv.H, pv.W, gx, gz, cq);
    f2 = bilinear_cl(pv.base + 2 * psz, pv.H, pv.W, gz, gx, cq);
}

// ---- OSG decoder (triplane.py:133-146) -------------------------------------------------------------------------
// Each thread decodes TWO samples whose 32 mean features sit in smem rows `ra`, `rb` (stride kRow); the 33 outputs
// (density, then 32 colours after the sigmoid clamp) overwrite the same rows.  Weight reads are warp-broadcast
// LDS.128 shared by both samples.  `has_b` = second sample exists.
__device__ __forceinline__ void decode_pair(const MlpSmem& w, float* __restrict__ ra, float* __restrict__ rb, bool has_b) {
    float xa[kC], xb[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) { xa[c] = ra[c]; xb[c] = has_b ? rb[c] : 0.0f; }
    float ya[kW2Pad], yb[kW2Pad];
#pragma unroll
    for (int o = 0; o < kW2Pad; ++o) { ya[o] = w.b2[o]; yb[o] = w.b2[o]; }

#pragma unroll 1
    for (int ch = 0; ch < 8; ++ch) {
        float ha[8], hb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { ha[q] = w.b1[ch * 8 + q]; hb[q] = ha[q]; }
        const float4* w1 = reinterpret_cast<const float4*>(w.w1c + ch * kC * 8);
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            const float4 u = w1[2 * c], v = w1[2 * c + 1];
            ha[0] = fmaf(xa[c], u.x, ha[0]); hb[0] = fmaf(xb[c], u.x, hb[0]);
            ha[1] = fmaf(xa[c], u.y, ha[1]); hb[1] = fmaf(xb[c], u.y, hb[1]);
            ha[2] = fmaf(xa[c], u.z, ha[2]); hb[2] = fmaf(xb[c], u.z, hb[2]);
            ha[3] = fmaf(xa[c], u.w, ha[3]); hb[3] = fmaf(xb[c], u.w, hb[3]);
            ha[4] = fmaf(xa[c], v.x, ha[4]); hb[4] = fmaf(xb[c], v.x, hb[4]);
            ha[5] = fmaf(xa[c], v.y, ha[5]); hb[5] = fmaf(xb[c], v.y, hb[5]);
            ha[6] = fmaf(xa[c], v.z, ha[6]); hb[6] = fmaf(xb[c], v.z, hb[6]);
            ha[7] = fmaf(xa[c], v.w, ha[7]); hb[7] = fmaf(xb[c], v.w, hb[7]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float sa = softplus_fast(ha[q]), sb = softplus_fast(hb[q]);
            const float4* w2 = reinterpret_cast<const float4*>(w.w2t + (ch * 8 + q) * kW2Pad);
#pragma unroll
            for (int o4 = 0; o4 < kW2Pad / 4; ++o4) {
                const float4 t = w2[o4];
                ya[4 * o4 + 0] = fmaf(sa, t.x, ya[4 * o4 + 0]); yb[4 * o4 + 0] = fmaf(sb, t.x, yb[4 * o4 + 0]);
                ya[4 * o4 + 1] = fmaf(sa, t.y, ya[4 * o4 + 1]); yb[4 * o4 + 1] = fmaf(sb, t.y, yb[4 * o4 + 1]);
                ya[4 * o4 + 2] = fmaf(sa, t.z, ya[4 * o4 + 2]); yb[4 * o4 + 2] = fmaf(sb, t.z, yb[4 * o4 + 2]);
                ya[4 * o4 + 3] = fmaf(sa, t.w, ya[4 * o4 + 3]); yb[4 * o4 + 3] = fmaf(sb, t.w, yb[4 * o4 + 3]);
            }
        }
    }
    ra[0] = ya[0];
    if (has_b) rb[0] = yb[0];
#pragma unroll
    for (int o = 1; o < kOut; ++o) {
        ra[o] = sigmoid_fast(ya[o]) * 1.002f - 0.001f;
        if (has_b) rb[o] = sigmoid_fast(yb[o]) * 1.002f - 0.001f;
    }
}


}  // namespace r3dp
