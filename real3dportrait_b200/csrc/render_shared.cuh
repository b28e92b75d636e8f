// Structures shared by the two fused-render translation units:
//   render.cu         entry points, limits / clamp kernels, the CTA-per-ray-tile kernel (two-pass importance renders, odd shapes)
//   render_stream.cu  the warp-specialised streaming kernel (single-pass renders: gather || tcgen05 decode || march overlap)
#pragma once
#include "render_core.cuh"
#include "tc_prims.cuh"

namespace r3dp {

// ---- tensor-core decoder operand image --------------------------------------------------------------------------------------------
// The OSG decoder is two GEMMs over the samples: [M x 32] x [32 x 64] -> softplus -> [M x 64] x [64 x 33].  They run on tcgen05 with
// fp16 operands and fp32 accumulation in TMEM; fp32 accuracy is kept by splitting every operand into two fp16 halves
// (v = hi + lo exactly to 2^-22 |v|) and summing the three significant partial products hi*hi + lo*hi + hi*lo (lo*lo ~ 2^-22 is dropped).
// Operand images (K-major, 128-byte swizzle: one 128 B row = 64 fp16, 8-row groups 1024 B apart):
//   A1 tile  128 samples x [x_hi(32) | x_lo(32)]             written by the gather
//   W1       64 hidden   x [w_hi(32) | w_lo(32)]             k-steps 0,1 = hi, 2,3 = lo
//   A2 tile  128 samples x [h_hi(64)] , [h_lo(64)]           two atoms, written by the layer-1 epilogue
//   W2       48 outputs  x [w_hi(64)] , [w_lo(64)]           two atoms, rows >= 33 are zero (UMMA N must be a multiple of 16)
struct alignas(16) MlpTcImage {
    uint8_t w1[kHidden * 128];
    uint8_t w2hi[48 * 128];
    uint8_t w2lo[48 * 128];
    float b1[kHidden];
    float b2[48];
};
static_assert(sizeof(MlpTcImage) == 8192 + 6144 + 6144 + 256 + 192, "MlpTcImage layout");

constexpr uint32_t kIdescL1 = (1u << 4) | ((uint32_t)(kHidden >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);    // M128 N64, f16 x f16 -> f32
constexpr uint32_t kIdescL2 = (1u << 4) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);         // M128 N48

__device__ __forceinline__ uint32_t sw128_off(int row, int k) {                // byte offset of fp16 element (row, k) of a swizzled atom
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7))) << 4) + (k & 7) * 2);
}

// ---- per-call workspace ---------------------------------------------------------------------------------------------------------------
// [RenderWs | MlpTcImage | MlpConst-free padding | limits float2[N*M]]: the decoder operand image lives in the CALLER's workspace, so two
// renders with different decoders on different streams never share state.
struct RenderWs {
    unsigned t0_min, t0_max;   // ordered-uint encoded floats over valid rays
    unsigned d_min, d_max;     // over every sample depth of the call
    unsigned n_valid;
    unsigned pad[3];
};
static_assert(sizeof(RenderWs) == 32, "ws header");
constexpr size_t kWsImageOff = 256;                                              // MlpTcImage (16-byte aligned)
constexpr size_t kWsLimitsOff = kWsImageOff + ((sizeof(MlpTcImage) + 255) / 256) * 256;

// channels-last plane addressing, strides in floats: texel (plane p, row y, col x) of frame n starts at
//   base + n*frame_stride + p*plane_stride + y*row_stride + x*texel_stride    and holds kC contiguous floats.
// [N,3,H,W,C]: plane = H*W*C, row = W*C, texel = C.   [N,H,W,3,C] (= torch channels_last of the producer's [N,3*C,H,W]): plane = C, row = W*3*C, texel = 3*C.
struct PlaneSet {
    const float* base;
    long long frame_stride;
    int plane_stride, row_stride, texel_stride;
};

struct RenderArgs {
    PlaneSet p0, p1;                    // p1.base == nullptr: single set; else both sets are sampled at the same points and added
    int N, H, W;
    const float* ray_o; const float* ray_d; const float* camera; int M, res;
    int S, S_imp; float box_warp; int white_back;
    const float* u_coarse; const float* u_fine;
    r3dp_mlp_t mlp;
    const MlpTcImage* image;            // tcgen05 decoder operands (workspace)
    float* rgb; float* depth; float* wsum;
    const float2* limits; const uint8_t* valid; RenderWs* ws;
    int tiles_per_frame, tile_cols;     // ray tiling of the CTA-per-tile kernel (see ray_of)
};

__device__ __forceinline__ Ray fetch_ray(const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                                         const float* __restrict__ camera, int res, int n, int M, int m) {
    if (ray_o != nullptr) {
        const float* o = ray_o + ((size_t)n * M + m) * 3; const float* d = ray_d + ((size_t)n * M + m) * 3;
        Ray r; r.ox = o[0]; r.oy = o[1]; r.oz = o[2]; r.dx = d[0]; r.dy = d[1]; r.dz = d[2];
        return r;
    }
    return make_ray(camera + n * 25, camera + n * 25 + 16, res, m);
}

// Descriptor of one bilinear tap quad with explicit strides: out[0] = float-punned offset (in floats, from the frame's base) of the
// 2x2 texel block (shifted inside the plane; outside taps get weight 0, see axis_taps), out[1..4] = weights.
__device__ __forceinline__ void tap_desc_s(float gu, float gv, int H, int W, int plane_off, int row_stride, int texel_stride, float* out) {
    const float px = ((gu + 1.0f) * (float)W - 1.0f) * 0.5f;      // align_corners=False
    const float py = ((gv + 1.0f) * (float)H - 1.0f) * 0.5f;
    int bx, by; float wxa, wxb, wya, wyb;
    axis_taps(px, W, bx, wxa, wxb);
    axis_taps(py, H, by, wya, wyb);
    out[0] = __int_as_float(plane_off + by * row_stride + bx * texel_stride);
    out[1] = wxa * wya; out[2] = wxb * wya; out[3] = wxa * wyb; out[4] = wxb * wyb;
}

// order-preserving key of a depth for the merge sort: NaN last (as torch.sort), -0 == +0
__device__ __forceinline__ unsigned sort_key(float d) {
    if (d != d) return 0xffffffffu;
    return f2ord(d == 0.f ? 0.f : d);
}

// render_stream.cu
extern int g_rs_chunk_log2;      // A/B knob (r3dp_set_option / R3DP_RS_D)
bool render_stream_fits(const RenderArgs& a);
int launch_render_stream(const RenderArgs& a, cudaStream_t st);

}  // namespace r3dp
