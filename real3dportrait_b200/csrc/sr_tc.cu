// Tensor-core super-resolution path for sm_100a: the four modulated 3x3 convolutions of SuperresolutionHybrid8XDC
// (197.6 GFLOP/frame) as TMA-fed tcgen05 implicit GEMMs.
//
//   activations  NHWC fp16, channels padded to a multiple of 64 (one 128-byte swizzle row = 64 channels)
//   weights      per-sample folded (modulated+demodulated) fp16, packed [n][tap][Cout][Cin_pad]  (K-major B operand)
//   accumulate   fp32 in TMEM; epilogue in fp32 (bias, lrelu*sqrt2, ToRGB) then fp16 / fp32 stores
//
// One CTA = one 128(pixels) x 128(couts) output tile: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane,
// tcgen05.mma cta_group::1 kind::f16), warps 2-5 = epilogue (tcgen05.ld of their 32-lane TMEM quarter).  The im2col
// is done by TMA itself: each (tap, 64-channel chunk) of the K loop is one 4-D box load {64 ch, 128 px, 1 row, 1 img}
// at the tap's shifted coordinates, zero-filled outside the image (= the conv's zero padding).  Two CTAs are resident
// per SM (96 KB smem, 128 TMEM columns each) so one CTA's epilogue overlaps the other's main loop.
//
// The stride-2 transposed convolution of the up layers (conv2d_resample.py:116-133) is run as its four output-parity
// phases, each an implicit GEMM over the low-resolution grid with 4/2/2/1 taps; its (2H+1)x(2W+1) result is then
// FIR-filtered (+bias, lrelu) by a bandwidth-bound kernel, exactly the reference's operation order.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>

namespace r3dp {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int kThreads = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers, bias, rgb weights*/ + 2048;

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor,
// mma_sm100_desc.hpp: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);
    d |= (uint64_t)1 << 16;                      // LBO (unused for swizzled K-major; canonical value 1)
    d |= (uint64_t)(1024 >> 4) << 32;            // SBO
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct Taps {
    int n;
    int dy[9], dx[9], widx[9];
};
enum Mode { kStoreAct = 0, kStoreRaw = 1, kToRgbFinal = 2 };
struct ConvArgs {
    Taps taps;
    int k_chunks;            // Cin_pad / 64
    int tiles_x, rows;       // grid of M tiles: rows x tiles_x (each tile = 128 consecutive grid columns of one row)
    int w_shared;            // 1: all samples use weight set 0
    int mode;
    // output (fp16 NHWC) for kStoreAct / kStoreRaw: pixel (row*oy_mul+oy_off, col*ox_mul+ox_off) of an out_H x out_W x out_C image
    __half* out; int out_H, out_W, out_C, oy_mul, oy_off, ox_mul, ox_off;
    const float* bias;       // [Cout] conv bias (kStoreAct, kToRgbFinal)
    // kToRgbFinal: img_out[n][3][out_H][out_W] = upsample2d(img_prev[n][3][out_H/2][out_W/2]) + torgb(act) + brgb
    const float* wrgb;       // [Nw][3][128] folded ToRGB weights (fp32)
    const float* brgb;       // [3]
    const float* img_prev;
    float* img_out;
};

__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, const ConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stage_base = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* s_bias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 128);       // [128]
    float* s_wrgb = s_bias + 128;                                                       // [3][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x / a.tiles_x, col0 = (blockIdx.x % a.tiles_x) * BM;
    const int nblk = blockIdx.y, n = blockIdx.z;
    const int wn = a.w_shared ? 0 : n;
    const int num_kb = a.taps.n * a.k_chunks;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        const int t = threadIdx.x - 64;
        if (a.bias) s_bias[t] = a.bias[nblk * BN + t];
        if (a.mode == kToRgbFinal) { for (int e = t; e < 3 * BN; e += 128) s_wrgb[e] = a.wrgb[(size_t)wn * 3 * BN + e]; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES, ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                const int t = kb / a.k_chunks, kc = kb - t * a.k_chunks;
                uint8_t* sa = stage_base + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                tma_load_4d(sa, &tmA, &full[s], kc * BK, col0 + a.taps.dx[t], row + a.taps.dy[t], n);
                tma_load_4d(sa + A_BYTES, &tmB, &full[s], kc * BK, nblk * BN, a.taps.widx[t], wn);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_u32(stage_base + s * STAGE_BYTES);
                const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    // advancing K by 16 fp16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr>>4) field
                    tc_mma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc, (kb | k) != 0);
                }
                tc_commit(&empty[s]);                       // frees the smem stage when these MMAs retire
                if (kb == num_kb - 1) tc_commit(tmem_full);  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: warps 2..5 own TMEM lanes [32*(warp%4), +32) =====
        const int q = warp & 3;
        const int m = q * 32 + lane;                         // row of the tile = grid column col0 + m
        const int gcol = col0 + m;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float rgb[3] = {0.f, 0.f, 0.f};
        const int Y = row * a.oy_mul + a.oy_off, X = gcol * a.ox_mul + a.ox_off;
        const bool in_img = (Y < a.out_H) && (X < a.out_W);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tc_ld32(taddr + c0, r);
            if (a.mode == kStoreRaw) {
                if (in_img) {
                    uint4* dst = reinterpret_cast<uint4*>(a.out + (((size_t)n * a.out_H + Y) * a.out_W + X) * a.out_C + nblk * BN + c0);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        __half2 h0 = __floats2half2_rn(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1]));
                        __half2 h1 = __floats2half2_rn(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3]));
                        __half2 h2 = __floats2half2_rn(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5]));
                        __half2 h3 = __floats2half2_rn(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7]));
                        uint4 pk;
                        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                        dst[v] = pk;
                    }
                }
            } else {
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float v = __uint_as_float(r[j]) + s_bias[c0 + j];
                    f[j] = (v < 0.f ? v * 0.2f : v) * 1.4142135623730951f;        // bias_act lrelu, gain sqrt(2)
                }
                if (a.mode == kStoreAct) {
                    if (in_img) {
                        uint4* dst = reinterpret_cast<uint4*>(a.out + (((size_t)n * a.out_H + Y) * a.out_W + X) * a.out_C + nblk * BN + c0);
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            __half2 h0 = __floats2half2_rn(f[8 * v + 0], f[8 * v + 1]), h1 = __floats2half2_rn(f[8 * v + 2], f[8 * v + 3]);
                            __half2 h2 = __floats2half2_rn(f[8 * v + 4], f[8 * v + 5]), h3 = __floats2half2_rn(f[8 * v + 6], f[8 * v + 7]);
                            uint4 pk;
                            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                            dst[v] = pk;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        rgb[0] = fmaf(f[j], s_wrgb[c0 + j], rgb[0]);
                        rgb[1] = fmaf(f[j], s_wrgb[BN + c0 + j], rgb[1]);
                        rgb[2] = fmaf(f[j], s_wrgb[2 * BN + c0 + j], rgb[2]);
                    }
                }
            }
        }
        if (a.mode == kToRgbFinal && in_img) {
            const int h = a.out_H / 2, w = a.out_W / 2;
            const float k4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float acc = 0.f;
                if (a.img_prev) {
                    const float* ip = a.img_prev + ((size_t)n * 3 + c) * h * w;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int zy = Y + u - 2;
                        if (zy < 0 || (zy & 1) || (zy >> 1) >= h) continue;
                        float rowv = 0.f;
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int zx = X + v - 2;
                            if (zx < 0 || (zx & 1) || (zx >> 1) >= w) continue;
                            rowv = fmaf(k4[v], __ldg(ip + (size_t)(zy >> 1) * w + (zx >> 1)), rowv);
                        }
                        acc = fmaf(k4[u], rowv, acc);
                    }
                }
                a.img_out[(((size_t)n * 3 + c) * a.out_H + Y) * a.out_W + X] = rgb[c] + a.brgb[c] + acc;
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

// ---- helpers around the GEMMs ---------------------------------------------------------------------------------------
// wf fp32 [Nw][O][I][3][3] -> packed fp16 [Nw][9][O][Ip]  (zero for i >= I)
__global__ void pack_weights_kernel(const float* __restrict__ wf, int Nw, int O, int I, int Ip, __half* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)Nw * 9 * O * Ip;
    if (idx >= total) return;
    const int i = (int)(idx % Ip); const int o = (int)((idx / Ip) % O); const int t = (int)((idx / ((long long)Ip * O)) % 9);
    const int nw = (int)(idx / ((long long)Ip * O * 9));
    out[idx] = __float2half_rn(i < I ? wf[(((size_t)nw * O + o) * I + i) * 9 + t] : 0.f);
}

// bilinear up-resize (or copy when size == h) of NCHW fp32 -> NHWC fp16 with channel padding to Cp
__global__ void resize_to_nhwc_f16_kernel(const float* __restrict__ x, int N, int C, int h, int w, int size, int Cp, __half* __restrict__ y) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * size * size * Cp) return;
    const int c = (int)(idx % Cp); const int ox = (int)((idx / Cp) % size); const int oy = (int)((idx / ((long long)Cp * size)) % size);
    const int n = (int)(idx / ((long long)Cp * size * size));
    float v = 0.f;
    if (c < C) {
        const float sy = fmaxf(((float)oy + 0.5f) * ((float)h / (float)size) - 0.5f, 0.f);
        const float sx = fmaxf(((float)ox + 0.5f) * ((float)w / (float)size) - 0.5f, 0.f);
        const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1), y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
        const float ty = sy - (float)y0, tx = sx - (float)x0;
        const float* p = x + ((size_t)n * C + c) * h * w;
        const float r0 = p[y0 * w + x0] * (1.f - ty) + p[y1 * w + x0] * ty;
        const float r1 = p[y0 * w + x1] * (1.f - ty) + p[y1 * w + x1] * ty;
        v = r0 * (1.f - tx) + r1 * tx;
    }
    y[idx] = __float2half_rn(v);
}

// last column X = 2W of the transposed-conv result (the only part of the (2H+1)x(2W+1) grid the 128-wide GEMM tiles do not
// cover): yb[n][Y][2W][co] = sum_{ci, ky == Y (mod 2)} x[(Y-ky)/2][W-1][ci] * w[ky*3+2][co][ci].  One warp per (Y, co).
__global__ void upconv_edge_kernel(const __half* __restrict__ x, const __half* __restrict__ wp, int H, int W, int Cp, int O,
                                   int w_shared, __half* __restrict__ yb) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.y;
    const int BH = 2 * H + 1, BW = 2 * W + 1;
    if (gw >= BH * O) return;
    const int Y = gw / O, co = gw - Y * O;
    const int wn = w_shared ? 0 : n;
    float acc = 0.f;
    for (int ky = (Y & 1); ky < 3; ky += 2) {
        const int iy = (Y - ky) >> 1;
        if ((Y - ky) < 0 || iy >= H) continue;
        const __half* xp = x + (((size_t)n * H + iy) * W + (W - 1)) * Cp;
        const __half* wq = wp + (((size_t)wn * 9 + ky * 3 + 2) * O + co) * Cp;
        for (int c = lane; c < Cp; c += 32) acc = fmaf(__half2float(xp[c]), __half2float(wq[c]), acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) yb[(((size_t)n * BH + Y) * BW + 2 * W) * O + co] = __float2half_rn(acc);
}

// FIR 4x4 (pad 1, gain 4) + bias + lrelu*sqrt2 on the transposed-conv result: yb [N][2H+1][2W+1][C] fp16 -> y [N][2H][2W][C] fp16.
// One thread = one output pixel x 8 channels (16-byte vectors); neighbouring threads share taps through L1.
__global__ void __launch_bounds__(256) fir_bias_lrelu_f16_kernel(const __half* __restrict__ yb, const float* __restrict__ bias, int N, int OH,
                                                                 int OW, int C, __half* __restrict__ y) {
    const int cv = C / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * OH * OW * cv) return;
    const int c8 = (int)(idx % cv); const int ox = (int)((idx / cv) % OW); const int oy = (int)((idx / ((long long)cv * OW)) % OH);
    const int n = (int)(idx / ((long long)cv * OW * OH));
    const int BH = OH + 1, BW = OW + 1;
    const float k4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int yy = oy + u - 1;
        if ((unsigned)yy >= (unsigned)BH) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int xx = ox + v - 1;
            if ((unsigned)xx >= (unsigned)BW) continue;
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(yb + (((size_t)n * BH + yy) * BW + xx) * C + c8 * 8));
            const __half2* h = reinterpret_cast<const __half2*>(&raw);
            const float wgt = k4[u] * k4[v];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                acc[2 * j] = fmaf(wgt, f.x, acc[2 * j]); acc[2 * j + 1] = fmaf(wgt, f.y, acc[2 * j + 1]);
            }
        }
    }
    uint4 pk; __half2* ph = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a0 = acc[2 * j] + bias[c8 * 8 + 2 * j], a1 = acc[2 * j + 1] + bias[c8 * 8 + 2 * j + 1];
        a0 = (a0 < 0.f ? a0 * 0.2f : a0) * 1.4142135623730951f; a1 = (a1 < 0.f ? a1 * 0.2f : a1) * 1.4142135623730951f;
        ph[j] = __floats2half2_rn(a0, a1);
    }
    *reinterpret_cast<uint4*>(y + (((size_t)n * OH + oy) * OW + ox) * C + c8 * 8) = pk;
}

// ToRGB for the first block: x NHWC fp16 [N][H][W][C] -> img_out NCHW fp32 = upsample2d(img_prev) + conv1x1 + bias.  One warp per
// 32 consecutive pixels is wasteful on loads, so: one thread per pixel, 16-byte channel vectors, weights in smem.
__global__ void __launch_bounds__(256) torgb_f16_kernel(const __half* __restrict__ x, const float* __restrict__ wrgb, const float* __restrict__ brgb,
                                                        const float* __restrict__ img_prev, int H, int W, int C, int w_shared,
                                                        float* __restrict__ img_out) {
    extern __shared__ float s_w[];                               // [3][C]
    const int n = blockIdx.y;
    const int wn = w_shared ? 0 : n;
    for (int e = threadIdx.x; e < 3 * C; e += blockDim.x) s_w[e] = wrgb[(size_t)wn * 3 * C + e];
    __syncthreads();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= H * W) return;
    const int Y = pix / W, X = pix - Y * W;
    const uint4* xp = reinterpret_cast<const uint4*>(x + ((size_t)n * H * W + pix) * C);
    float r = 0.f, g = 0.f, b = 0.f;
    for (int c8 = 0; c8 < C / 8; ++c8) {
        const uint4 raw = __ldg(xp + c8);
        const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            const int c = c8 * 8 + 2 * j;
            r = fmaf(f.x, s_w[c], r); r = fmaf(f.y, s_w[c + 1], r);
            g = fmaf(f.x, s_w[C + c], g); g = fmaf(f.y, s_w[C + c + 1], g);
            b = fmaf(f.x, s_w[2 * C + c], b); b = fmaf(f.y, s_w[2 * C + c + 1], b);
        }
    }
    float out[3] = {r + brgb[0], g + brgb[1], b + brgb[2]};
    if (img_prev) {
        const int h = H / 2, w = W / 2;
        const float k4[4] = {0.25f, 0.75f, 0.75f, 0.25f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* ip = img_prev + ((size_t)n * 3 + c) * h * w;
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int zy = Y + u - 2;
                if (zy < 0 || (zy & 1) || (zy >> 1) >= h) continue;
                float rowv = 0.f;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int zx = X + v - 2;
                    if (zx < 0 || (zx & 1) || (zx >> 1) >= w) continue;
                    rowv = fmaf(k4[v], ip[(size_t)(zy >> 1) * w + (zx >> 1)], rowv);
                }
                acc = fmaf(k4[u], rowv, acc);
            }
            out[c] += acc;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) img_out[((size_t)n * 3 + c) * H * W + pix] = out[c];
}

// ---- host side: tensor maps -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// fp16 tensor [d3][d2][d1][d0] (d0 innermost, dense), box {64, box1, 1, 1}, 128-byte swizzle, zero fill outside
static int make_map_4d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint32_t box1) {
    EncodeTiledFn fn = encode_fn();
    R3DP_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {d0, d1, d2, d3};
    cuuint64_t strides[3] = {d0 * 2, d0 * d1 * 2, d0 * d1 * d2 * 2};
    cuuint32_t box[4] = {64, box1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    R3DP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu x %llu x %llu)", (int)r,
                 (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)d3);
    return 0;
}

static int launch_conv(const void* x, int N, int H, int W, int Cp, const void* wp, int Nw, int O, ConvArgs a, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    if (make_map_4d(&tmA, x, (uint64_t)Cp, (uint64_t)W, (uint64_t)H, (uint64_t)N, BM)) return 1;
    if (make_map_4d(&tmB, wp, (uint64_t)Cp, (uint64_t)O, 9, (uint64_t)Nw, BN)) return 1;
    a.k_chunks = Cp / BK;
    a.w_shared = (Nw == 1);
    static bool attr_set = false;
    if (!attr_set) {
        R3DP_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(a.tiles_x * a.rows, O / BN, N);
    conv_tc_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(tmA, tmB, a);
    R3DP_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // namespace tc
}  // namespace r3dp

using namespace r3dp;
using namespace r3dp::tc;

extern "C" int r3dp_sr_tc_pack_weights(const float* wf, int Nw, int O, int I, void* packed_f16, r3dp_stream_t stream) {
    R3DP_REQUIRE(wf && packed_f16, "sr_tc_pack_weights: null pointer");
    R3DP_REQUIRE(Nw > 0 && O > 0 && I > 0, "sr_tc_pack_weights: bad shape");
    const int Ip = (I + 63) / 64 * 64;
    const long long total = (long long)Nw * 9 * O * Ip;
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(wf, Nw, O, I, Ip, reinterpret_cast<__half*>(packed_f16));
    R3DP_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

extern "C" int r3dp_sr_tc_input(const float* x, int N, int C, int h, int w, int size, void* y_f16, r3dp_stream_t stream) {
    R3DP_REQUIRE(x && y_f16, "sr_tc_input: null pointer");
    R3DP_REQUIRE(N > 0 && C > 0 && h > 0 && w > 0 && size >= h && size >= w, "sr_tc_input: up-scaling (or copy) only");
    const int Cp = (C + 63) / 64 * 64;
    const long long total = (long long)N * size * size * Cp;
    resize_to_nhwc_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, N, C, h, w, size, Cp, reinterpret_cast<__half*>(y_f16));
    R3DP_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

extern "C" size_t r3dp_sr_tc_scratch_bytes(int N, int O, int H, int W) { return (size_t)N * (2 * H + 1) * (2 * W + 1) * O * sizeof(__half); }

// SynthesisLayer on tensor cores.  x [N][H][W][Ip] fp16 NHWC (Ip = I rounded up to 64), wp packed weights [Nw][9][O][Ip] fp16
// (Nw == N per-sample, or 1 shared), bias [O] fp32.  up == 1: y [N][H][W][O]; up == 2: y [N][2H][2W][O], scratch >= r3dp_sr_tc_scratch_bytes.
extern "C" int r3dp_sr_tc_layer(const void* x_f16, const void* wp_f16, const float* bias, int N, int Nw, int I, int O, int H, int W, int up,
                                void* y_f16, void* scratch, r3dp_stream_t stream) {
    R3DP_REQUIRE(x_f16 && wp_f16 && bias && y_f16, "sr_tc_layer: null pointer");
    R3DP_REQUIRE(N > 0 && (Nw == N || Nw == 1) && I > 0 && H > 0, "sr_tc_layer: bad shape");
    R3DP_REQUIRE(W % BM == 0 && O % BN == 0, "sr_tc_layer: needs W %% 128 == 0 and Cout %% 128 == 0 (got W=%d, Cout=%d)", W, O);
    R3DP_REQUIRE(up == 1 || up == 2, "sr_tc_layer: up must be 1 or 2");
    const int Ip = (I + 63) / 64 * 64;
    cudaStream_t st = as_stream(stream);
    ConvArgs a = {};
    a.bias = bias;
    if (up == 1) {
        a.taps.n = 9;
        for (int t = 0; t < 9; ++t) { a.taps.dy[t] = t / 3 - 1; a.taps.dx[t] = t % 3 - 1; a.taps.widx[t] = t; }
        a.tiles_x = W / BM; a.rows = H; a.mode = kStoreAct;
        a.out = reinterpret_cast<__half*>(y_f16); a.out_H = H; a.out_W = W; a.out_C = O; a.oy_mul = a.ox_mul = 1;
        return launch_conv(x_f16, N, H, W, Ip, wp_f16, Nw, O, a, st);
    }
    R3DP_REQUIRE(scratch, "sr_tc_layer: up=2 needs scratch");
    __half* yb = reinterpret_cast<__half*>(scratch);
    a.mode = kStoreRaw; a.out = yb; a.out_H = 2 * H + 1; a.out_W = 2 * W + 1; a.out_C = O; a.oy_mul = a.ox_mul = 2;
    a.tiles_x = W / BM;
    for (int pa = 0; pa < 2; ++pa)
        for (int pb = 0; pb < 2; ++pb) {
            a.oy_off = pa; a.ox_off = pb; a.rows = pa ? H : H + 1;
            a.taps.n = 0;
            for (int ky = pa; ky < 3; ky += 2)
                for (int kx = pb; kx < 3; kx += 2) {
                    const int t = a.taps.n++;
                    a.taps.dy[t] = -(ky >> 1); a.taps.dx[t] = -(kx >> 1); a.taps.widx[t] = ky * 3 + kx;
                }
            if (launch_conv(x_f16, N, H, W, Ip, wp_f16, Nw, O, a, st)) return 1;
        }
    {
        const int warps = (2 * H + 1) * O;
        dim3 grid((warps * 32 + 255) / 256, N);
        upconv_edge_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(x_f16), reinterpret_cast<const __half*>(wp_f16), H, W, Ip, O,
                                                 Nw == 1, yb);
    }
    {
        const long long total = (long long)N * (2 * H) * (2 * W) * (O / 8);
        fir_bias_lrelu_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(yb, bias, N, 2 * H, 2 * W, O, reinterpret_cast<__half*>(y_f16));
    }
    R3DP_LAUNCH_CHECK();
    count_launches(2);
    return 0;
}

// Last layer fused with ToRGB: conv3x3 (I -> 128) + bias + lrelu, then img_out = upsample2d(img_prev) + torgb + brgb; the 128-channel
// activation itself is never written (SynthesisBlock is_last: only the image leaves the block).
extern "C" int r3dp_sr_tc_last_layer(const void* x_f16, const void* wp_f16, const float* bias, const float* wrgb, const float* brgb,
                                     const float* img_prev, int N, int Nw, int I, int H, int W, float* img_out, r3dp_stream_t stream) {
    R3DP_REQUIRE(x_f16 && wp_f16 && bias && wrgb && brgb && img_out, "sr_tc_last_layer: null pointer");
    R3DP_REQUIRE(N > 0 && (Nw == N || Nw == 1) && W % BM == 0 && H % 2 == 0, "sr_tc_last_layer: bad shape");
    const int Ip = (I + 63) / 64 * 64;
    ConvArgs a = {};
    a.bias = bias; a.wrgb = wrgb; a.brgb = brgb; a.img_prev = img_prev; a.img_out = img_out;
    a.taps.n = 9;
    for (int t = 0; t < 9; ++t) { a.taps.dy[t] = t / 3 - 1; a.taps.dx[t] = t % 3 - 1; a.taps.widx[t] = t; }
    a.tiles_x = W / BM; a.rows = H; a.mode = kToRgbFinal; a.out_H = H; a.out_W = W; a.out_C = BN; a.oy_mul = a.ox_mul = 1;
    return launch_conv(x_f16, N, H, W, Ip, wp_f16, Nw, BN, a, as_stream(stream));
}

// ToRGB of a non-final block: x NHWC fp16 [N][H][W][C] -> img_out NCHW fp32 [N][3][H][W] (+ upsample2d(img_prev) + bias).
extern "C" int r3dp_sr_tc_torgb(const void* x_f16, const float* wrgb, const float* brgb, const float* img_prev, int N, int Nw, int C, int H,
                                int W, float* img_out, r3dp_stream_t stream) {
    R3DP_REQUIRE(x_f16 && wrgb && brgb && img_out, "sr_tc_torgb: null pointer");
    R3DP_REQUIRE(N > 0 && (Nw == N || Nw == 1) && C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "sr_tc_torgb: bad shape");
    dim3 grid((H * W + 255) / 256, N);
    torgb_f16_kernel<<<grid, 256, 3 * C * sizeof(float), as_stream(stream)>>>(reinterpret_cast<const __half*>(x_f16), wrgb, brgb, img_prev, H, W, C,
                                                                              Nw == 1, img_out);
    R3DP_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}
