// zs_orb.cu -- cv::ORB::create()->compute(image, keypoints, descriptors) for provided keypoints
// (reference call: zenslam_core/source/detection/keypoint_detector_grid.cpp:138; semantics SURVEY A.2-A.4).
//
//   k_orb_blur      7x7 sigma-2 float32 separable blur with OpenCV's exact FMA order (A.3), tiled through
//                   shared memory; reads the REFLECT_101-padded level-0 plane, writes the blurred plane.
//                   Streaming, HBM-bound: 2 bytes per pixel.
//   k_orb_filter    order-preserving compaction of the keypoints that survive the 31-px border filter.
//   k_orb_describe  warp per keypoint, lane = descriptor byte: 8 tests x 2 gathers from the (L2-resident)
//                   blurred plane; pattern rows fetched as int4 from a 1 KB table.
#include <math.h>

#include "zs_common.cuh"
#include "../../include/zs_orb_pattern.h"

#define BLUR_TW 64
#define BLUR_TH 32

// taps of cv::getGaussianKernel(7, 2, CV_32F)
#define G0 0x1.1f5f62p-4f
#define G1 0x1.0c70fcp-3f
#define G2 0x1.869472p-3f
#define G3 0x1.ba95c0p-3f

// u8 -> float without a conversion instruction: drop the byte into the mantissa of 2^23 and subtract 2^23 (exact)
__device__ __forceinline__ float u8f(uint32_t w, int k)
{
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + k)), 8388608.f);     // 0x4B0000bb = 2^23 + bb
}

// grid: (ceil(w/64), ceil(h/32), count), 256 threads.  Four horizontally adjacent outputs per thread in both
// passes: the row pass reads 3 words of the staged u8 tile and writes one float4, the column pass reads seven
// float4 and writes one u8x4 word.  The per-output arithmetic (and its order) is OpenCV's, see SURVEY A.3.
__global__ void __launch_bounds__(256) k_orb_blur(zs_pyr_view v, int first)
{
    __shared__ __align__(16) uint8_t s_in[BLUR_TH + 6][BLUR_TW + 8];     // 38 x 72
    __shared__ __align__(16) float s_row[BLUR_TH + 6][BLUR_TW];          // row-pass results
    const int w = v.w[0], h = v.h[0], pitch = v.pitch[0];
    const int slot = zs_slot(first, blockIdx.z, v.slots);
    const int x0 = blockIdx.x * BLUR_TW, y0 = blockIdx.y * BLUR_TH;
    const uint8_t* src = v.img[0] + (size_t)slot * v.slot_stride[0] + (size_t)(v.pad_y + y0 - 3) * pitch + v.pad_x + x0 - 4;
    // load (BLUR_TH+6) rows x 72 bytes (columns x0-4 .. x0+67) as 32-bit words; the padded plane makes every
    // address valid (pad_x >= 16 and the pitch is rounded up to 128 beyond w + 2 pad_x)
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * 18; i += 256) {
        const int r = i / 18, c = i - r * 18;
        uint32_t val = 0;
        if (y0 - 3 + r < h + 3)                        // rows past the bottom padding we need are never used
            val = *(const uint32_t*)(src + (size_t)r * pitch + 4 * c);
        *(uint32_t*)&s_in[r][4 * c] = val;
    }
    __syncthreads();
    // row pass: s = g0*x0; s = fma(x_i, g_i, s), i = 1..6.  Output column c needs staged bytes c+1 .. c+7.
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * (BLUR_TW / 4); i += 256) {
        const int r = i / (BLUR_TW / 4), c = (i - r * (BLUR_TW / 4)) * 4;
        const uint32_t* wp = (const uint32_t*)&s_in[r][c];
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
        float f[10];                                        // staged bytes c+1 .. c+10
        f[0] = u8f(w0, 1); f[1] = u8f(w0, 2); f[2] = u8f(w0, 3);
        f[3] = u8f(w1, 0); f[4] = u8f(w1, 1); f[5] = u8f(w1, 2); f[6] = u8f(w1, 3);
        f[7] = u8f(w2, 0); f[8] = u8f(w2, 1); f[9] = u8f(w2, 2);
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float t = __fmul_rn(G0, f[k]);
            t = fmaf(f[k + 1], G1, t);
            t = fmaf(f[k + 2], G2, t);
            t = fmaf(f[k + 3], G3, t);
            t = fmaf(f[k + 4], G2, t);
            t = fmaf(f[k + 5], G1, t);
            t = fmaf(f[k + 6], G0, t);
            o[k] = t;
        }
        *(float4*)&s_row[r][c] = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    // column pass: s = g3*r0; s = fma(r_{+k} + r_{-k}, g_{3+k}, s), k = 1..3; round half even, saturate
    uint8_t* dst = v.blur + (size_t)slot * v.blur_slot;
    for (int i = threadIdx.x; i < BLUR_TH * (BLUR_TW / 4); i += 256) {
        const int r = i / (BLUR_TW / 4), c = (i - r * (BLUR_TW / 4)) * 4;
        const int x = x0 + c, y = y0 + r;
        if (x >= w || y >= h) continue;
        float4 rr[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) rr[k] = *(const float4*)&s_row[r + k][c];
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a0 = ((const float*)&rr[0])[k], a1 = ((const float*)&rr[1])[k], a2 = ((const float*)&rr[2])[k],
                        a3 = ((const float*)&rr[3])[k], a4 = ((const float*)&rr[4])[k], a5 = ((const float*)&rr[5])[k],
                        a6 = ((const float*)&rr[6])[k];
            float t = __fmul_rn(G3, a3);
            t = fmaf(__fadd_rn(a4, a2), G2, t);
            t = fmaf(__fadd_rn(a5, a1), G1, t);
            t = fmaf(__fadd_rn(a6, a0), G0, t);
            int q = __float2int_rn(t);
            q = q < 0 ? 0 : q > 255 ? 255 : q;
            packed |= (uint32_t)q << (8 * k);
        }
        uint8_t* o = dst + (size_t)y * v.blur_pitch + x;
        if (x + 4 <= w) *(uint32_t*)o = packed;            // blur_pitch is a multiple of 128 and x of 4
        else for (int k = 0; x + k < w; ++k) o[k] = (uint8_t)(packed >> (8 * k));
    }
}

zs_status zs_launch_orb_blur(zs_context* ctx, const zs_pyramid* p, int first, int count)
{
    if (count == 0) return ZS_OK;
    k_orb_blur<<<dim3(zs_div_up(p->width, BLUR_TW), zs_div_up(p->height, BLUR_TH), count), 256, 0, ctx->stream>>>(p->v, first);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// Border filter + order-preserving compaction (KeyPointsFilter::runByImageBorder with edge 31):
// keep iff 31 <= cvRound(x) < w-31 and 31 <= cvRound(y) < h-31.  One block per image.
__global__ void __launch_bounds__(1024) k_orb_filter(const float2* __restrict__ xyi,
                                                     const float* __restrict__ ri, const float* __restrict__ ai,
                                                     const int* __restrict__ ni, int cap, int w, int h,
                                                     float2* __restrict__ xyo,
                                                     float* __restrict__ ro, float* __restrict__ ao,
                                                     int* __restrict__ src_index, int* __restrict__ no)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int img = blockIdx.x;
    const size_t off = (size_t)img * cap;
    const int n = min(ni[img], cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        float x = 0.f, y = 0.f;
        bool f = false;
        if (i < n) {
            const float2 xy = xyi[off + i]; x = xy.x; y = xy.y;
            const int rx = __float2int_rn(x), ry = __float2int_rn(y);
            f = (w > 62 && h > 62) && rx >= 31 && rx < w - 31 && ry >= 31 && ry < h - 31;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        const int wpre = __popc(m & ((1u << lane) - 1));
        if (lane == 0) warp_sums[warp] = __popc(m);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
        if (f) {
            const int pos = carry + woff + wpre;
            xyo[off + pos] = make_float2(x, y);
            if (ro) ro[off + pos] = ri ? ri[off + i] : 0.f;
            if (ao) ao[off + pos] = ai ? ai[off + i] : -1.f;
            if (src_index) src_index[off + pos] = i;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) no[img] = carry;
}

__device__ int4 g_orb_pattern[256] = ZS_ORB_PATTERN_INIT;   // global copy: lane-divergent reads go through L1

// grid: (ceil(cap/8), count), 256 threads = 8 warps = 8 keypoints.
// The 512 sample points of a keypoint lie within 19 px of its centre (pattern radius 18.4 + rounding), and the border
// filter keeps every centre >= 31 px inside the image, so the warp first stages the 39 x 44-byte neighbourhood with
// coalesced 32-bit loads (39 rows x 11 words = ~14 loads per lane, each row one or two 32-byte sectors) and then
// gathers from shared memory -- instead of 512 single-byte gathers that each touch their own sector in L2.
#define ORB_R 19
#define ORB_PW 11                                      // words per staged row (39 + up to 3 alignment bytes <= 44)
#define ORB_ROWS (2 * ORB_R + 1)
// cos / sin of the keypoint angles, once per keypoint (double math like OpenCV; FP64 is scarce on B200, so this
// is kept out of the describe kernel, where every lane of a warp would repeat it)
__global__ void k_orb_trig(const float* __restrict__ angles, const int* __restrict__ counts, int cap, float2* __restrict__ ab)
{
    const int img = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(counts[img], cap)) return;
    const size_t off = (size_t)img * cap + i;
    // OpenCV: angle *= (float)(CV_PI/180.f); a = (float)cos(angle), b = (float)sin(angle)  (double math)
    const float angle = __fmul_rn(angles[off], (float)(3.14159265358979323846 / 180.0));
    ab[off] = make_float2((float)cos((double)angle), (float)sin((double)angle));
}

// ab: per-keypoint (cos, sin) or null, in which case every keypoint uses (a0, b0) -- the constants for the angle
// -1 degree that FAST keypoints carry, computed once on the host
__global__ void __launch_bounds__(256) k_orb_describe(zs_pyr_view v, int first, const float2* __restrict__ xys,
                                                      const float2* __restrict__ ab, float a0, float b0,
                                                      const int* __restrict__ counts, int cap, uint8_t* __restrict__ desc)
{
    __shared__ uint32_t s_patch[8][ORB_ROWS * ORB_PW];
    __shared__ int4 s_pattern[256];                    // transposed: [bit][lane], so a warp reads 512 contiguous bytes per bit
    const int img = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const int kp = blockIdx.x * 8 + warp;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x * 8 >= counts[img]) return;         // whole block idle
    // (the table row a lane needs for bit b is 8*lane + b: read straight from global memory that is one 16-byte
    // piece out of 32 different cache lines per load -- the kernel used to spend most of its time there)
    {
        int4 pt = g_orb_pattern[threadIdx.x];
        if (!ab) {
            // constant angle: store the two byte offsets (relative to the patch centre) instead of the coordinates
            const float x0f = __fsub_rn(__fmul_rn((float)pt.x, a0), __fmul_rn((float)pt.y, b0));
            const float y0f = __fadd_rn(__fmul_rn((float)pt.x, b0), __fmul_rn((float)pt.y, a0));
            const float x1f = __fsub_rn(__fmul_rn((float)pt.z, a0), __fmul_rn((float)pt.w, b0));
            const float y1f = __fadd_rn(__fmul_rn((float)pt.z, b0), __fmul_rn((float)pt.w, a0));
            pt.x = __float2int_rn(y0f) * (ORB_PW * 4) + __float2int_rn(x0f);
            pt.y = __float2int_rn(y1f) * (ORB_PW * 4) + __float2int_rn(x1f);
        }
        s_pattern[(threadIdx.x & 7) * 32 + (threadIdx.x >> 3)] = pt;
    }
    __syncthreads();
    if (kp >= counts[img]) return;
    const size_t off = (size_t)img * cap + kp;
    const int slot = zs_slot(first, img, v.slots);
    const int pitch = v.blur_pitch;
    const float2 xy = xys[off];
    const int cx = __float2int_rn(xy.x), cy = __float2int_rn(xy.y);
    const int x0 = (cx - ORB_R) & ~3;                  // first staged column (4-byte aligned; the plane pitch is a multiple of 128)
    const uint8_t* src = v.blur + (size_t)slot * v.blur_slot + (size_t)(cy - ORB_R) * pitch + x0;
    uint32_t* sp = s_patch[warp];
    // 429 words = 13 full rounds of 32 lanes + 13 words: all loads are issued before the first store
    uint32_t stage[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) {
        const int i = lane + 32 * k;
        const int r = i / ORB_PW, cc = i - r * ORB_PW;
        stage[k] = (i < ORB_ROWS * ORB_PW) ? *(const uint32_t*)(src + (size_t)r * pitch + 4 * cc) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 14; ++k) {
        const int i = lane + 32 * k;
        if (i < ORB_ROWS * ORB_PW) sp[i] = stage[k];
    }
    __syncwarp();
    const uint8_t* c = (const uint8_t*)sp + ORB_R * (ORB_PW * 4) + (cx - x0);     // the centre pixel inside the staged patch
    int val = 0;
    if (ab) {
        const float2 t = ab[off];
        const float a = t.x, b = t.y;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
            const int4 pt = s_pattern[bit * 32 + lane];
            const float x0f = __fsub_rn(__fmul_rn((float)pt.x, a), __fmul_rn((float)pt.y, b));
            const float y0f = __fadd_rn(__fmul_rn((float)pt.x, b), __fmul_rn((float)pt.y, a));
            const float x1f = __fsub_rn(__fmul_rn((float)pt.z, a), __fmul_rn((float)pt.w, b));
            const float y1f = __fadd_rn(__fmul_rn((float)pt.z, b), __fmul_rn((float)pt.w, a));
            const int t0 = c[__float2int_rn(y0f) * (ORB_PW * 4) + __float2int_rn(x0f)];
            const int t1 = c[__float2int_rn(y1f) * (ORB_PW * 4) + __float2int_rn(x1f)];
            val |= (t0 < t1) << bit;
        }
    } else {
        // every keypoint carries the same angle (FAST keypoints: -1 degree): the rotated, rounded sample offsets were
        // computed once per block into s_pattern (see above), so a bit costs two shared loads and a compare
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
            const int4 pt = s_pattern[bit * 32 + lane];
            const int t0 = c[pt.x], t1 = c[pt.y];
            val |= (t0 < t1) << bit;
        }
    }
    desc[off * 32 + lane] = (uint8_t)val;
}

extern "C" zs_status zs_orb_compute(zs_context* ctx, const zs_pyramid* p, int first, int count, const float* d_xy_in,
                                    const float* d_resp_in, const float* d_angle_in,
                                    const int* d_count_in, int cap, float* d_xy, float* d_resp,
                                    int* d_src_index, int* d_count, uint8_t* d_desc)
{
    ZS_REQUIRE(ctx && p && d_xy_in && d_count_in && d_xy && d_count && d_desc, "null argument");
    ZS_REQUIRE(count >= 0 && count <= p->slots && first >= 0 && cap > 0, "bad range");
    ZS_REQUIRE(d_xy != d_xy_in, "in-place compaction is not supported");
    if (count == 0) return ZS_OK;
    zs_status st = zs_launch_orb_blur(ctx, p, first, count);
    if (st != ZS_OK) return st;
    float* d_angle = nullptr;
    if (d_angle_in) {
        void* s;
        st = zs_scratch(ctx, sizeof(float) * 3 * (size_t)cap * count + 64, &s);
        if (st != ZS_OK) return st;
        d_angle = (float*)s;
    }
    k_orb_filter<<<count, 1024, 0, ctx->stream>>>((const float2*)d_xy_in, d_resp_in, d_angle_in, d_count_in, cap, p->width,
                                                  p->height, (float2*)d_xy, d_resp, d_angle, d_src_index, d_count);
    ZS_LAUNCH_CHECK(ctx);
    float2* d_ab = nullptr;
    // angle -1 degree: (float)cos / (float)sin of the double value of the float product, as OpenCV computes them
    const float ang = -1.f * (float)(3.14159265358979323846 / 180.0);
    const float a0 = (float)cos((double)ang), b0 = (float)sin((double)ang);
    if (d_angle) {
        // the compacted angles sit in the first half of the scratch block; (cos, sin) pairs go behind them
        d_ab = (float2*)(d_angle + (((size_t)cap * count + 3) & ~(size_t)3));
        k_orb_trig<<<dim3(zs_div_up(cap, 256), count), 256, 0, ctx->stream>>>(d_angle, d_count, cap, d_ab);
        ZS_LAUNCH_CHECK(ctx);
    }
    k_orb_describe<<<dim3(zs_div_up(cap, 8), count), 256, 0, ctx->stream>>>(p->v, first, (const float2*)d_xy, d_ab, a0, b0, d_count, cap, d_desc);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}
