// zs_orb.cu -- cv::ORB::create()->compute(image, keypoints, descriptors) for provided keypoints
// (reference call: zenslam_core/source/detection/keypoint_detector_grid.cpp:138; semantics SURVEY A.2-A.4).
//
//   k_orb_blur      7x7 sigma-2 float32 separable blur with OpenCV's exact FMA order (A.3), tiled through
//                   shared memory; reads the REFLECT_101-padded level-0 plane, writes the blurred plane.
//                   Streaming, HBM-bound: 2 bytes per pixel.
//   k_orb_filter    order-preserving compaction of the keypoints that survive the 31-px border filter.
//   k_orb_describe  warp per keypoint, lane = descriptor byte: 8 tests x 2 gathers from the (L2-resident)
//                   blurred plane; pattern rows fetched as int4 from a 1 KB table.
#include "zs_common.cuh"
#include "../../include/zs_orb_pattern.h"

#define BLUR_TW 64
#define BLUR_TH 32

// taps of cv::getGaussianKernel(7, 2, CV_32F)
#define G0 0x1.1f5f62p-4f
#define G1 0x1.0c70fcp-3f
#define G2 0x1.869472p-3f
#define G3 0x1.ba95c0p-3f

// grid: (ceil(w/64), ceil(h/32), count), 256 threads
__global__ void __launch_bounds__(256) k_orb_blur(zs_pyr_view v, int first)
{
    __shared__ uint8_t s_in[BLUR_TH + 6][BLUR_TW + 8];     // 38 x 72
    __shared__ float s_row[BLUR_TH + 6][BLUR_TW];          // row-pass results
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
    // row pass: s = g0*x0; s = fma(x_i, g_i, s), i = 1..6
    for (int i = threadIdx.x; i < (BLUR_TH + 6) * BLUR_TW; i += 256) {
        const int r = i / BLUR_TW, c = i - r * BLUR_TW;
        const uint8_t* p = &s_in[r][c + 1];                // column x0 + c - 3
        float s = __fmul_rn(G0, (float)p[0]);
        s = fmaf((float)p[1], G1, s);
        s = fmaf((float)p[2], G2, s);
        s = fmaf((float)p[3], G3, s);
        s = fmaf((float)p[4], G2, s);
        s = fmaf((float)p[5], G1, s);
        s = fmaf((float)p[6], G0, s);
        s_row[r][c] = s;
    }
    __syncthreads();
    // column pass: s = g3*r0; s = fma(r_{+k} + r_{-k}, g_{3+k}, s), k = 1..3; round half even, saturate
    uint8_t* dst = v.blur + (size_t)slot * v.blur_slot;
    for (int i = threadIdx.x; i < BLUR_TH * BLUR_TW; i += 256) {
        const int r = i / BLUR_TW, c = i - r * BLUR_TW;
        const int x = x0 + c, y = y0 + r;
        if (x >= w || y >= h) continue;
        float s = __fmul_rn(G3, s_row[r + 3][c]);
        s = fmaf(__fadd_rn(s_row[r + 4][c], s_row[r + 2][c]), G2, s);
        s = fmaf(__fadd_rn(s_row[r + 5][c], s_row[r + 1][c]), G1, s);
        s = fmaf(__fadd_rn(s_row[r + 6][c], s_row[r + 0][c]), G0, s);
        int q = __float2int_rn(s);
        q = q < 0 ? 0 : q > 255 ? 255 : q;
        dst[(size_t)y * v.blur_pitch + x] = (uint8_t)q;
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

// grid: (ceil(cap/8), count), 256 threads = 8 warps = 8 keypoints
__global__ void __launch_bounds__(256) k_orb_describe(zs_pyr_view v, int first, const float2* __restrict__ xys,
                                                      const float* __restrict__ angles,
                                                      const int* __restrict__ counts, int cap, uint8_t* __restrict__ desc)
{
    const int img = blockIdx.y;
    const int kp = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (kp >= counts[img]) return;
    const size_t off = (size_t)img * cap + kp;
    const int slot = zs_slot(first, img, v.slots);
    const int pitch = v.blur_pitch;
    const float2 xy = xys[off];
    const int cx = __float2int_rn(xy.x), cy = __float2int_rn(xy.y);
    const uint8_t* c = v.blur + (size_t)slot * v.blur_slot + (size_t)cy * pitch + cx;
    float angle = angles ? angles[off] : -1.f;
    // OpenCV: angle *= (float)(CV_PI/180.f); a = (float)cos(angle), b = (float)sin(angle)  (double math)
    angle = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.0));
    const float a = (float)cos((double)angle), b = (float)sin((double)angle);
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const int4 pt = g_orb_pattern[lane * 8 + bit];
        const float x0 = __fsub_rn(__fmul_rn((float)pt.x, a), __fmul_rn((float)pt.y, b));
        const float y0 = __fadd_rn(__fmul_rn((float)pt.x, b), __fmul_rn((float)pt.y, a));
        const float x1 = __fsub_rn(__fmul_rn((float)pt.z, a), __fmul_rn((float)pt.w, b));
        const float y1 = __fadd_rn(__fmul_rn((float)pt.z, b), __fmul_rn((float)pt.w, a));
        const int t0 = c[__float2int_rn(y0) * pitch + __float2int_rn(x0)];
        const int t1 = c[__float2int_rn(y1) * pitch + __float2int_rn(x1)];
        val |= (t0 < t1) << bit;
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
        st = zs_scratch(ctx, sizeof(float) * (size_t)cap * count, &s);
        if (st != ZS_OK) return st;
        d_angle = (float*)s;
    }
    k_orb_filter<<<count, 1024, 0, ctx->stream>>>((const float2*)d_xy_in, d_resp_in, d_angle_in, d_count_in, cap, p->width,
                                                  p->height, (float2*)d_xy, d_resp, d_angle, d_src_index, d_count);
    ZS_LAUNCH_CHECK(ctx);
    k_orb_describe<<<dim3(zs_div_up(cap, 8), count), 256, 0, ctx->stream>>>(p->v, first, (const float2*)d_xy, d_angle, d_count, cap, d_desc);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}
