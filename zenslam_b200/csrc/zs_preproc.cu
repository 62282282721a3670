// zs_preproc.cu -- the per-frame pre-processing of processor::process on the device
// (zenslam_core/source/processor.cpp:25-55):
//   utils::convert_color(image, cv::COLOR_BGR2GRAY)                       -> k_bgr2gray
//   utils::apply_clahe(image, cv::createCLAHE(4.0))   (processor.h:38)     -> k_clahe_lut + k_clahe_apply
//   utils::rectify = cv::remap(image, map_x, map_y, cv::INTER_LINEAR)      -> k_remap_linear
//   (utils.cpp:119-124; maps are CV_32FC1 from cv::initUndistortRectifyMap, calibration.cpp:60-70)
// OpenCV does all three in integer / fixed-point arithmetic for 8-bit images (the CLAHE interpolation in float32
// with a fixed operation order), so the kernels reproduce cv2 bit for bit.  Streaming, HBM-bound: the remap can
// write straight into level 0 of a zs_pyramid so the frame is never copied again.
#include "zs_common.cuh"

// ---- BGR -> gray: (B*3735 + G*19235 + R*9798 + 2^14) >> 15, four pixels (three words in, one word out) per thread
__global__ void __launch_bounds__(256) k_bgr2gray(const uint8_t* __restrict__ src, size_t spitch, size_t sstride, int w, int h,
                                                  uint8_t* __restrict__ dst, size_t dpitch, size_t dstride, int vec)
{
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* s = src + (size_t)blockIdx.z * sstride + (size_t)y * spitch + 3 * (size_t)x;
    uint8_t* d = dst + (size_t)blockIdx.z * dstride + (size_t)y * dpitch + x;
    if (vec && x + 4 <= w) {
        const uint32_t w0 = ((const uint32_t*)s)[0], w1 = ((const uint32_t*)s)[1], w2 = ((const uint32_t*)s)[2];
        const uint32_t b[12] = { w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, w0 >> 24, w1 & 255, (w1 >> 8) & 255, (w1 >> 16) & 255, w1 >> 24,
                                 w2 & 255, (w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24 };
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) out |= ((b[3 * k] * 3735u + b[3 * k + 1] * 19235u + b[3 * k + 2] * 9798u + (1u << 14)) >> 15) << (8 * k);
        *(uint32_t*)d = out;
    } else {
        for (int k = 0; k < 4 && x + k < w; ++k)
            d[k] = (uint8_t)((s[3 * k] * 3735u + s[3 * k + 1] * 19235u + s[3 * k + 2] * 9798u + (1u << 14)) >> 15);
    }
}

extern "C" zs_status zs_cvt_bgr2gray(zs_context* ctx, const uint8_t* d_bgr, size_t pitch, size_t stride, int width, int height,
                                     int count, uint8_t* d_gray, size_t gray_pitch, size_t gray_stride)
{
    ZS_REQUIRE(ctx && d_bgr && d_gray, "null argument");
    ZS_REQUIRE(width > 0 && height > 0 && count >= 0 && pitch >= 3 * (size_t)width && gray_pitch >= (size_t)width, "bad geometry");
    if (count == 0) return ZS_OK;
    const int vec = ((uintptr_t)d_bgr % 4 == 0 && pitch % 4 == 0 && stride % 4 == 0 && (uintptr_t)d_gray % 4 == 0 && gray_pitch % 4 == 0 &&
                     gray_stride % 4 == 0) ? 1 : 0;
    k_bgr2gray<<<dim3(zs_div_up(zs_div_up(width, 4), 256), height, count), 256, 0, ctx->stream>>>(d_bgr, pitch, stride, width, height,
                                                                                              d_gray, gray_pitch, gray_stride, vec);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// ---- CLAHE -------------------------------------------------------------------------------------------
struct clahe_args {
    const uint8_t* src; size_t spitch, sstride;
    uint8_t* dst; size_t dpitch, dstride;
    int w, h, tiles_x, tiles_y, tw, th, clip;
    float lut_scale, inv_tw, inv_th;
    uint8_t* lut;                 // [count][tiles_y*tiles_x][256]
};

// one block per tile: histogram (shared-memory atomics), clip + redistribute, prefix sum -> 256-entry LUT
__global__ void __launch_bounds__(256) k_clahe_lut(clahe_args a)
{
    __shared__ int hist[256];
    __shared__ int wsum[8];
    const int tile = blockIdx.x, img = blockIdx.y, t = threadIdx.x;
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    hist[t] = 0;
    __syncthreads();
    const uint8_t* s = a.src + (size_t)img * a.sstride;
    for (int i = t; i < a.tw * a.th; i += 256) {
        const int yy = i / a.tw, xx = i - yy * a.tw;
        // the image is extended REFLECT_101 on the right / bottom up to a multiple of the tile grid
        const int y = zs_reflect101(ty * a.th + yy, a.h), x = zs_reflect101(tx * a.tw + xx, a.w);
        atomicAdd(&hist[s[(size_t)y * a.spitch + x]], 1);
    }
    __syncthreads();
    int v = hist[t];
    if (a.clip > 0) {
        int excess = v > a.clip ? v - a.clip : 0;
        v -= excess;
        // block sum of the excess
        int e = excess;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if ((t & 31) == 0) wsum[t >> 5] = e;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) clipped += wsum[k];
        const int batch = clipped / 256;
        const int residual = clipped - batch * 256;
        v += batch;
        if (residual != 0) {
            const int step = max(256 / residual, 1);
            if (t % step == 0 && t / step < residual) v += 1;
        }
        __syncthreads();
    }
    // inclusive prefix sum over the 256 bins
    int s_ = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, s_, o); if ((t & 31) >= o) s_ += n; }
    if ((t & 31) == 31) wsum[t >> 5] = s_;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < (t >> 5); ++k) base += wsum[k];
    const int sum = base + s_;
    int q = __float2int_rn(__fmul_rn((float)sum, a.lut_scale));
    q = q < 0 ? 0 : q > 255 ? 255 : q;
    a.lut[((size_t)img * a.tiles_x * a.tiles_y + tile) * 256 + t] = (uint8_t)q;
}

// bilinear interpolation between the LUTs of the four surrounding tiles (float32, OpenCV's operation order)
__global__ void __launch_bounds__(256) k_clahe_apply(clahe_args a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= a.w) return;
    const float tyf = __fsub_rn(__fmul_rn((float)y, a.inv_th), 0.5f), txf = __fsub_rn(__fmul_rn((float)x, a.inv_tw), 0.5f);
    int ty1 = __float2int_rd(tyf), tx1 = __float2int_rd(txf);
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.f, ya), xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.f, xa);
    int ty2 = ty1 + 1, tx2 = tx1 + 1;
    ty1 = max(ty1, 0); ty2 = min(ty2, a.tiles_y - 1); tx1 = max(tx1, 0); tx2 = min(tx2, a.tiles_x - 1);
    const int v = a.src[(size_t)img * a.sstride + (size_t)y * a.spitch + x];
    const uint8_t* lut = a.lut + (size_t)img * a.tiles_x * a.tiles_y * 256 + v;
    const float l11 = (float)lut[(ty1 * a.tiles_x + tx1) * 256], l12 = (float)lut[(ty1 * a.tiles_x + tx2) * 256];
    const float l21 = (float)lut[(ty2 * a.tiles_x + tx1) * 256], l22 = (float)lut[(ty2 * a.tiles_x + tx2) * 256];
    const float res = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, xa)), ya1),
                                __fmul_rn(__fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, xa)), ya));
    int q = __float2int_rn(res);
    q = q < 0 ? 0 : q > 255 ? 255 : q;
    a.dst[(size_t)img * a.dstride + (size_t)y * a.dpitch + x] = (uint8_t)q;
}

extern "C" zs_status zs_clahe(zs_context* ctx, const uint8_t* d_src, size_t pitch, size_t stride, int width, int height, int count,
                              double clip_limit, int tiles_x, int tiles_y, uint8_t* d_dst, size_t dst_pitch, size_t dst_stride)
{
    ZS_REQUIRE(ctx && d_src && d_dst, "null argument");
    ZS_REQUIRE(width > 0 && height > 0 && count >= 0 && tiles_x > 0 && tiles_y > 0 && tiles_x <= 64 && tiles_y <= 64, "bad geometry");
    if (count == 0) return ZS_OK;
    clahe_args a;
    a.src = d_src; a.spitch = pitch; a.sstride = stride; a.dst = d_dst; a.dpitch = dst_pitch; a.dstride = dst_stride;
    a.w = width; a.h = height; a.tiles_x = tiles_x; a.tiles_y = tiles_y;
    int ew = width, eh = height;
    if (width % tiles_x != 0 || height % tiles_y != 0) { ew = width + tiles_x - width % tiles_x; eh = height + tiles_y - height % tiles_y; }
    a.tw = ew / tiles_x; a.th = eh / tiles_y;
    const int total = a.tw * a.th;
    a.lut_scale = (float)255 / total;                                   // static_cast<float>(histSize - 1) / tileSizeTotal
    a.clip = 0;
    if (clip_limit > 0.0) { a.clip = (int)(clip_limit * total / 256); if (a.clip < 1) a.clip = 1; }
    a.inv_tw = 1.0f / a.tw; a.inv_th = 1.0f / a.th;
    void* s;
    zs_status st = zs_scratch(ctx, (size_t)count * tiles_x * tiles_y * 256, &s);
    if (st != ZS_OK) return st;
    a.lut = (uint8_t*)s;
    k_clahe_lut<<<dim3(tiles_x * tiles_y, count), 256, 0, ctx->stream>>>(a);
    ZS_LAUNCH_CHECK(ctx);
    k_clahe_apply<<<dim3(zs_div_up(width, 256), height, count), 256, 0, ctx->stream>>>(a);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// ---- remap (INTER_LINEAR, BORDER_CONSTANT 0, CV_32FC1 maps) -------------------------------------------
// cvRound(v * 32) as x86 cvtss2si does it: NaN and out-of-range values become INT_MIN
__device__ __forceinline__ int cv_round_q5(float v)
{
    const float f = __fmul_rn(v, 32.f);
    if (!(f >= -2147483648.f && f < 2147483648.f)) return (int)0x80000000;
    return __float2int_rn(f);
}

__global__ void __launch_bounds__(256) k_remap_linear(const uint8_t* __restrict__ src, size_t spitch, size_t sstride, int sw, int sh,
                                                      const float* __restrict__ mx, const float* __restrict__ my, size_t mpitch,
                                                      size_t mstride, int dw, int dh, uint8_t* __restrict__ dst, size_t dpitch,
                                                      size_t dstride)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, img = blockIdx.z;
    if (x0 >= dw) return;
    const uint8_t* s = src + (size_t)img * sstride;
    const float* px = mx + (size_t)img * mstride + (size_t)y * mpitch + x0;
    const float* py = my + (size_t)img * mstride + (size_t)y * mpitch + x0;
    uint8_t out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        out[k] = 0;
        if (x0 + k >= dw) continue;
        const int sx = cv_round_q5(px[k]), sy = cv_round_q5(py[k]);
        int ix = sx >> 5, iy = sy >> 5;
        const int ax = sx & 31, ay = sy & 31;
        ix = min(max(ix, -32768), 32767); iy = min(max(iy, -32768), 32767);       // saturate_cast<short>
        int acc = 1 << 14;
        const bool x0in = ix >= 0 && ix < sw, x1in = ix + 1 >= 0 && ix + 1 < sw, y0in = iy >= 0 && iy < sh, y1in = iy + 1 >= 0 && iy + 1 < sh;
        if (y0in) {
            const uint8_t* r = s + (size_t)iy * spitch;
            if (x0in) acc += (32 - ax) * (32 - ay) * 32 * (int)r[ix];
            if (x1in) acc += ax * (32 - ay) * 32 * (int)r[ix + 1];
        }
        if (y1in) {
            const uint8_t* r = s + (size_t)(iy + 1) * spitch;
            if (x0in) acc += (32 - ax) * ay * 32 * (int)r[ix];
            if (x1in) acc += ax * ay * 32 * (int)r[ix + 1];
        }
        out[k] = (uint8_t)(acc >> 15);
    }
    uint8_t* d = dst + (size_t)img * dstride + (size_t)y * dpitch + x0;
    if (x0 + 4 <= dw && (((uintptr_t)d) & 3) == 0) *(uint32_t*)d = (uint32_t)out[0] | ((uint32_t)out[1] << 8) | ((uint32_t)out[2] << 16) | ((uint32_t)out[3] << 24);
    else for (int k = 0; k < 4 && x0 + k < dw; ++k) d[k] = out[k];
}

extern "C" zs_status zs_remap_linear(zs_context* ctx, const uint8_t* d_src, size_t pitch, size_t stride, int src_width, int src_height,
                                     int count, const float* d_map_x, const float* d_map_y, size_t map_pitch, size_t map_stride,
                                     int dst_width, int dst_height, uint8_t* d_dst, size_t dst_pitch, size_t dst_stride)
{
    ZS_REQUIRE(ctx && d_src && d_map_x && d_map_y && d_dst, "null argument");
    ZS_REQUIRE(src_width > 0 && src_height > 0 && dst_width > 0 && dst_height > 0 && count >= 0, "bad geometry");
    ZS_REQUIRE(src_width < 32767 && src_height < 32767, "cv::remap addresses sources through 16-bit coordinates");
    if (count == 0) return ZS_OK;
    k_remap_linear<<<dim3(zs_div_up(zs_div_up(dst_width, 4), 256), dst_height, count), 256, 0, ctx->stream>>>(
        d_src, pitch, stride, src_width, src_height, d_map_x, d_map_y, map_pitch, map_stride, dst_width, dst_height, d_dst, dst_pitch, dst_stride);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

// level-0 interior of a pyramid slot, so that zs_remap_linear / zs_clahe / zs_cvt_bgr2gray can produce the frame in place
extern "C" zs_status zs_pyramid_level0(const zs_pyramid* p, int slot, uint8_t** d_ptr, size_t* pitch, size_t* slot_stride)
{
    ZS_REQUIRE(p && d_ptr && pitch && slot_stride, "null argument");
    ZS_REQUIRE(slot >= 0 && slot < p->slots, "bad slot");
    const zs_pyr_view& v = p->v;
    *d_ptr = v.img[0] + (size_t)slot * v.slot_stride[0] + (size_t)v.pad_y * v.pitch[0] + v.pad_x;
    *pitch = (size_t)v.pitch[0];
    *slot_stride = v.slot_stride[0];
    return ZS_OK;
}

// ---- host mirror of processor::process's image path for one image (synchronous) ---------------------
extern "C" zs_status zs_process_image_host(zs_context* ctx, const uint8_t* image, int channels, int width, int height, size_t pitch,
                                           int clahe_enabled, double clahe_clip_limit, const float* map_x, const float* map_y,
                                           uint8_t* undistorted)
{
    ZS_REQUIRE(ctx && image && undistorted, "null argument");
    ZS_REQUIRE(channels == 1 || channels == 3, "channels must be 1 (gray) or 3 (BGR)");
    ZS_REQUIRE((map_x == nullptr) == (map_y == nullptr), "map_x and map_y go together");
    ZS_REQUIRE(width > 0 && height > 0 && pitch >= (size_t)width * channels, "bad geometry");
    ZS_CUDA(cudaSetDevice(ctx->device));
    const size_t px = (size_t)width * height;
    const size_t gpitch = ((size_t)width + 15) / 16 * 16;
    const size_t o_in = 0, o_g0 = o_in + (pitch * height + 255) / 256 * 256, o_g1 = o_g0 + (gpitch * height + 255) / 256 * 256,
                 o_mx = o_g1 + (gpitch * height + 255) / 256 * 256, o_my = o_mx + (px * 4 + 255) / 256 * 256, total = o_my + (px * 4 + 255) / 256 * 256;
    zs_async_buffer buf(ctx->stream);
    ZS_CUDA(cudaMallocAsync((void**)&buf.p, total, ctx->stream));
    uint8_t* base = buf.p;
    zs_status st = ZS_OK;
    cudaError_t e = cudaMemcpyAsync(base + o_in, image, pitch * height, cudaMemcpyHostToDevice, ctx->stream);
    uint8_t* cur = base + o_g0; uint8_t* other = base + o_g1;
    if (e == cudaSuccess) {
        if (channels == 3) st = zs_cvt_bgr2gray(ctx, base + o_in, pitch, 0, width, height, 1, cur, gpitch, 0);
        else e = cudaMemcpy2DAsync(cur, gpitch, base + o_in, pitch, width, height, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e == cudaSuccess && st == ZS_OK && clahe_enabled) {
        st = zs_clahe(ctx, cur, gpitch, 0, width, height, 1, clahe_clip_limit, 8, 8, other, gpitch, 0);
        uint8_t* t = cur; cur = other; other = t;
    }
    if (e == cudaSuccess && st == ZS_OK && map_x) {
        e = cudaMemcpyAsync(base + o_mx, map_x, px * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_my, map_y, px * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess)
            st = zs_remap_linear(ctx, cur, gpitch, 0, width, height, 1, (const float*)(base + o_mx), (const float*)(base + o_my), width, 0,
                                 width, height, other, gpitch, 0);
        uint8_t* t = cur; cur = other; other = t;
    }
    if (e == cudaSuccess && st == ZS_OK)
        e = cudaMemcpy2DAsync(undistorted, width, cur, gpitch, width, height, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return zs_cuda_fail(e, "zs_process_image_host", __FILE__, __LINE__);
    return st;
}
