// zs_context.cu -- context, error reporting, scratch, pyramid storage (allocation / upload / download).
#include <stdarg.h>
#include <stdlib.h>

#include <cuda.h>

#include "zs_common.cuh"

// cuTensorMapEncodeTiled is resolved through the runtime so that the library does not link libcuda
zs_encode_tiled_fn zs_get_encode_tiled()
{
    static zs_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (zs_encode_tiled_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// TMA descriptors of a pyramid's planes (used by the KLT kernel to stage 32x32 patches)
static zs_status make_tensor_maps(zs_pyramid* p)
{
    zs_pyr_view& v = p->v;
    v.tmaps = nullptr; v.fast_maps = nullptr; p->tmaps_dev = nullptr;
    zs_encode_tiled_fn enc = zs_get_encode_tiled();
    if (!enc) { zs_set_error("cuTensorMapEncodeTiled is not available from this driver"); return ZS_ERR_CUDA; }
    CUtensorMap maps[2 * ZS_MAX_LEVELS + 3];
    for (int l = 0; l < v.levels; ++l) {
        const cuuint64_t rows = (cuuint64_t)(v.h[l] + 2 * v.pad_y);
        const cuuint64_t dims[3] = { (cuuint64_t)v.pitch[l], rows, (cuuint64_t)v.slots };
        // TMA needs a 16-byte aligned box origin, so the boxes are wider than the 33 columns a 32-wide window tile needs:
        // 48 bytes cover any byte offset 0..15, 36 (dx,dy) words cover any word offset 0..3; 33 rows = 32 + the bilinear row
        const cuuint32_t box[3] = { 48, 33, 1 }, boxd[3] = { 36, 33, 1 }, es[3] = { 1, 1, 1 };
        const cuuint64_t st8[2] = { (cuuint64_t)v.pitch[l], (cuuint64_t)v.slot_stride[l] };
        CUresult r = enc(&maps[2 * l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, v.img[l], dims, st8, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { zs_set_error("cuTensorMapEncodeTiled(image level %d) failed: %d", l, (int)r); return ZS_ERR_CUDA; }
        const cuuint64_t st32[2] = { (cuuint64_t)v.pitch[l] * 4, (cuuint64_t)v.slot_stride[l] * 4 };
        r = enc(&maps[2 * l + 1], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, v.der[l], dims, st32, boxd, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { zs_set_error("cuTensorMapEncodeTiled(deriv level %d) failed: %d", l, (int)r); return ZS_ERR_CUDA; }
    }
    // grid-FAST strips of level 0 (zs_fast.cu, k_fast_grid_v2): the same plane as u32 elements, one box = one block's strip
    // of cells plus 16 bytes of slack; kept behind the per-level maps
    bool fast_ok = (v.pitch[0] % 16) == 0 && (v.slot_stride[0] % 16) == 0;
    {
        const cuuint64_t rows = (cuuint64_t)(v.h[0] + 2 * v.pad_y);
        const cuuint64_t dims[3] = { (cuuint64_t)v.pitch[0] / 4, rows, (cuuint64_t)v.slots };
        const cuuint64_t st[2] = { (cuuint64_t)v.pitch[0], (cuuint64_t)v.slot_stride[0] };
        const cuuint32_t es[3] = { 1, 1, 1 };
        const cuuint32_t boxes[3][3] = { { (16 * 16 + 16) / 4, 16, 1 }, { (4 * 32 + 16) / 4, 32, 1 }, { (64 + 16) / 4, 64, 1 } };
        for (int k = 0; k < 3 && fast_ok; ++k) {
            const CUresult r = enc(&maps[2 * ZS_MAX_LEVELS + k], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, v.img[0], dims, st, boxes[k], es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) fast_ok = false;       // e.g. an image lower than a 64-row box: the kernels keep their vector loads
        }
    }
    ZS_CUDA(cudaMalloc(&p->tmaps_dev, sizeof(CUtensorMap) * (2 * ZS_MAX_LEVELS + 3)));
    ZS_CUDA(cudaMemcpyAsync(p->tmaps_dev, maps, sizeof(CUtensorMap) * 2 * v.levels, cudaMemcpyHostToDevice, p->ctx->stream));
    if (fast_ok)
        ZS_CUDA(cudaMemcpyAsync((CUtensorMap*)p->tmaps_dev + 2 * ZS_MAX_LEVELS, maps + 2 * ZS_MAX_LEVELS, sizeof(CUtensorMap) * 3,
                                cudaMemcpyHostToDevice, p->ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(p->ctx->stream));        // `maps` is on this stack frame
    v.tmaps = p->tmaps_dev;
    v.fast_maps = fast_ok ? (const void*)((CUtensorMap*)p->tmaps_dev + 2 * ZS_MAX_LEVELS) : nullptr;
    return ZS_OK;
}

void zs_read_switches(zs_switches* s)
{
    auto on = [](const char* n) { return getenv(n) != nullptr; };
    auto num = [](const char* n) { const char* e = getenv(n); return e ? atoi(e) : 0; };
    s->fe_no_graph = on("ZS_FE_NO_GRAPH"); s->klt_no_tma = on("ZS_KLT_NO_TMA"); s->klt_no_share = on("ZS_KLT_NO_SHARE");
    s->lk_no_cache = on("ZS_LK_NO_CACHE"); s->fast_v1 = on("ZS_FAST_V1"); s->subpix_v1 = on("ZS_SUBPIX_V1"); s->fast_no_tma = on("ZS_FAST_NO_TMA"); s->klt63_four_warps = on("ZS_KLT63_FOUR_WARPS"); s->klt63_unpacked = on("ZS_KLT63_UNPACKED"); s->klt_persist_min = num("ZS_KLT_PERSIST_MIN"); s->klt63_packed = num("ZS_KLT63_PACKED"); s->klt31_packed = num("ZS_KLT31_PACKED"); s->klt_no_persist = on("ZS_KLT_NO_PERSIST"); s->l2_no_tensor = on("ZS_L2_NO_TENSOR");
    s->l2_one_tile = on("ZS_L2_ONE_TILE"); s->l2_chains = on("ZS_L2_CHAINS"); s->hamming_no_tensor = on("ZS_HAMMING_NO_TENSOR"); s->hamming_tensor_min = num("ZS_HAMMING_TENSOR_MIN"); s->fast_pretest = on("ZS_FAST_PRETEST");
    s->pyr_force = on("ZS_PYR_SPLIT") ? 1 : on("ZS_PYR_FUSED") ? 2 : 0;
    s->hamming_splits = num("ZS_HAMMING_SPLITS"); s->hamming_variant = num("ZS_HAMMING_VARIANT");
    s->l2_splits = num("ZS_L2_SPLITS"); s->l2_epi_groups = num("ZS_L2_EPI_GROUPS");
}

static thread_local char g_err[512] = "";

void zs_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

zs_status zs_cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    zs_set_error("%s:%d: %s -> %s", file, line, what, cudaGetErrorString(e));
    return ZS_ERR_CUDA;
}

extern "C" {

const char* zs_version(void) { return "zenslam_cuda 0.1 (sm_100a)"; }

const char* zs_last_error_string(void) { return g_err; }

const char* zs_status_string(zs_status s)
{
    switch (s) {
    case ZS_OK: return "ok";
    case ZS_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required)";
    case ZS_ERR_INVALID: return "invalid argument";
    case ZS_ERR_CUDA: return "CUDA runtime error";
    case ZS_ERR_CAPACITY: return "output capacity too small";
    case ZS_ERR_UNSUPPORTED: return "unsupported input";
    default: return "unknown status";
    }
}

// cf. zenslam::metal::is_available (zenslam_metal/source/pyr_lk.cpp:23-30): true only when the
// kernels in this library can actually run (they are compiled for sm_100a and nothing else).
int zs_is_available(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) return 1;
    }
    return 0;
}

zs_status zs_context_create(int device, void* stream, zs_context** out)
{
    ZS_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        zs_set_error("no CUDA device visible; this backend has no CPU fallback");
        return ZS_ERR_NO_DEVICE;
    }
    ZS_REQUIRE(device >= 0 && device < n, "device ordinal out of range");
    cudaDeviceProp prop;
    ZS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        zs_set_error("device %d is sm_%d%d; libzenslam_cuda is built for sm_100a only", device, prop.major, prop.minor);
        return ZS_ERR_NO_DEVICE;
    }
    ZS_CUDA(cudaSetDevice(device));
    zs_context* c = (zs_context*)calloc(1, sizeof(zs_context));
    c->device = device;
    zs_read_switches(&c->sw);
    c->sm_count = prop.multiProcessorCount;
    if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { free(c); return zs_cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
        c->own_stream = true;
    }
    {
        cudaError_t e = cudaMalloc((void**)&c->d_async_err, 256);
        if (e == cudaSuccess) e = cudaMemset(c->d_async_err, 0, 256);
        if (e != cudaSuccess) {
            if (c->own_stream) cudaStreamDestroy(c->stream);
            free(c);
            return zs_cuda_fail(e, "cudaMalloc(async error flags)", __FILE__, __LINE__);
        }
        c->d_klt_work = c->d_async_err + 32;
    }
    *out = c;
    return ZS_OK;
}

// Errors that kernels of stream-asynchronous entries found in their DATA (not their arguments): waits for the stream, reports
// and clears them.  [0] / [1]: zs_match_l2_* got descriptors that are not integers in 0..255 ([0] belongs to the call in flight
// and blanks its results, [1] stays until it is reported here: a later valid call is not affected by an earlier bad one).
zs_status zs_context_async_error(zs_context* c)
{
    ZS_REQUIRE(c, "ctx is null");
    ZS_CUDA(cudaSetDevice(c->device));
    int flags[4] = { 0, 0, 0, 0 };
    ZS_CUDA(cudaMemcpyAsync(flags, c->d_async_err, sizeof(flags), cudaMemcpyDeviceToHost, c->stream));
    ZS_CUDA(cudaStreamSynchronize(c->stream));
    if (flags[1]) {                                          // [1] is the sticky copy of the per-call flag [0]
        ZS_CUDA(cudaMemsetAsync(c->d_async_err, 0, sizeof(flags), c->stream));
        zs_set_error("L2 matching is exact only for integer-valued descriptors in 0..255 (cv::SIFT); got other values "
                     "(every match of that call was reported as -1)");
        return ZS_ERR_UNSUPPORTED;
    }
    return ZS_OK;
}

void zs_context_destroy(zs_context* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 2; ++i) if (c->host_pyr[i]) zs_pyramid_destroy(c->host_pyr[i]);
    for (int i = 0; i < ZS_LK_CACHE_SLOTS; ++i) free(c->lk_copy[i]);
    for (int i = 0; i < 2; ++i) if (c->frame_pin[i]) cudaFreeHost(c->frame_pin[i]);
    if (c->host_orb) zs_orb_detector_destroy(c->host_orb);
    if (c->d_async_err) cudaFree(c->d_async_err);
    if (c->scratch) cudaFree(c->scratch);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    free(c);
}

zs_status zs_context_synchronize(zs_context* c)
{
    ZS_REQUIRE(c, "ctx is null");
    ZS_CUDA(cudaStreamSynchronize(c->stream));
    return ZS_OK;
}

zs_status zs_context_reload_switches(zs_context* c)
{
    ZS_REQUIRE(c, "ctx is null");
    zs_read_switches(&c->sw);
    return ZS_OK;
}

void* zs_context_stream(zs_context* c) { return c ? (void*)c->stream : nullptr; }
uint64_t zs_context_launch_count(const zs_context* c) { return c ? c->launches : 0; }

}  // extern "C"

zs_status zs_scratch(zs_context* ctx, size_t bytes, void** out)
{
    if (bytes > ctx->scratch_bytes) {
        // single-stream ordering: everything queued before still uses the old block, so free asynchronously
        if (ctx->scratch) ZS_CUDA(cudaFreeAsync(ctx->scratch, ctx->stream));
        ctx->scratch = nullptr; ctx->scratch_bytes = 0;
        size_t want = bytes + bytes / 4 + 4096;
        ZS_CUDA(cudaMallocAsync(&ctx->scratch, want, ctx->stream));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return ZS_OK;
}

zs_status zs_pinned(zs_context* ctx, size_t bytes, void** out)
{
    if (bytes > ctx->pinned_bytes) {
        ZS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->pinned) ZS_CUDA(cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr; ctx->pinned_bytes = 0;
        size_t want = bytes + bytes / 4 + 4096;
        ZS_CUDA(cudaMallocHost(&ctx->pinned, want));
        ctx->pinned_bytes = want;
    }
    *out = ctx->pinned;
    return ZS_OK;
}

// ------------------------------------------------------------------------------------------------
// pyramid storage
// ------------------------------------------------------------------------------------------------
static int num_levels(int w, int h, int win_w, int win_h, int max_level)
{
    // cv::buildOpticalFlowPyramid: stop before a level with width <= win.width or height <= win.height
    int levels = 1;
    for (int l = 0; l < max_level; ++l) {
        w = (w + 1) / 2; h = (h + 1) / 2;
        if (w <= win_w || h <= win_h) break;
        ++levels;
    }
    return levels;
}

extern "C" {

zs_status zs_pyramid_create(zs_context* ctx, int width, int height, int slots, int win_w, int win_h,
                            int max_level, zs_pyramid** out)
{
    ZS_REQUIRE(ctx && out, "null argument");
    ZS_REQUIRE(width > 0 && height > 0 && slots > 0, "bad geometry");
    ZS_REQUIRE(win_w >= 3 && win_h >= 3 && win_w <= 127 && win_h <= 127, "window must be within 3..127");
    ZS_REQUIRE(max_level >= 0, "max_level < 0");
    if (max_level > ZS_MAX_LEVELS - 1) max_level = ZS_MAX_LEVELS - 1;
    ZS_CUDA(cudaSetDevice(ctx->device));
    zs_pyramid* p = (zs_pyramid*)calloc(1, sizeof(zs_pyramid));
    p->ctx = ctx; p->width = width; p->height = height; p->slots = slots;
    p->win_w = win_w; p->win_h = win_h; p->max_level = max_level;
    zs_pyr_view& v = p->v;
    v.levels = num_levels(width, height, win_w, win_h, max_level);
    v.slots = slots;
    v.pad_x = (win_w + 15) / 16 * 16;
    v.pad_y = win_h;
    size_t total = 0, img_off[ZS_MAX_LEVELS], der_off[ZS_MAX_LEVELS];
    int w = width, h = height;
    for (int l = 0; l < v.levels; ++l) {
        v.w[l] = w; v.h[l] = h;
        v.pitch[l] = (w + 2 * v.pad_x + 127) / 128 * 128;
        v.slot_stride[l] = (size_t)v.pitch[l] * (h + 2 * v.pad_y);
        img_off[l] = total; total += v.slot_stride[l] * slots;
        total = (total + 255) / 256 * 256;
        der_off[l] = total; total += v.slot_stride[l] * slots * sizeof(short2);
        total = (total + 255) / 256 * 256;
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
    v.blur_pitch = (width + 127) / 128 * 128;
    v.blur_slot = (size_t)v.blur_pitch * height;
    size_t blur_off = total; total += v.blur_slot * slots;
    cudaError_t e = cudaMalloc(&p->block, total);
    if (e != cudaSuccess) { free(p); return zs_cuda_fail(e, "cudaMalloc(pyramid)", __FILE__, __LINE__); }
    p->block_bytes = total;
    // derivative padding must be zero (OpenCV pads the Scharr planes with BORDER_CONSTANT 0); the kernels
    // only ever write plane interiors, so one clear at creation suffices.
    e = cudaMemsetAsync(p->block, 0, total, ctx->stream);
    if (e != cudaSuccess) { cudaFree(p->block); free(p); return zs_cuda_fail(e, "cudaMemset(pyramid)", __FILE__, __LINE__); }
    for (int l = 0; l < v.levels; ++l) {
        v.img[l] = (uint8_t*)p->block + img_off[l];
        v.der[l] = (short2*)((uint8_t*)p->block + der_off[l]);
    }
    v.blur = (uint8_t*)p->block + blur_off;
    zs_status st = make_tensor_maps(p);
    if (st != ZS_OK) { cudaFree(p->block); free(p); return st; }
    *out = p;
    return ZS_OK;
}

void zs_pyramid_destroy(zs_pyramid* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->block);
    if (p->tmaps_dev) cudaFree(p->tmaps_dev);
    free(p);
}

int zs_pyramid_levels(const zs_pyramid* p) { return p ? p->v.levels : 0; }

zs_status zs_pyramid_level_size(const zs_pyramid* p, int level, int* width, int* height)
{
    ZS_REQUIRE(p && level >= 0 && level < p->v.levels, "bad level");
    if (width) *width = p->v.w[level];
    if (height) *height = p->v.h[level];
    return ZS_OK;
}

zs_status zs_pyramid_upload(zs_context* ctx, zs_pyramid* p, const uint8_t* src, size_t pitch, size_t stride,
                            int first, int count, int src_is_host)
{
    ZS_REQUIRE(ctx && p && src, "null argument");
    ZS_REQUIRE(count >= 0 && count <= p->slots && first >= 0, "bad slot range");
    ZS_REQUIRE(pitch >= (size_t)p->width, "pitch < width");
    const zs_pyr_view& v = p->v;
    const cudaMemcpyKind kind = src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    // slots first.. are contiguous in memory except at the ring wrap: at most two strided 2-D copies
    int done = 0;
    while (done < count) {
        const int s = (first + done) % p->slots;
        const int run = (count - done < p->slots - s) ? count - done : p->slots - s;
        uint8_t* dst = v.img[0] + (size_t)s * v.slot_stride[0] + (size_t)v.pad_y * v.pitch[0] + v.pad_x;
        cudaMemcpy3DParms c;
        memset(&c, 0, sizeof(c));
        c.srcPtr = make_cudaPitchedPtr((void*)(src + (size_t)done * stride), pitch, pitch, stride / pitch);
        c.dstPtr = make_cudaPitchedPtr((void*)dst, v.pitch[0], v.pitch[0], p->height + 2 * v.pad_y);
        c.extent = make_cudaExtent(p->width, p->height, run);
        c.kind = kind;
        if (stride % pitch != 0) {
            // slot stride not a whole number of rows: fall back to one 2-D copy per image
            for (int i = 0; i < run; ++i)
                ZS_CUDA(cudaMemcpy2DAsync(dst + (size_t)i * v.slot_stride[0], v.pitch[0],
                                          src + (size_t)(done + i) * stride, pitch, p->width, p->height, kind,
                                          ctx->stream));
        } else {
            ZS_CUDA(cudaMemcpy3DAsync(&c, ctx->stream));
        }
        done += run;
    }
    return ZS_OK;
}

zs_status zs_pyramid_download_image(zs_context* ctx, const zs_pyramid* p, int slot, int level, uint8_t* dst)
{
    ZS_REQUIRE(ctx && p && dst, "null argument");
    ZS_REQUIRE(slot >= 0 && slot < p->slots && level >= 0 && level < p->v.levels, "bad slot/level");
    const zs_pyr_view& v = p->v;
    const uint8_t* src = v.img[level] + (size_t)slot * v.slot_stride[level] + (size_t)v.pad_y * v.pitch[level] + v.pad_x;
    ZS_CUDA(cudaMemcpy2DAsync(dst, v.w[level], src, v.pitch[level], v.w[level], v.h[level], cudaMemcpyDeviceToHost,
                              ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZS_OK;
}

zs_status zs_pyramid_download_deriv(zs_context* ctx, const zs_pyramid* p, int slot, int level, int16_t* dst)
{
    ZS_REQUIRE(ctx && p && dst, "null argument");
    ZS_REQUIRE(slot >= 0 && slot < p->slots && level >= 0 && level < p->v.levels, "bad slot/level");
    const zs_pyr_view& v = p->v;
    const short2* src = v.der[level] + (size_t)slot * v.slot_stride[level] + (size_t)v.pad_y * v.pitch[level] + v.pad_x;
    ZS_CUDA(cudaMemcpy2DAsync(dst, (size_t)v.w[level] * 4, src, (size_t)v.pitch[level] * 4, (size_t)v.w[level] * 4,
                              v.h[level], cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZS_OK;
}

zs_status zs_orb_download_blur(zs_context* ctx, const zs_pyramid* p, int slot, uint8_t* dst)
{
    ZS_REQUIRE(ctx && p && dst, "null argument");
    ZS_REQUIRE(slot >= 0 && slot < p->slots, "bad slot");
    const zs_pyr_view& v = p->v;
    ZS_CUDA(cudaMemcpy2DAsync(dst, p->width, v.blur + (size_t)slot * v.blur_slot, v.blur_pitch, p->width, p->height,
                              cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZS_OK;
}

}  // extern "C"
