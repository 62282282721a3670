// zs_host.cu -- host-pointer mirrors of the reference's calls (single image, synchronous): what the C++
// adapter in zenslam_cuda/ binds for drop-in use.  They stage through the context's device scratch, run the
// same kernels as the batched path with batch 1, and copy the results back.
#include <stdlib.h>

#include "zs_common.cuh"
#if defined(__x86_64__)
#include <nmmintrin.h>
#endif

static zs_status host_pyramid(zs_context* ctx, int which, int w, int h, int slots, int win_w, int win_h, int max_level,
                              zs_pyramid** out)
{
    zs_pyramid* p = ctx->host_pyr[which];
    if (p && (p->width != w || p->height != h || p->slots != slots || p->win_w != win_w || p->win_h != win_h ||
              p->max_level != max_level)) {
        zs_pyramid_destroy(p);
        p = ctx->host_pyr[which] = nullptr;
    }
    if (!p) {
        zs_status st = zs_pyramid_create(ctx, w, h, slots, win_w, win_h, max_level, &p);
        if (st != ZS_OK) return st;
        ctx->host_pyr[which] = p;
    }
    *out = p;
    return ZS_OK;
}

static inline size_t al256(size_t v) { return (v + 255) / 256 * 256; }

// One pageable host frame into level 0 of a pyramid slot, for the detector entries (whose kernels wait for the frame): the
// frame goes through one of two pinned buffers of the context in four row chunks, each chunk's DMA running while the host
// copies the next one, and the call returns with the last DMA still in flight -- 25 us less per 752 x 480 detection call than
// cudaMemcpy2DAsync from the pageable buffer (tools/bench_seams.py: 0.355 -> 0.304 ms for the two calls of a stereo frame).
// Every host entry ends with a stream synchronisation, so a buffer is never reused while a DMA reads it.
static zs_status upload_frame(zs_context* ctx, zs_pyramid* p, const uint8_t* img, size_t pitch, int slot)
{
    const int w = p->width, h = p->height;
    const int b = ctx->frame_pin_next; ctx->frame_pin_next ^= 1;
    const size_t bytes = (size_t)w * h;
    if (ctx->frame_pin_bytes[b] < bytes) {
        ZS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->frame_pin[b]) ZS_CUDA(cudaFreeHost(ctx->frame_pin[b]));
        ctx->frame_pin[b] = nullptr; ctx->frame_pin_bytes[b] = 0;
        ZS_CUDA(cudaMallocHost((void**)&ctx->frame_pin[b], bytes));
        ctx->frame_pin_bytes[b] = bytes;
    }
    uint8_t* pin = ctx->frame_pin[b];
    uint8_t* dst; size_t dpitch, dstride;
    zs_status st = zs_pyramid_level0(p, slot, &dst, &dpitch, &dstride);
    if (st != ZS_OK) return st;
    const int chunks = h >= 64 ? 4 : 1;
    for (int c = 0; c < chunks; ++c) {
        const int y0 = (int)((long long)h * c / chunks), y1 = (int)((long long)h * (c + 1) / chunks);
        if (pitch == (size_t)w) memcpy(pin + (size_t)y0 * w, img + (size_t)y0 * pitch, (size_t)(y1 - y0) * w);
        else for (int y = y0; y < y1; ++y) memcpy(pin + (size_t)y * w, img + (size_t)y * pitch, w);
        ZS_CUDA(cudaMemcpy2DAsync(dst + (size_t)y0 * dpitch, dpitch, pin + (size_t)y0 * w, w, w, y1 - y0, cudaMemcpyHostToDevice,
                                  ctx->stream));
    }
    return ZS_OK;
}


// 64-bit content hash of a frame.  It keys the device-side pyramid cache below: a frame whose bytes were seen before keeps
// its slot, so its upload and pyramid build are skipped.  The key is the content, not the pointer -- buffers are recycled
// between frames.  Every LK call hashes both of its frames in full, so the hash is on the critical path of the per-frame
// seam calls: four interleaved CRC32C lanes (one `crc32` per 8 bytes, 8 bytes per cycle: the speed of reading the frame,
// ~25 us for 752 x 480 from DRAM) where the CPU has SSE4.2, four multiply-xorshift lanes (~50 us) elsewhere.
static uint64_t frame_hash_scalar(const uint8_t* img, int w, int h, size_t pitch)
{
    const uint64_t K = 0x9E3779B97F4A7C15ull;
    uint64_t a = K ^ (uint64_t)w, b = K * 3 ^ (uint64_t)h, c = K * 5, d = K * 7;
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = img + (size_t)y * pitch;
        int x = 0;
        for (; x + 32 <= w; x += 32) {
            uint64_t v[4];
            memcpy(v, r + x, 32);
            a = (a ^ v[0]) * K; a ^= a >> 29;
            b = (b ^ v[1]) * K; b ^= b >> 29;
            c = (c ^ v[2]) * K; c ^= c >> 29;
            d = (d ^ v[3]) * K; d ^= d >> 29;
        }
        uint64_t t = 0;
        for (; x < w; ++x) t = t * 257 + r[x];
        a = (a ^ t ^ (uint64_t)y) * K; a ^= a >> 29;
    }
    uint64_t hsh = a ^ (b * K) ^ ((c * K) >> 7) ^ (d << 3);
    hsh ^= hsh >> 31; hsh *= K; hsh ^= hsh >> 29;
    return hsh ? hsh : 1;                      // 0 marks an empty slot
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) static uint64_t frame_hash_crc(const uint8_t* img, int w, int h, size_t pitch)
{
    const uint64_t K = 0x9E3779B97F4A7C15ull;
    uint64_t a = (uint32_t)w, b = (uint32_t)h, c = 0x243F6A88u, d = 0x85A308D3u;
    for (int y = 0; y < h; ++y) {
        const uint8_t* r = img + (size_t)y * pitch;
        int x = 0;
        for (; x + 32 <= w; x += 32) {
            uint64_t v[4];
            memcpy(v, r + x, 32);
            a = _mm_crc32_u64(a, v[0]); b = _mm_crc32_u64(b, v[1]); c = _mm_crc32_u64(c, v[2]); d = _mm_crc32_u64(d, v[3]);
        }
        for (; x < w; ++x) a = _mm_crc32_u8((uint32_t)a, r[x]);
        d = _mm_crc32_u32((uint32_t)d, (uint32_t)y);      // the row number: equal rows in a different order differ
    }
    uint64_t hsh = ((a | (b << 32)) * K) ^ (c | (d << 32));
    hsh ^= hsh >> 31; hsh *= K; hsh ^= hsh >> 29;
    return hsh ? hsh : 1;
}
#endif

static uint64_t frame_hash(const uint8_t* img, int w, int h, size_t pitch)
{
#if defined(__x86_64__)
    static const bool crc = __builtin_cpu_supports("sse4.2");
    if (crc) return frame_hash_crc(img, w, h, pitch);
#endif
    return frame_hash_scalar(img, w, h, pitch);
}

// A hash match is only a candidate (a collision of the 64-bit hash would silently track against a stale pyramid):
// it is confirmed against retained bytes of the slot's frame -- every LK_SAMPLE_STEP-th row and the last one, an eighth of
// the frame.  (Comparing and retaining the whole frame costs 50 us per LK call at 752x480, a quarter of the call; a false
// hit now needs a 64-bit collision over all bytes AND equality of 45 KB of them.)
#define LK_SAMPLE_STEP 8
static int lk_sample_rows(int h) { return (h + LK_SAMPLE_STEP - 1) / LK_SAMPLE_STEP + 1; }
static int lk_sample_row(int k, int h) { const int y = k * LK_SAMPLE_STEP; return y < h ? y : h - 1; }
static bool frame_equals(const uint8_t* copy, const uint8_t* img, int w, int h, size_t pitch)
{
    if (!copy) return false;
    for (int k = 0; k < lk_sample_rows(h); ++k)
        if (memcmp(copy + (size_t)k * w, img + (size_t)lk_sample_row(k, h) * pitch, w) != 0) return false;
    return true;
}

// the LK pyramid of the host mirrors: ZS_LK_CACHE_SLOTS slots, least-recently-used replacement
static zs_status lk_pyramid(zs_context* ctx, int w, int h, const zs_lk_params* prm, zs_pyramid** out)
{
    zs_pyramid* before = ctx->host_pyr[0];
    zs_status st = host_pyramid(ctx, 0, w, h, ZS_LK_CACHE_SLOTS, prm->win_w, prm->win_h, prm->max_level, out);
    if (st == ZS_OK && *out != before) {       // new geometry: nothing cached yet
        memset(ctx->lk_hash, 0, sizeof(ctx->lk_hash));
        memset(ctx->lk_stamp, 0, sizeof(ctx->lk_stamp));
    }
    return st;
}

static zs_status lk_slot(zs_context* ctx, zs_pyramid* p, const uint8_t* img, int w, int h, size_t pitch, int avoid, int* slot)
{
    const bool no_cache = ctx->sw.lk_no_cache;
    const uint64_t hsh = frame_hash(img, w, h, pitch);
    int lru = -1;
    for (int s = 0; s < ZS_LK_CACHE_SLOTS; ++s) {
        if (s == avoid) continue;
        if (!no_cache && ctx->lk_hash[s] == hsh && frame_equals(ctx->lk_copy[s], img, w, h, pitch)) {
            ctx->lk_stamp[s] = ++ctx->lk_clock; ctx->lk_hits++;
            *slot = s;
            return ZS_OK;
        }
        if (lru < 0 || ctx->lk_stamp[s] < ctx->lk_stamp[lru]) lru = s;
    }
    ctx->lk_misses++;
    ctx->lk_hash[lru] = 0;                      // not valid until both calls below have been queued
    // (plain pageable copy here: through upload_frame's pinned chunks an LK call with a new frame was 7 us SLOWER -- the
    // driver's own staging of this copy already overlaps the hashing of the call's second frame)
    zs_status st = zs_pyramid_upload(ctx, p, img, pitch, pitch * h, lru, 1, 1);
    if (st != ZS_OK) return st;
    if ((st = zs_pyramid_build(ctx, p, lru, 1)) != ZS_OK) return st;
    if (!no_cache) {                            // retained so that a later hash match can be confirmed (a 64-bit hash alone can collide)
        const size_t bytes = (size_t)w * lk_sample_rows(h);
        if (ctx->lk_copy_bytes[lru] < bytes) {
            free(ctx->lk_copy[lru]);
            ctx->lk_copy[lru] = (uint8_t*)malloc(bytes);
            ctx->lk_copy_bytes[lru] = ctx->lk_copy[lru] ? bytes : 0;
        }
        if (!ctx->lk_copy[lru]) { zs_set_error("out of host memory for the LK frame cache"); return ZS_ERR_CUDA; }
        for (int k = 0; k < lk_sample_rows(h); ++k) memcpy(ctx->lk_copy[lru] + (size_t)k * w, img + (size_t)lk_sample_row(k, h) * pitch, w);
    }
    ctx->lk_hash[lru] = hsh; ctx->lk_stamp[lru] = ++ctx->lk_clock;
    *slot = lru;
    return ZS_OK;
}

extern "C" zs_status zs_calc_optical_flow_pyr_lk_host(zs_context* ctx, const uint8_t* prev_img, const uint8_t* next_img,
                                                      int width, int height, size_t pitch, const float* prev_pts,
                                                      float* next_pts, int n, uint8_t* status, float* err,
                                                      const zs_lk_params* prm)
{
    ZS_REQUIRE(ctx && prev_img && next_img && prm, "null argument");
    ZS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return ZS_OK;
    ZS_REQUIRE(prev_pts && next_pts && status && err, "null argument");
    ZS_CUDA(cudaSetDevice(ctx->device));
    zs_pyramid* p;
    zs_status st = lk_pyramid(ctx, width, height, prm, &p);
    if (st != ZS_OK) return st;
    int slot_prev = 0, slot_next = 1;
    if ((st = lk_slot(ctx, p, prev_img, width, height, pitch, -1, &slot_prev)) != ZS_OK) return st;
    if ((st = lk_slot(ctx, p, next_img, width, height, pitch, slot_prev, &slot_next)) != ZS_OK) return st;
    // device scratch and pinned staging share one layout: [slots(2) count(1)] | prev n*2 f | next n*2 f | err n f | status n.
    // One H2D copy brings the header and the points (the initial flow too, when it is used), one D2H copy returns next / err /
    // status: the seven small pageable copies this call used to make cost as much as its kernel.  The KLT kernel itself uses
    // no context scratch, so one block serves the whole call
    const size_t o_prev = 256, o_next = o_prev + al256(sizeof(float) * 2 * n), o_err = o_next + al256(sizeof(float) * 2 * n),
                 o_st = o_err + al256(sizeof(float) * n), total = o_st + al256(n);
    void *s, *pin;
    if ((st = zs_scratch(ctx, total, &s)) != ZS_OK) return st;
    if ((st = zs_pinned(ctx, total, &pin)) != ZS_OK) return st;
    uint8_t* base = (uint8_t*)s;
    uint8_t* hb = (uint8_t*)pin;
    const bool init = (prm->flags & ZS_LK_USE_INITIAL_FLOW) != 0;
    const int hdr[3] = { slot_prev, slot_next, n };
    memcpy(hb, hdr, sizeof(hdr));
    memcpy(hb + o_prev, prev_pts, sizeof(float) * 2 * n);
    if (init) memcpy(hb + o_next, next_pts, sizeof(float) * 2 * n);
    ZS_CUDA(cudaMemcpyAsync(base, hb, init ? o_err : o_next, cudaMemcpyHostToDevice, ctx->stream));
    st = zs_klt_track(ctx, p, (const int*)base, (const int*)base + 1, (const float*)(base + o_prev), (float*)(base + o_next),
                      (const int*)base + 2, 1, n, prm, base + o_st, (float*)(base + o_err));
    if (st != ZS_OK) return st;
    ZS_CUDA(cudaMemcpyAsync(hb + o_next, base + o_next, total - o_next, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(next_pts, hb + o_next, sizeof(float) * 2 * n);
    memcpy(status, hb + o_st, n);
    memcpy(err, hb + o_err, sizeof(float) * n);
    return ZS_OK;
}

extern "C" zs_status zs_lk_cache_stats(const zs_context* ctx, uint64_t* hits, uint64_t* misses)
{
    ZS_REQUIRE(ctx, "null argument");
    if (hits) *hits = ctx->lk_hits;
    if (misses) *misses = ctx->lk_misses;
    return ZS_OK;
}

// keypoint_tracker::track_keypoints in one call (keypoint_tracker.cpp:129-197 stereo, :343-434 temporal): forward LK
// (initial flow optional), backward LK from the forward results, forward-backward gate.  The reference makes two pyr_lk
// calls for this; through the pyr_lk seam each of them uploads both frames and rebuilds both pyramids, here that happens
// once and both passes + the gate run in one kernel (zs_klt_track_fb).
extern "C" zs_status zs_track_keypoints_host(zs_context* ctx, const uint8_t* img_0, const uint8_t* img_1, int width, int height,
                                             size_t pitch, const float* points_0, const float* predicted_1, int n,
                                             const zs_lk_params* prm, double klt_threshold, float* points_1, uint8_t* status,
                                             float* err, uint8_t* keep)
{
    ZS_REQUIRE(ctx && img_0 && img_1 && prm, "null argument");
    ZS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return ZS_OK;
    ZS_REQUIRE(points_0 && points_1 && keep, "null argument");
    ZS_CUDA(cudaSetDevice(ctx->device));
    zs_pyramid* p;
    zs_status st = lk_pyramid(ctx, width, height, prm, &p);
    if (st != ZS_OK) return st;
    int slot_0 = 0, slot_1 = 1;
    if ((st = lk_slot(ctx, p, img_0, width, height, pitch, -1, &slot_0)) != ZS_OK) return st;
    if ((st = lk_slot(ctx, p, img_1, width, height, pitch, slot_0, &slot_1)) != ZS_OK) return st;
    // one layout for the device scratch and the pinned staging, one copy in, one copy out (see zs_calc_optical_flow_pyr_lk_host)
    const size_t o_prev = 256, o_next = o_prev + al256(sizeof(float) * 2 * n), o_err = o_next + al256(sizeof(float) * 2 * n),
                 o_st = o_err + al256(sizeof(float) * n), o_keep = o_st + al256(n), total = o_keep + al256(n);
    void *s, *pin;
    if ((st = zs_scratch(ctx, total, &s)) != ZS_OK) return st;
    if ((st = zs_pinned(ctx, total, &pin)) != ZS_OK) return st;
    uint8_t* base = (uint8_t*)s;
    uint8_t* hb = (uint8_t*)pin;
    const int hdr[3] = { slot_0, slot_1, n };
    memcpy(hb, hdr, sizeof(hdr));
    memcpy(hb + o_prev, points_0, sizeof(float) * 2 * n);
    zs_lk_params q = *prm;
    if (predicted_1) {
        q.flags |= ZS_LK_USE_INITIAL_FLOW;
        memcpy(hb + o_next, predicted_1, sizeof(float) * 2 * n);
    } else {
        q.flags &= ~ZS_LK_USE_INITIAL_FLOW;
    }
    ZS_CUDA(cudaMemcpyAsync(base, hb, predicted_1 ? o_err : o_next, cudaMemcpyHostToDevice, ctx->stream));
    st = zs_klt_track_fb(ctx, p, (const int*)base, (const int*)base + 1, (const float*)(base + o_prev), (float*)(base + o_next),
                         (const int*)base + 2, 1, n, &q, klt_threshold, base + o_st, (float*)(base + o_err), base + o_keep);
    if (st != ZS_OK) return st;
    ZS_CUDA(cudaMemcpyAsync(hb + o_next, base + o_next, total - o_next, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(points_1, hb + o_next, sizeof(float) * 2 * n);
    if (status) memcpy(status, hb + o_st, n);
    if (err) memcpy(err, hb + o_err, sizeof(float) * n);
    memcpy(keep, hb + o_keep, n);
    return ZS_OK;
}

static zs_status detect_grid_host(zs_context* ctx, const uint8_t* img, int width, int height, size_t pitch, int cell_w, int cell_h,
                                  int threshold, const uint8_t* occupied, float* x, float* y, float* response, uint8_t* desc,
                                  int* n_out, int subpix)
{
    ZS_REQUIRE(ctx && img && x && y && response && desc && n_out, "null argument");
    ZS_REQUIRE(cell_w > 0 && cell_h > 0, "bad cell size");
    ZS_CUDA(cudaSetDevice(ctx->device));
    *n_out = 0;
    const int gw = width / cell_w, gh = height / cell_h, cells = gw * gh;
    if (cells == 0) return ZS_OK;
    zs_pyramid* p;
    zs_status st = host_pyramid(ctx, 1, width, height, 1, 16, 16, 0, &p);
    if (st != ZS_OK) return st;
    if ((st = upload_frame(ctx, p, img, pitch, 0)) != ZS_OK) return st;
    if ((st = zs_pyramid_build(ctx, p, 0, 1)) != ZS_OK) return st;      // fills the reflect padding ORB's blur reads
    // This call's buffers come from cudaMallocAsync rather than the context scratch, because the detection
    // kernels use that scratch themselves.
    // Everything the caller gets back -- both counts, keypoints, responses, descriptors -- is one contiguous tail of the
    // buffer, fetched at full capacity by ONE copy into pinned staging (63 KB at 752 x 480 / 16-px cells) and trimmed to the
    // count on the host: reading the count first and the rows afterwards cost a second round trip and four pageable copies.
    const size_t o_occ = 0, o_xy0 = al256(cells), o_r0 = o_xy0 + al256(sizeof(float) * 2 * cells),
                 o_n0 = o_r0 + al256(sizeof(float) * cells), o_n = o_n0 + 256, o_xy = o_n + 256,
                 o_r = o_xy + al256(sizeof(float) * 2 * cells), o_desc = o_r + al256(sizeof(float) * cells),
                 total = o_desc + al256((size_t)cells * 32);
    void* pin;
    if ((st = zs_pinned(ctx, total - o_n0, &pin)) != ZS_OK) return st;
    zs_async_buffer buf(ctx->stream);
    ZS_CUDA(cudaMallocAsync((void**)&buf.p, total, ctx->stream));
    uint8_t* base = buf.p;
    if (occupied) ZS_CUDA(cudaMemcpyAsync(base + o_occ, occupied, cells, cudaMemcpyHostToDevice, ctx->stream));
    st = zs_fast_grid_detect(ctx, p, 0, 1, cell_w, cell_h, threshold, occupied ? base + o_occ : nullptr, (float*)(base + o_xy0),
                             (float*)(base + o_r0), (int*)(base + o_n0), cells);
    // PARALLEL_GRID: cv::cornerSubPix(win 5x5, 30 iterations, eps 0.01) on every selected corner before ORB::compute
    // (keypoint_detector_parallel.cpp:160-170)
    if (st == ZS_OK && subpix)
        st = zs_corner_subpix(ctx, p, 0, 1, (float*)(base + o_xy0), (const int*)(base + o_n0), cells, 5, 5, 30, 0.01);
    if (st == ZS_OK)
        st = zs_orb_compute(ctx, p, 0, 1, (const float*)(base + o_xy0), (const float*)(base + o_r0), nullptr,
                            (const int*)(base + o_n0), cells, (float*)(base + o_xy), (float*)(base + o_r), nullptr,
                            (int*)(base + o_n), base + o_desc);
    if (st != ZS_OK) return st;
    const uint8_t* tail = (const uint8_t*)pin;                    // host image of the buffer from o_n0 on
    ZS_CUDA(cudaMemcpyAsync(pin, base + o_n0, total - o_n0, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = *(const int*)(tail + (o_n - o_n0)), n_raw = *(const int*)tail;
    if (!subpix && cell_w >= 63 && cell_h >= 63) {
        // GRID only (PARALLEL_GRID has no such step): a free cell where FAST finds nothing goes through _describer->detect
        // (keypoint_detector_grid.cpp:92-95) -- the 8-level ORB detector, which CAN return a keypoint once the cell is wider
        // than twice its 31-px edge threshold.  That fallback is not implemented: fail loudly instead of diverging silently.
        int free_cells = cells;
        if (occupied) for (int i = 0; i < cells; ++i) free_cells -= occupied[i] ? 1 : 0;
        if (n_raw < free_cells) {
            zs_set_error("GRID detector with cells >= 63 px: %d free cell(s) without a FAST corner would take the reference's "
                         "ORB::detect fallback (keypoint_detector_grid.cpp:92-95), which this backend does not implement; use "
                         "PARALLEL_GRID (no fallback in the reference) or cells <= 62 px", free_cells - n_raw);
            return ZS_ERR_UNSUPPORTED;
        }
    }
    if (n > 0) {
        const float* xy = (const float*)(tail + (o_xy - o_n0));
        for (int i = 0; i < n; ++i) { x[i] = xy[2 * i]; y[i] = xy[2 * i + 1]; }
        memcpy(response, tail + (o_r - o_n0), sizeof(float) * n);
        memcpy(desc, tail + (o_desc - o_n0), (size_t)n * 32);
    }
    *n_out = n;
    return ZS_OK;
}

extern "C" zs_status zs_detect_keypoints_grid_host(zs_context* ctx, const uint8_t* img, int width, int height, size_t pitch,
                                                   int cell_w, int cell_h, int threshold, const uint8_t* occupied, float* x,
                                                   float* y, float* response, uint8_t* desc, int* n_out)
{
    return detect_grid_host(ctx, img, width, height, pitch, cell_w, cell_h, threshold, occupied, x, y, response, desc, n_out, 0);
}

extern "C" zs_status zs_detect_keypoints_parallel_host(zs_context* ctx, const uint8_t* img, int width, int height, size_t pitch,
                                                       int cell_w, int cell_h, int threshold, const uint8_t* occupied, float* x,
                                                       float* y, float* response, uint8_t* desc, int* n_out)
{
    return detect_grid_host(ctx, img, width, height, pitch, cell_w, cell_h, threshold, occupied, x, y, response, desc, n_out, 1);
}

// keypoint_detector_simple::detect_keypoints with `feature: FAST` (keypoint_detector_simple.cpp:38-63): full-frame
// cv::FAST + mask, ORB::compute; no cap on the count in the reference, so the caller sizes the outputs (cap)
extern "C" zs_status zs_detect_keypoints_simple_host(zs_context* ctx, const uint8_t* img, int width, int height, size_t pitch,
                                                     const uint8_t* mask, size_t mask_pitch, int threshold, float* x, float* y,
                                                     float* response, uint8_t* desc, int cap, int* n_out)
{
    ZS_REQUIRE(ctx && img && x && y && response && desc && n_out, "null argument");
    ZS_REQUIRE(width > 0 && height > 0 && cap > 0, "bad sizes");
    ZS_CUDA(cudaSetDevice(ctx->device));
    *n_out = 0;
    zs_pyramid* p;
    zs_status st = host_pyramid(ctx, 1, width, height, 1, 16, 16, 0, &p);
    if (st != ZS_OK) return st;
    if ((st = upload_frame(ctx, p, img, pitch, 0)) != ZS_OK) return st;
    if ((st = zs_pyramid_build(ctx, p, 0, 1)) != ZS_OK) return st;
    const size_t plane = al256((size_t)width * height);
    const size_t o_mask = 0, o_xy0 = plane, o_r0 = o_xy0 + al256(sizeof(float) * 2 * cap), o_n0 = o_r0 + al256(sizeof(float) * cap),
                 o_xy = o_n0 + 256, o_r = o_xy + al256(sizeof(float) * 2 * cap), o_n = o_r + al256(sizeof(float) * cap),
                 o_desc = o_n + 256, total = o_desc + al256((size_t)cap * 32);
    zs_async_buffer buf(ctx->stream);
    ZS_CUDA(cudaMallocAsync((void**)&buf.p, total, ctx->stream));
    uint8_t* base = buf.p;
    if (mask)
        ZS_CUDA(cudaMemcpy2DAsync(base + o_mask, width, mask, mask_pitch, width, height, cudaMemcpyHostToDevice, ctx->stream));
    st = zs_fast_detect(ctx, p, 0, 1, threshold, mask ? base + o_mask : nullptr, (float*)(base + o_xy0), (float*)(base + o_r0),
                        (int*)(base + o_n0), cap);
    int found = 0;
    if (st == ZS_OK) {
        ZS_CUDA(cudaMemcpyAsync(&found, base + o_n0, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaStreamSynchronize(ctx->stream));
        if (found > cap) {
            *n_out = found;
            zs_set_error("full-frame FAST found %d corners, capacity is %d", found, cap);
            return ZS_ERR_CAPACITY;
        }
        st = zs_orb_compute(ctx, p, 0, 1, (const float*)(base + o_xy0), (const float*)(base + o_r0), nullptr,
                            (const int*)(base + o_n0), cap, (float*)(base + o_xy), (float*)(base + o_r), nullptr,
                            (int*)(base + o_n), base + o_desc);
    }
    if (st != ZS_OK) return st;
    int n = 0;
    ZS_CUDA(cudaMemcpyAsync(&n, base + o_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n > 0) {
        void* pin;
        if ((st = zs_pinned(ctx, sizeof(float) * 2 * n, &pin)) != ZS_OK) return st;
        ZS_CUDA(cudaMemcpyAsync(pin, base + o_xy, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(response, base + o_r, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(desc, base + o_desc, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaStreamSynchronize(ctx->stream));
        const float* xy = (const float*)pin;
        for (int i = 0; i < n; ++i) { x[i] = xy[2 * i]; y[i] = xy[2 * i + 1]; }
    }
    *n_out = n;
    return ZS_OK;
}

extern "C" zs_status zs_match_host(zs_context* ctx, const void* q, int nq, const void* t, int nt, int dim, int norm, int mode,
                                   double ratio, int* query_idx, int* train_idx, float* distance, int* n_out)
{
    ZS_REQUIRE(ctx && n_out, "null argument");
    *n_out = 0;
    if (nq <= 0 || nt <= 0) return ZS_OK;     // matcher.cpp:55-56,120-123: empty in, empty out
    ZS_REQUIRE(q && t && query_idx && train_idx && distance, "null argument");
    ZS_REQUIRE(norm == 0 || norm == 1 || norm == 2, "norm must be 0 (Hamming), 1 (L2, float rows) or 2 (L2, u8 rows)");
    ZS_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (KNN + ratio) or 1 (BRUTE cross-check)");
    ZS_CUDA(cudaSetDevice(ctx->device));
    const size_t row = norm == 0 ? 32 : norm == 2 ? (size_t)dim : sizeof(float) * (size_t)dim;
    if (norm == 0) ZS_REQUIRE(dim == 32 || dim == 256 || dim == 0, "Hamming descriptors are 32-byte rows");
    const size_t o_q = 256, o_t = o_q + al256(row * nq), o_idx = o_t + al256(row * nt), o_dist = o_idx + al256(sizeof(int) * 2 * nq),
                 o_pass = o_dist + al256(sizeof(float) * 2 * nq), total = o_pass + al256(nq);
    zs_async_buffer buf(ctx->stream);
    ZS_CUDA(cudaMallocAsync((void**)&buf.p, total, ctx->stream));
    uint8_t* base = buf.p;
    const int hdr[2] = { nq, nt };
    ZS_CUDA(cudaMemcpyAsync(base, hdr, sizeof(hdr), cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(base + o_q, q, row * nq, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(base + o_t, t, row * nt, cudaMemcpyHostToDevice, ctx->stream));
    zs_status st;
    const int* d_nq = (const int*)base; const int* d_nt = d_nq + 1;
    int* d_idx = (int*)(base + o_idx); float* d_dist = (float*)(base + o_dist); uint8_t* d_pass = base + o_pass;
    if (norm == 0) {
        if (mode == 0) st = zs_match_hamming_knn2(ctx, base + o_q, d_nq, 0, base + o_t, d_nt, 0, 1, nq, nt, ratio, d_idx, d_dist, d_pass);
        else st = zs_match_hamming_cross(ctx, base + o_q, d_nq, 0, base + o_t, d_nt, 0, 1, nq, nt, d_idx, d_dist);
    } else {
        if (norm == 2) {
            if (mode == 0) st = zs_match_l2_knn2_u8(ctx, base + o_q, d_nq, base + o_t, d_nt, 1, nq, nt, dim, ratio, d_idx, d_dist, d_pass);
            else st = zs_match_l2_cross_u8(ctx, base + o_q, d_nq, base + o_t, d_nt, 1, nq, nt, dim, d_idx, d_dist);
        }
        else if (mode == 0) st = zs_match_l2_knn2(ctx, (const float*)(base + o_q), d_nq, 0, (const float*)(base + o_t), d_nt, 0, 1, nq, nt, dim, ratio, d_idx, d_dist, d_pass);
        else st = zs_match_l2_cross(ctx, (const float*)(base + o_q), d_nq, 0, (const float*)(base + o_t), d_nt, 0, 1, nq, nt, dim, d_idx, d_dist);
    }
    if (st != ZS_OK) return st;
    void* pin;
    const size_t hb = sizeof(int) * 2 * nq + sizeof(float) * 2 * nq + nq;
    if ((st = zs_pinned(ctx, hb, &pin)) != ZS_OK) return st;
    int* h_idx = (int*)pin; float* h_dist = (float*)(h_idx + 2 * nq); uint8_t* h_pass = (uint8_t*)(h_dist + 2 * nq);
    const int per = mode == 0 ? 2 : 1;
    ZS_CUDA(cudaMemcpyAsync(h_idx, d_idx, sizeof(int) * per * nq, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(h_dist, d_dist, sizeof(float) * per * nq, cudaMemcpyDeviceToHost, ctx->stream));
    if (mode == 0) ZS_CUDA(cudaMemcpyAsync(h_pass, d_pass, nq, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (norm == 1 && (st = zs_context_async_error(ctx)) != ZS_OK) return st;     // float rows that were not integers in 0..255
    int m = 0;
    for (int i = 0; i < nq; ++i) {
        if (mode == 0) {
            if (!h_pass[i]) continue;
            query_idx[m] = i; train_idx[m] = h_idx[2 * i]; distance[m] = h_dist[2 * i]; ++m;
        } else {
            if (h_idx[i] < 0) continue;
            query_idx[m] = i; train_idx[m] = h_idx[i]; distance[m] = h_dist[i]; ++m;
        }
    }
    *n_out = m;
    return ZS_OK;
}

extern "C" zs_status zs_knn_match_host(zs_context* ctx, const void* q, int nq, const void* t, int nt, int dim, int norm, int k,
                                       int cross_check, int* idx, float* dist)
{
    ZS_REQUIRE(ctx, "null argument");
    ZS_REQUIRE(k == 1 || k == 2, "k must be 1 or 2");
    ZS_REQUIRE(!cross_check || k == 1, "cross check requires k = 1 (cv::BFMatcher asserts the same)");
    ZS_REQUIRE(norm == 0 || norm == 1 || norm == 2, "norm must be 0 (Hamming), 1 (L2, float rows) or 2 (L2, u8 rows)");
    if (nq <= 0) return ZS_OK;
    ZS_REQUIRE(idx && dist, "null argument");
    if (nt <= 0) {
        for (int i = 0; i < nq * k; ++i) { idx[i] = -1; dist[i] = 0.f; }
        return ZS_OK;
    }
    ZS_REQUIRE(q && t, "null argument");
    if (norm == 0) ZS_REQUIRE(dim == 32 || dim == 256 || dim == 0, "Hamming descriptors are 32-byte rows");
    ZS_CUDA(cudaSetDevice(ctx->device));
    const size_t row = norm == 0 ? 32 : norm == 2 ? (size_t)dim : sizeof(float) * (size_t)dim;
    const size_t o_q = 256, o_t = o_q + al256(row * nq), o_idx = o_t + al256(row * nt), o_dist = o_idx + al256(sizeof(int) * 2 * nq),
                 total = o_dist + al256(sizeof(float) * 2 * nq);
    zs_async_buffer buf(ctx->stream);
    ZS_CUDA(cudaMallocAsync((void**)&buf.p, total, ctx->stream));
    uint8_t* base = buf.p;
    const int hdr[2] = { nq, nt };
    ZS_CUDA(cudaMemcpyAsync(base, hdr, sizeof(hdr), cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(base + o_q, q, row * nq, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(base + o_t, t, row * nt, cudaMemcpyHostToDevice, ctx->stream));
    const int* d_nq = (const int*)base; const int* d_nt = d_nq + 1;
    int* d_idx = (int*)(base + o_idx); float* d_dist = (float*)(base + o_dist);
    zs_status st;
    if (cross_check) {
        st = norm == 0 ? zs_match_hamming_cross(ctx, base + o_q, d_nq, 0, base + o_t, d_nt, 0, 1, nq, nt, d_idx, d_dist)
             : norm == 2 ? zs_match_l2_cross_u8(ctx, base + o_q, d_nq, base + o_t, d_nt, 1, nq, nt, dim, d_idx, d_dist)
                         : zs_match_l2_cross(ctx, (const float*)(base + o_q), d_nq, 0, (const float*)(base + o_t), d_nt, 0, 1, nq, nt,
                                             dim, d_idx, d_dist);
    } else {
        st = norm == 0 ? zs_match_hamming_knn2(ctx, base + o_q, d_nq, 0, base + o_t, d_nt, 0, 1, nq, nt, 1.0, d_idx, d_dist, nullptr)
             : norm == 2 ? zs_match_l2_knn2_u8(ctx, base + o_q, d_nq, base + o_t, d_nt, 1, nq, nt, dim, 1.0, d_idx, d_dist, nullptr)
                         : zs_match_l2_knn2(ctx, (const float*)(base + o_q), d_nq, 0, (const float*)(base + o_t), d_nt, 0, 1, nq, nt,
                                            dim, 1.0, d_idx, d_dist, nullptr);
    }
    if (st != ZS_OK) return st;
    void* pin;
    const int per = cross_check ? 1 : 2;
    if ((st = zs_pinned(ctx, (sizeof(int) + sizeof(float)) * per * (size_t)nq, &pin)) != ZS_OK) return st;
    int* h_idx = (int*)pin; float* h_dist = (float*)(h_idx + (size_t)per * nq);
    ZS_CUDA(cudaMemcpyAsync(h_idx, d_idx, sizeof(int) * per * nq, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(h_dist, d_dist, sizeof(float) * per * nq, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (norm == 1 && (st = zs_context_async_error(ctx)) != ZS_OK) return st;     // float rows that were not integers in 0..255
    for (int i = 0; i < nq; ++i)
        for (int j = 0; j < k; ++j) { idx[i * k + j] = h_idx[i * per + j]; dist[i * k + j] = h_dist[i * per + j]; }
    return ZS_OK;
}

// keypoint_tracker::assign_landmark_indices, descriptor stage (keypoint_tracker.cpp:262-287): cross-checked 1-NN of the
// new keypoints' descriptors against the landmark descriptors, kept when distance <= max_descriptor_distance.
extern "C" zs_status zs_assign_landmarks_host(zs_context* ctx, const uint8_t* keypoint_desc, int n, const uint8_t* landmark_desc,
                                              int m, double max_descriptor_distance, int* landmark_row, float* distance)
{
    ZS_REQUIRE(ctx, "null argument");
    if (n <= 0) return ZS_OK;
    ZS_REQUIRE(landmark_row && distance, "null argument");
    zs_status st = zs_knn_match_host(ctx, keypoint_desc, n, landmark_desc, m, 32, 0, 1, 1, landmark_row, distance);
    if (st != ZS_OK) return st;
    for (int i = 0; i < n; ++i)
        if (landmark_row[i] >= 0 && !((double)distance[i] <= max_descriptor_distance)) { landmark_row[i] = -1; distance[i] = 0.f; }
    return ZS_OK;
}
