// zs_orb_detect.cu -- the multi-scale ORB detector behind `feature: ORB`:
//   cv::ORB::create(500, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31, fast_threshold)->detect(image, keypoints, mask)
//   (zenslam_core/source/detection/keypoint_detector_simple.cpp:17,49; same switch in keypoint_detector_grid.cpp:18
//   and keypoint_detector_parallel.cpp) followed by cv::ORB::create()->compute(image, keypoints, descriptors)
//   (keypoint_detector_simple.cpp:27,54).
//
// Per image, all on the device (semantics pinned bit-exact to cv2 4.13 by the test oracle):
//   k_orb_resize       level l = resize(level l-1, INTER_LINEAR_EXACT): 8.8 fixed-point weights from host tables;
//                      the mask pyramid uses the same kernel + threshold(254, TOZERO)
//   k_orb_fast_score   FAST-9-16 score map of a level (same ring test / score as zs_fast.cu)
//   k_orb_fast_nms     3x3 NMS + mask + edge filter -> unordered candidate list + score histogram (atomics)
//   k_orb_thr1         retainBest(2 n_l) on the integer FAST score = a threshold read off the histogram
//   k_orb_harris       Harris response (7x7 block, k 0.04, OpenCV's float expression) of the surviving candidates
//   k_orb_select       retainBest(n_l) on the Harris response by rank counting (ties kept, like OpenCV), kept
//                      keypoints emitted in canonical order: level ascending, then raster (y, x)
//   k_orb_angle        intensity-centroid angle, warp per keypoint, exact integer moments + cv::fastAtan2's polynomial
//   k_orb_blur_plane   ORB's 7x7 sigma-2 float blur of every level (same arithmetic order as zs_orb.cu / SURVEY A.3)
//   k_orb_describe_ms  rBRIEF on the blurred level the keypoint came from, pattern rotated by its angle
// OpenCV's KeyPointsFilter::retainBest reorders with std::nth_element, so cv2's output ORDER depends on the C++
// standard library; the keypoint SET (and every attribute) is identical and the order here is canonical.
#include <math.h>
#include <stdlib.h>

#include "zs_common.cuh"
#include "zs_fast_core.cuh"
#include "../../include/zs_orb_pattern.h"

#define ORBD_MAX_LEVELS 16

struct orbd_level {
    int w, h, pitch;             // pitch: bytes per row (multiple of 16)
    float scale;                 // (float)pow((double)scale_factor, level)
    int nper;                    // features wanted at this level (computeKeyPoints' nfeaturesPerLevel)
    int ccap;                    // candidate capacity per image
    size_t plane;                // bytes per image plane (pitch * h)
    uint8_t* img; uint8_t* mask; uint8_t* blur;          // [B][h][pitch]
    int* ox; int* wx; int* oy; int* wy;                  // resize tables from level-1 (device)
    int* key; int* sc;           // candidates after NMS + mask + edge: [B][ccap]
    int* key2; float* val2;      // candidates that survive retainBest(2n): [B][ccap]
    uint8_t* keep2;              // [B][ccap]
};

struct zs_orb_detector {
    zs_context* ctx;
    int width, height, max_images, nfeatures, nlevels, edge, patch, fast_threshold, cap;
    float scale_factor, harris_scale4;
    orbd_level lv[ORBD_MAX_LEVELS];
    void* block; size_t block_bytes;
    uint8_t* score;              // [B][h0][pitch0] scratch score map (levels are processed one after another)
    int* cnt; int* cnt2; int* hist; int* thr1;           // [B][L], [B][L], [B][L][256], [B][L]
    int* umax;                   // [patch/2 + 2] row ends of the circular patch (device)
    int* lxy;                    // [B][cap][2] level coordinates of the emitted keypoints
    orbd_level* d_lv;            // device copy of lv[]
};

__device__ int4 g_orbd_pattern[256] = ZS_ORB_PATTERN_INIT;

// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_orb_copy_plane(const uint8_t* __restrict__ src, size_t spitch, size_t sstride,
                                                        uint8_t* __restrict__ dst, int w, int h, int dpitch, size_t dplane)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    dst[blockIdx.z * dplane + (size_t)y * dpitch + x] = src[blockIdx.z * sstride + (size_t)y * spitch + x];
}

// cv::resize(INTER_LINEAR_EXACT), u8: horizontal pass exact in 8.8, vertical pass rounded half up from 16.16
__global__ void __launch_bounds__(256) k_orb_resize(const uint8_t* __restrict__ src, int sw, int sh, int spitch, size_t splane,
                                                    uint8_t* __restrict__ dst, int dw, int dh, int dpitch, size_t dplane,
                                                    const int* __restrict__ ox, const int* __restrict__ wx,
                                                    const int* __restrict__ oy, const int* __restrict__ wy, int tozero254)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const uint8_t* s = src + blockIdx.z * splane;
    const int x0 = ox[x], x1 = min(x0 + 1, sw - 1), y0 = oy[y], y1 = min(y0 + 1, sh - 1);
    const int a = wx[x], b = wy[y];
    const uint8_t* r0 = s + (size_t)y0 * spitch; const uint8_t* r1 = s + (size_t)y1 * spitch;
    const int h0 = r0[x0] * (256 - a) + r0[x1] * a, h1 = r1[x0] * (256 - a) + r1[x1] * a;
    int v = (h0 * (256 - b) + h1 * b + (1 << 15)) >> 16;
    if (tozero254 && v <= 254) v = 0;
    dst[blockIdx.z * dplane + (size_t)y * dpitch + x] = (uint8_t)v;
}

__global__ void __launch_bounds__(256) k_orb_fast_score(const uint8_t* __restrict__ img, int w, int h, int pitch, size_t plane,
                                                        int threshold, uint8_t* __restrict__ score, size_t splane)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    uint8_t out = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {
        int d[16];
        if (fast_is_corner(img + blockIdx.z * plane, pitch, x, y, threshold, d)) out = (uint8_t)fast_score(d);
    }
    score[blockIdx.z * splane + (size_t)y * pitch + x] = out;
}

// NMS (strictly greater than the 8 neighbours), KeyPointsFilter::runByPixelsMask, runByImageBorder(edge)
__global__ void __launch_bounds__(256) k_orb_fast_nms(const uint8_t* __restrict__ score, size_t splane, const uint8_t* __restrict__ mask,
                                                      size_t mplane, int w, int h, int pitch, int edge, int ccap, int level,
                                                      int nlevels, int* __restrict__ key, int* __restrict__ sc,
                                                      int* __restrict__ cnt, int* __restrict__ hist)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int lo = max(edge, 3);
    if (x < lo || x >= w - lo || y < lo || y >= h - lo) return;
    const int img = blockIdx.z;
    const uint8_t* r1 = score + img * splane + (size_t)y * pitch;
    const int s = r1[x];
    if (s == 0) return;
    const uint8_t* r0 = r1 - pitch; const uint8_t* r2 = r1 + pitch;
    if (!(s > r0[x - 1] && s > r0[x] && s > r0[x + 1] && s > r1[x - 1] && s > r1[x + 1] && s > r2[x - 1] && s > r2[x] && s > r2[x + 1]))
        return;
    if (mask && mask[img * mplane + (size_t)y * pitch + x] == 0) return;
    const int slot = atomicAdd(&cnt[img * nlevels + level], 1);
    if (slot < ccap) { key[(size_t)img * ccap + slot] = y * w + x; sc[(size_t)img * ccap + slot] = s; }
    atomicAdd(&hist[(img * nlevels + level) * 256 + s], 1);
}

// retainBest(2 n_l) on the FAST score: the smallest score that is still among the 2 n_l best (ties kept)
__global__ void k_orb_thr1(const orbd_level* __restrict__ lv, int nlevels, const int* __restrict__ cnt, const int* __restrict__ hist,
                           int* __restrict__ thr1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= gridDim.x * blockDim.x) return;
    const int level = i % nlevels;
    const int n = 2 * lv[level].nper, total = cnt[i];
    int t = 0;
    if (total > n) {
        if (n == 0) t = 256;
        else {
            int cum = 0;
            for (t = 255; t > 0; --t) { cum += hist[i * 256 + t]; if (cum >= n) break; }
        }
    }
    thr1[i] = t;
}

// HarrisResponses (orb.cpp): blockSize 7, integer gradient sums, float response with OpenCV's expression order
__global__ void __launch_bounds__(256) k_orb_harris(orbd_level L, int level, int nlevels, const int* __restrict__ cnt,
                                                    const int* __restrict__ thr1, int* __restrict__ cnt2, float scale4)
{
    const int img = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = min(cnt[img * nlevels + level], L.ccap);
    if (i >= m) return;
    const size_t o = (size_t)img * L.ccap + i;
    if (L.sc[o] < thr1[img * nlevels + level]) return;
    const int k = L.key[o], x = k % L.w, y = k / L.w, p = L.pitch;
    const uint8_t* c = L.img + img * L.plane + (size_t)y * p + x;
    int a = 0, b = 0, cc = 0;
    for (int dy = -3; dy <= 3; ++dy) {
        const uint8_t* r = c + dy * p;
#pragma unroll
        for (int dx = -3; dx <= 3; ++dx) {
            const uint8_t* q = r + dx;
            const int Ix = ((int)q[1] - (int)q[-1]) * 2 + ((int)q[-p + 1] - (int)q[-p - 1]) + ((int)q[p + 1] - (int)q[p - 1]);
            const int Iy = ((int)q[p] - (int)q[-p]) * 2 + ((int)q[p - 1] - (int)q[-p - 1]) + ((int)q[p + 1] - (int)q[-p + 1]);
            a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
        }
    }
    const float fa = (float)a, fb = (float)b, fc = (float)cc;
    const float tr = __fadd_rn(fa, fb);
    const float r = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, tr), tr)), scale4);
    const int slot = atomicAdd(&cnt2[img * nlevels + level], 1);
    L.key2[(size_t)img * L.ccap + slot] = k;
    L.val2[(size_t)img * L.ccap + slot] = r;
}

// One block per image, levels in turn: keep candidate i iff fewer than n_l candidates have a strictly larger Harris
// response (== response >= the n_l-th largest: KeyPointsFilter::retainBest with its ties), then place the kept ones in
// raster order behind the previous levels' keypoints.
__global__ void __launch_bounds__(1024) k_orb_select(const orbd_level* __restrict__ lv, int nlevels, const int* __restrict__ cnt2,
                                                     int cap, int patch, float* __restrict__ oxy, float* __restrict__ osize,
                                                     float* __restrict__ oresp, int* __restrict__ ooct, int* __restrict__ lxy,
                                                     int* __restrict__ ocount)
{
    __shared__ int s_kept;
    const int img = blockIdx.x;
    int base = 0;
    for (int level = 0; level < nlevels; ++level) {
        const orbd_level L = lv[level];
        const int m = min(cnt2[img * nlevels + level], L.ccap), n = L.nper;
        const int* key = L.key2 + (size_t)img * L.ccap;
        const float* val = L.val2 + (size_t)img * L.ccap;
        uint8_t* keep = L.keep2 + (size_t)img * L.ccap;
        if (threadIdx.x == 0) s_kept = 0;
        __syncthreads();
        int mine = 0;
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            bool k = true;
            if (m > n) {
                if (n == 0) k = false;
                else {
                    const float v = val[i];
                    int g = 0;
                    for (int j = 0; j < m; ++j) g += val[j] > v;
                    k = g < n;
                }
            }
            keep[i] = k;
            mine += k;
        }
        if (mine) atomicAdd(&s_kept, mine);
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            if (!keep[i]) continue;
            const int ki = key[i];
            int pos = 0;
            for (int j = 0; j < m; ++j) pos += (keep[j] && key[j] < ki);
            const int o = base + pos;
            if (o < cap) {
                const size_t q = (size_t)img * cap + o;
                const int x = ki % L.w, y = ki / L.w;
                oxy[2 * q] = __fmul_rn((float)x, L.scale); oxy[2 * q + 1] = __fmul_rn((float)y, L.scale);
                osize[q] = __fmul_rn((float)patch, L.scale);
                oresp[q] = val[i]; ooct[q] = level;
                lxy[2 * q] = x; lxy[2 * q + 1] = y;
            }
        }
        base += s_kept;
        __syncthreads();
    }
    if (threadIdx.x == 0) ocount[img] = base;
}

// cv::fastAtan2 (scalar path, degrees)
__device__ __forceinline__ float orbd_fast_atan2(float y, float x)
{
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k),
                p5 = __fmul_rn(0.1555786518463281f, k), p7 = __fmul_rn(-0.04432655554792128f, k);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps)); c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps)); c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// ICAngles (orb.cpp): m10 = sum u I, m01 = sum v I over the circular patch; warp per keypoint, lanes over u
__global__ void __launch_bounds__(256) k_orb_angle(const orbd_level* __restrict__ lv, const int* __restrict__ umax, int half,
                                                   const int* __restrict__ ocount, int cap, const int* __restrict__ ooct,
                                                   const int* __restrict__ lxy, float* __restrict__ oangle)
{
    const int img = blockIdx.y, kp = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (kp >= min(ocount[img], cap)) return;
    const size_t q = (size_t)img * cap + kp;
    const orbd_level L = lv[ooct[q]];
    const uint8_t* c = L.img + img * L.plane + (size_t)lxy[2 * q + 1] * L.pitch + lxy[2 * q];
    int m01 = 0, m10 = 0;
    for (int u = -half + lane; u <= half; u += 32) {
        m10 += u * (int)c[u];
        for (int v = 1; v <= half; ++v) {
            if (abs(u) > umax[v]) continue;
            const int vp = c[u + v * L.pitch], vm = c[u - v * L.pitch];
            m01 += v * (vp - vm);
            m10 += u * (vp + vm);
        }
    }
    m01 = __reduce_add_sync(0xffffffffu, m01); m10 = __reduce_add_sync(0xffffffffu, m10);
    if (lane == 0) oangle[q] = orbd_fast_atan2((float)m01, (float)m10);
}

// taps of cv::getGaussianKernel(7, 2, CV_32F)
#define OG0 0x1.1f5f62p-4f
#define OG1 0x1.0c70fcp-3f
#define OG2 0x1.869472p-3f
#define OG3 0x1.ba95c0p-3f

// ORB's blur on an arbitrary un-padded plane (REFLECT_101 by index): 32 x 8 outputs per block
__global__ void __launch_bounds__(256) k_orb_blur_plane(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int w, int h,
                                                        int pitch, size_t plane)
{
    __shared__ uint8_t s_in[14][40];
    __shared__ float s_row[14][32];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const uint8_t* s = src + blockIdx.z * plane;
    for (int i = threadIdx.x; i < 14 * 38; i += 256) {
        const int r = i / 38, c = i - r * 38;
        s_in[r][c] = s[(size_t)zs_reflect101(y0 + r - 3, h) * pitch + zs_reflect101(x0 + c - 3, w)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 14 * 32; i += 256) {
        const int r = i >> 5, c = i & 31;
        const uint8_t* p = &s_in[r][c];
        float t = __fmul_rn(OG0, (float)p[0]);
        t = fmaf((float)p[1], OG1, t); t = fmaf((float)p[2], OG2, t); t = fmaf((float)p[3], OG3, t);
        t = fmaf((float)p[4], OG2, t); t = fmaf((float)p[5], OG1, t); t = fmaf((float)p[6], OG0, t);
        s_row[r][c] = t;
    }
    __syncthreads();
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int x = x0 + c, y = y0 + r;
    if (x >= w || y >= h) return;
    float t = __fmul_rn(OG3, s_row[r + 3][c]);
    t = fmaf(__fadd_rn(s_row[r + 4][c], s_row[r + 2][c]), OG2, t);
    t = fmaf(__fadd_rn(s_row[r + 5][c], s_row[r + 1][c]), OG1, t);
    t = fmaf(__fadd_rn(s_row[r + 6][c], s_row[r][c]), OG0, t);
    int q = __float2int_rn(t);
    q = q < 0 ? 0 : q > 255 ? 255 : q;
    dst[blockIdx.z * plane + (size_t)y * pitch + x] = (uint8_t)q;
}

// computeOrbDescriptors (orb.cpp) for keypoints that carry an octave and an angle: centre = cvRound(pt * (1.f / scale))
// on the blurred level image; warp per keypoint, lane = descriptor byte
__global__ void __launch_bounds__(256) k_orb_describe_ms(const orbd_level* __restrict__ lv, const int* __restrict__ ocount, int cap,
                                                         const float* __restrict__ oxy, const float* __restrict__ oangle,
                                                         const int* __restrict__ ooct, uint8_t* __restrict__ desc)
{
    const int img = blockIdx.y, kp = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (kp >= min(ocount[img], cap)) return;
    const size_t q = (size_t)img * cap + kp;
    const orbd_level L = lv[ooct[q]];
    float a = 0.f, b = 0.f;
    if (lane == 0) {
        // OpenCV: angle *= (float)(CV_PI/180.f); a = (float)cos(angle), b = (float)sin(angle)  (double math)
        const float ang = __fmul_rn(oangle[q], (float)(3.14159265358979323846 / 180.0));
        a = (float)cos((double)ang); b = (float)sin((double)ang);
    }
    a = __shfl_sync(0xffffffffu, a, 0); b = __shfl_sync(0xffffffffu, b, 0);
    const float inv = __fdiv_rn(1.f, L.scale);
    const int cx = __float2int_rn(__fmul_rn(oxy[2 * q], inv)), cy = __float2int_rn(__fmul_rn(oxy[2 * q + 1], inv));
    const uint8_t* c = L.blur + img * L.plane + (size_t)cy * L.pitch + cx;
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const int4 pt = g_orbd_pattern[lane * 8 + bit];
        const float x0f = __fsub_rn(__fmul_rn((float)pt.x, a), __fmul_rn((float)pt.y, b));
        const float y0f = __fadd_rn(__fmul_rn((float)pt.x, b), __fmul_rn((float)pt.y, a));
        const float x1f = __fsub_rn(__fmul_rn((float)pt.z, a), __fmul_rn((float)pt.w, b));
        const float y1f = __fadd_rn(__fmul_rn((float)pt.z, b), __fmul_rn((float)pt.w, a));
        const int t0 = c[__float2int_rn(y0f) * L.pitch + __float2int_rn(x0f)];
        const int t1 = c[__float2int_rn(y1f) * L.pitch + __float2int_rn(x1f)];
        val |= (t0 < t1) << bit;
    }
    desc[q * 32 + lane] = (uint8_t)val;
}

// ------------------------------------------------------------------------------------------------------
static void orbd_lin_coeffs(int dn, int sn, int* ofs, int* w1)
{
    const double scale = (double)sn / (double)dn;
    for (int d = 0; d < dn; ++d) {
        double f = ((double)d + 0.5) * scale - 0.5;
        int si = (int)floor(f);
        f -= si;
        if (si < 0) { si = 0; f = 0.0; }
        if (si >= sn - 1) { si = sn - 1; f = 0.0; }
        ofs[d] = si;
        w1[d] = (int)floor(f * 256.0 + 0.5);
    }
}

static inline size_t orbd_al(size_t v) { return (v + 255) / 256 * 256; }

extern "C" zs_status zs_orb_detector_create(zs_context* ctx, int width, int height, int max_images, int nfeatures,
                                            float scale_factor, int nlevels, int edge_threshold, int patch_size,
                                            int fast_threshold, zs_orb_detector** out)
{
    ZS_REQUIRE(ctx && out, "null argument");
    ZS_REQUIRE(width > 0 && height > 0 && max_images > 0, "bad image size / count");
    ZS_REQUIRE(nfeatures >= 0 && nlevels >= 1 && nlevels <= ORBD_MAX_LEVELS, "nfeatures < 0 or nlevels outside 1..16");
    ZS_REQUIRE(scale_factor > 1.f, "scale_factor must be > 1");
    ZS_REQUIRE(edge_threshold >= 19 && patch_size >= 3 && patch_size <= 63 && (patch_size / 2) <= edge_threshold - 4,
               "edge_threshold must cover the descriptor pattern (>= 19), the Harris block and the orientation patch");
    ZS_CUDA(cudaSetDevice(ctx->device));
    zs_orb_detector* d = new zs_orb_detector();
    memset(d, 0, sizeof(*d));
    d->ctx = ctx; d->width = width; d->height = height; d->max_images = max_images; d->nfeatures = nfeatures;
    d->nlevels = nlevels; d->edge = edge_threshold; d->patch = patch_size; d->scale_factor = scale_factor;
    d->fast_threshold = fast_threshold < 0 ? 0 : fast_threshold > 255 ? 255 : fast_threshold;
    // retainBest keeps ties, so a level can emit more than its share: leave room
    d->cap = nfeatures + 32 * nlevels + 64;
    {
        const float scale = 1.f / ((1 << 2) * 7 * 255.f);
        d->harris_scale4 = scale * scale * scale * scale;
    }
    // computeKeyPoints: features per level
    {
        const float factor = (float)(1.0 / (double)scale_factor);
        float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; ++l) {
            d->lv[l].nper = (int)lrint((double)nd);
            sum += d->lv[l].nper;
            nd *= factor;
        }
        d->lv[nlevels - 1].nper = nfeatures - sum > 0 ? nfeatures - sum : 0;
    }
    const size_t B = (size_t)max_images;
    size_t total = 0;
    for (int l = 0; l < nlevels; ++l) {
        orbd_level& L = d->lv[l];
        L.scale = (float)pow((double)scale_factor, (double)l);
        const float inv = 1.0f / L.scale;
        L.w = (int)lrint((double)((float)width * inv)); L.h = (int)lrint((double)((float)height * inv));
        if (L.w < 1 || L.h < 1) { delete d; zs_set_error("image too small for %d pyramid levels", nlevels); return ZS_ERR_INVALID; }
        L.pitch = (L.w + 15) & ~15;
        L.plane = (size_t)L.pitch * L.h;
        L.ccap = ((L.w + 1) / 2) * ((L.h + 1) / 2) + 1;      // NMS survivors are never 8-neighbours
        total += 3 * orbd_al(B * L.plane);                   // img, mask, blur
        total += 4 * orbd_al(sizeof(int) * (size_t)(L.w + L.h));
        total += 4 * orbd_al(sizeof(int) * B * L.ccap) + orbd_al(B * L.ccap);
    }
    total += orbd_al(B * d->lv[0].plane);                    // score
    total += orbd_al(sizeof(int) * B * nlevels * (3 + 256)); // cnt, cnt2, thr1, hist
    total += orbd_al(sizeof(int) * 128) + orbd_al(sizeof(int) * 2 * B * d->cap) + orbd_al(sizeof(orbd_level) * ORBD_MAX_LEVELS);
    cudaError_t e = cudaMalloc(&d->block, total);
    if (e != cudaSuccess) { delete d; return zs_cuda_fail(e, "cudaMalloc(orb detector)", __FILE__, __LINE__); }
    d->block_bytes = total;
    uint8_t* p = (uint8_t*)d->block;
    auto take = [&](size_t bytes) { uint8_t* r = p; p += orbd_al(bytes); return r; };
    for (int l = 0; l < nlevels; ++l) {
        orbd_level& L = d->lv[l];
        L.img = take(B * L.plane); L.mask = take(B * L.plane); L.blur = take(B * L.plane);
        L.ox = (int*)take(sizeof(int) * (L.w + L.h)); L.wx = (int*)take(sizeof(int) * (L.w + L.h));
        L.oy = (int*)take(sizeof(int) * (L.w + L.h)); L.wy = (int*)take(sizeof(int) * (L.w + L.h));
        L.key = (int*)take(sizeof(int) * B * L.ccap); L.sc = (int*)take(sizeof(int) * B * L.ccap);
        L.key2 = (int*)take(sizeof(int) * B * L.ccap); L.val2 = (float*)take(sizeof(int) * B * L.ccap);
        L.keep2 = take(B * L.ccap);
    }
    d->score = take(B * d->lv[0].plane);
    int* counters = (int*)take(sizeof(int) * B * nlevels * (3 + 256));
    d->cnt = counters; d->cnt2 = counters + B * nlevels; d->thr1 = counters + 2 * B * nlevels; d->hist = counters + 3 * B * nlevels;
    d->umax = (int*)take(sizeof(int) * 128);
    d->lxy = (int*)take(sizeof(int) * 2 * B * d->cap);
    d->d_lv = (orbd_level*)take(sizeof(orbd_level) * ORBD_MAX_LEVELS);
    // host tables -> device (synchronous copies: creation is not on the hot path)
    zs_status st = ZS_OK;
    for (int l = 1; l < nlevels && st == ZS_OK; ++l) {
        orbd_level& L = d->lv[l];
        const orbd_level& S = d->lv[l - 1];
        int* t = (int*)malloc(sizeof(int) * 2 * (size_t)(L.w + L.h));
        orbd_lin_coeffs(L.w, S.w, t, t + L.w);
        orbd_lin_coeffs(L.h, S.h, t + 2 * L.w, t + 2 * L.w + L.h);
        if (cudaMemcpy(L.ox, t, sizeof(int) * L.w, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(L.wx, t + L.w, sizeof(int) * L.w, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(L.oy, t + 2 * L.w, sizeof(int) * L.h, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(L.wy, t + 2 * L.w + L.h, sizeof(int) * L.h, cudaMemcpyHostToDevice) != cudaSuccess)
            st = zs_cuda_fail(cudaGetLastError(), "cudaMemcpy(resize tables)", __FILE__, __LINE__);
        free(t);
    }
    if (st == ZS_OK) {
        // row ends of the circular patch (orb.cpp computeKeyPoints)
        const int half = patch_size / 2;
        int umax[128];
        memset(umax, 0, sizeof(umax));
        const int vmax = (int)floor(half * sqrt(2.f) / 2 + 1), vmin = (int)ceil(half * sqrt(2.f) / 2);
        for (int v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt((double)half * half - v * v));
        for (int v = half, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
        if (cudaMemcpy(d->umax, umax, sizeof(umax), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(d->d_lv, d->lv, sizeof(orbd_level) * ORBD_MAX_LEVELS, cudaMemcpyHostToDevice) != cudaSuccess)
            st = zs_cuda_fail(cudaGetLastError(), "cudaMemcpy(orb tables)", __FILE__, __LINE__);
    }
    if (st != ZS_OK) { cudaFree(d->block); delete d; return st; }
    *out = d;
    return ZS_OK;
}

extern "C" void zs_orb_detector_destroy(zs_orb_detector* d)
{
    if (!d) return;
    cudaFree(d->block);
    delete d;
}

extern "C" int zs_orb_detector_capacity(const zs_orb_detector* d) { return d ? d->cap : 0; }

extern "C" zs_status zs_orb_detector_level(const zs_orb_detector* d, int level, int* width, int* height, float* scale, int* nfeatures)
{
    ZS_REQUIRE(d && level >= 0 && level < d->nlevels, "bad level");
    if (width) *width = d->lv[level].w;
    if (height) *height = d->lv[level].h;
    if (scale) *scale = d->lv[level].scale;
    if (nfeatures) *nfeatures = d->lv[level].nper;
    return ZS_OK;
}

extern "C" zs_status zs_orb_detector_download_level(zs_context* ctx, const zs_orb_detector* d, int image, int level, int which,
                                                    uint8_t* dst)
{
    ZS_REQUIRE(ctx && d && dst && level >= 0 && level < d->nlevels && image >= 0 && image < d->max_images, "bad argument");
    const orbd_level& L = d->lv[level];
    const uint8_t* src = (which == 0 ? L.img : which == 1 ? L.mask : L.blur) + (size_t)image * L.plane;
    ZS_CUDA(cudaMemcpy2DAsync(dst, L.w, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    return ZS_OK;
}

extern "C" zs_status zs_orb_detect_and_compute(zs_context* ctx, zs_orb_detector* d, const uint8_t* d_img, size_t pitch, size_t stride,
                                               const uint8_t* d_mask, size_t mask_pitch, size_t mask_stride, int count,
                                               float* d_xy, float* d_size, float* d_angle, float* d_response, int* d_octave,
                                               int* d_count, uint8_t* d_desc)
{
    ZS_REQUIRE(ctx && d && d_img && d_xy && d_size && d_angle && d_response && d_octave && d_count, "null argument");
    ZS_REQUIRE(count >= 0 && count <= d->max_images, "count exceeds the detector's max_images");
    ZS_REQUIRE(pitch >= (size_t)d->width && (!d_mask || mask_pitch >= (size_t)d->width), "pitch smaller than the width");
    ZS_REQUIRE(!d_desc || (d->edge >= 31 && d->patch == 31),
               "describing needs edge_threshold >= 31 and patch_size 31 (cv::ORB::create() defaults on the compute side)");
    if (count == 0) return ZS_OK;
    cudaStream_t s = ctx->stream;
    const int NL = d->nlevels;
    ZS_CUDA(cudaMemsetAsync(d->cnt, 0, sizeof(int) * (size_t)d->max_images * NL * (3 + 256), s));
    for (int l = 0; l < NL; ++l) {
        const orbd_level& L = d->lv[l];
        const dim3 grid(zs_div_up(L.w, 32), zs_div_up(L.h, 8), count);
        if (l == 0) {
            k_orb_copy_plane<<<grid, 256, 0, s>>>(d_img, pitch, stride, L.img, L.w, L.h, L.pitch, L.plane);
            ZS_LAUNCH_CHECK(ctx);
            if (d_mask) {
                k_orb_copy_plane<<<grid, 256, 0, s>>>(d_mask, mask_pitch, mask_stride, L.mask, L.w, L.h, L.pitch, L.plane);
                ZS_LAUNCH_CHECK(ctx);
            }
        } else {
            const orbd_level& S = d->lv[l - 1];
            k_orb_resize<<<grid, 256, 0, s>>>(S.img, S.w, S.h, S.pitch, S.plane, L.img, L.w, L.h, L.pitch, L.plane, L.ox, L.wx, L.oy, L.wy, 0);
            ZS_LAUNCH_CHECK(ctx);
            if (d_mask) {
                k_orb_resize<<<grid, 256, 0, s>>>(S.mask, S.w, S.h, S.pitch, S.plane, L.mask, L.w, L.h, L.pitch, L.plane, L.ox, L.wx, L.oy,
                                                  L.wy, 1);
                ZS_LAUNCH_CHECK(ctx);
            }
        }
        if (L.w > 2 * d->edge && L.h > 2 * d->edge) {
            k_orb_fast_score<<<grid, 256, 0, s>>>(L.img, L.w, L.h, L.pitch, L.plane, d->fast_threshold, d->score, d->lv[0].plane);
            ZS_LAUNCH_CHECK(ctx);
            k_orb_fast_nms<<<grid, 256, 0, s>>>(d->score, d->lv[0].plane, d_mask ? L.mask : nullptr, L.plane, L.w, L.h, L.pitch, d->edge,
                                                L.ccap, l, NL, L.key, L.sc, d->cnt, d->hist);
            ZS_LAUNCH_CHECK(ctx);
        }
    }
    k_orb_thr1<<<count, NL, 0, s>>>(d->d_lv, NL, d->cnt, d->hist, d->thr1);
    ZS_LAUNCH_CHECK(ctx);
    for (int l = 0; l < NL; ++l) {
        const orbd_level& L = d->lv[l];
        if (!(L.w > 2 * d->edge && L.h > 2 * d->edge)) continue;
        k_orb_harris<<<dim3(zs_div_up(L.ccap, 256), count), 256, 0, s>>>(L, l, NL, d->cnt, d->thr1, d->cnt2, d->harris_scale4);
        ZS_LAUNCH_CHECK(ctx);
    }
    k_orb_select<<<count, 1024, 0, s>>>(d->d_lv, NL, d->cnt2, d->cap, d->patch, d_xy, d_size, d_response, d_octave, d->lxy, d_count);
    ZS_LAUNCH_CHECK(ctx);
    k_orb_angle<<<dim3(zs_div_up(d->cap, 8), count), 256, 0, s>>>(d->d_lv, d->umax, d->patch / 2, d_count, d->cap, d_octave, d->lxy, d_angle);
    ZS_LAUNCH_CHECK(ctx);
    if (d_desc) {
        for (int l = 0; l < NL; ++l) {
            const orbd_level& L = d->lv[l];
            if (!(L.w > 2 * d->edge && L.h > 2 * d->edge)) continue;
            k_orb_blur_plane<<<dim3(zs_div_up(L.w, 32), zs_div_up(L.h, 8), count), 256, 0, s>>>(L.img, L.blur, L.w, L.h, L.pitch, L.plane);
            ZS_LAUNCH_CHECK(ctx);
        }
        k_orb_describe_ms<<<dim3(zs_div_up(d->cap, 8), count), 256, 0, s>>>(d->d_lv, d_count, d->cap, d_xy, d_angle, d_octave, d_desc);
        ZS_LAUNCH_CHECK(ctx);
    }
    return ZS_OK;
}

// keypoint_detector_simple::detect_keypoints with `feature: ORB` (keypoint_detector_simple.cpp:38-63) on host buffers
extern "C" zs_status zs_detect_keypoints_orb_host(zs_context* ctx, const uint8_t* img, int width, int height, size_t pitch,
                                                  const uint8_t* mask, size_t mask_pitch, int nfeatures, float scale_factor, int nlevels,
                                                  int edge_threshold, int patch_size, int fast_threshold, float* x, float* y,
                                                  float* size, float* angle, float* response, int* octave, uint8_t* desc, int cap,
                                                  int* n_out)
{
    ZS_REQUIRE(ctx && img && x && y && size && angle && response && octave && n_out, "null argument");
    ZS_REQUIRE(width > 0 && height > 0 && pitch >= (size_t)width && cap >= 0, "bad image geometry");
    ZS_CUDA(cudaSetDevice(ctx->device));
    *n_out = 0;
    const int key[8] = { width, height, nfeatures, nlevels, edge_threshold, patch_size, fast_threshold, 1 };
    if (ctx->host_orb && (memcmp(key, ctx->host_orb_key, sizeof(key)) != 0 || ctx->host_orb_sf != scale_factor)) {
        ZS_CUDA(cudaStreamSynchronize(ctx->stream));
        zs_orb_detector_destroy(ctx->host_orb);
        ctx->host_orb = nullptr;
    }
    zs_status st;
    if (!ctx->host_orb) {
        if ((st = zs_orb_detector_create(ctx, width, height, 1, nfeatures, scale_factor, nlevels, edge_threshold, patch_size,
                                         fast_threshold, &ctx->host_orb)) != ZS_OK)
            return st;
        memcpy(ctx->host_orb_key, key, sizeof(key));
        ctx->host_orb_sf = scale_factor;
    }
    zs_orb_detector* d = ctx->host_orb;
    const int dc = d->cap;
    const size_t ipitch = ((size_t)width + 15) & ~(size_t)15, iplane = orbd_al(ipitch * height);
    const size_t o_img = 0, o_mask = iplane, o_xy = 2 * iplane, o_size = o_xy + orbd_al(sizeof(float) * 2 * dc),
                 o_ang = o_size + orbd_al(sizeof(float) * dc), o_resp = o_ang + orbd_al(sizeof(float) * dc),
                 o_oct = o_resp + orbd_al(sizeof(float) * dc), o_n = o_oct + orbd_al(sizeof(int) * dc), o_desc = o_n + 256,
                 total = o_desc + orbd_al((size_t)dc * 32);
    void* sc;
    if ((st = zs_scratch(ctx, total, &sc)) != ZS_OK) return st;
    uint8_t* base = (uint8_t*)sc;
    ZS_CUDA(cudaMemcpy2DAsync(base + o_img, ipitch, img, pitch, width, height, cudaMemcpyHostToDevice, ctx->stream));
    if (mask) ZS_CUDA(cudaMemcpy2DAsync(base + o_mask, ipitch, mask, mask_pitch, width, height, cudaMemcpyHostToDevice, ctx->stream));
    st = zs_orb_detect_and_compute(ctx, d, base + o_img, ipitch, iplane, mask ? base + o_mask : nullptr, ipitch, iplane, 1,
                                   (float*)(base + o_xy), (float*)(base + o_size), (float*)(base + o_ang), (float*)(base + o_resp),
                                   (int*)(base + o_oct), (int*)(base + o_n), desc ? base + o_desc : nullptr);
    if (st != ZS_OK) return st;
    int n = 0;
    ZS_CUDA(cudaMemcpyAsync(&n, base + o_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    if (n > dc || n > cap) { zs_set_error("ORB detector found %d keypoints, capacity is %d", n, n > dc ? dc : cap); return ZS_ERR_CAPACITY; }
    if (n > 0) {
        void* pin;
        if ((st = zs_pinned(ctx, sizeof(float) * 2 * n, &pin)) != ZS_OK) return st;
        ZS_CUDA(cudaMemcpyAsync(pin, base + o_xy, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(size, base + o_size, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(angle, base + o_ang, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(response, base + o_resp, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(octave, base + o_oct, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
        if (desc) ZS_CUDA(cudaMemcpyAsync(desc, base + o_desc, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaStreamSynchronize(ctx->stream));
        const float* xy = (const float*)pin;
        for (int i = 0; i < n; ++i) { x[i] = xy[2 * i]; y[i] = xy[2 * i + 1]; }
    }
    return ZS_OK;
}
