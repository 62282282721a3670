// zs_subpix.cu -- cv::cornerSubPix for the PARALLEL_GRID detector.
// Reference: keypoint_detector_parallel::detect_keypoints refines every selected corner with
//   cv::cornerSubPix(image, pts, Size(5,5), Size(-1,-1), {EPS+COUNT, 30, 0.01})
// (zenslam_core/source/detection/keypoint_detector_parallel.cpp:160-170); arithmetic in SURVEY A.10.
//
// One warp per point.  Per iteration the (2w+3)^2 float patch around the current estimate is produced in shared
// memory exactly as OpenCV's 8u->32f getRectSubPix does it -- inside the image that is a per-row recurrence
// (lane = patch row), at the image border the replicate-border bilinear form (lanes stride over the patch) --
// and the five double-precision sums of the normal equations are accumulated by five lanes, each adding its terms in
// raster order, so that every rounding happens in the same order as on the CPU: results are bit-identical to OpenCV
// built without IPP (the oracle pins that), not merely within tolerance.  The TERMS are order-free, so all 32 lanes
// compute them first (one window pixel per lane and step) and park them in shared memory; only the 121 dependent
// additions per sum stay serial (the first version had five lanes do everything: 10 ms per 57 k points at tumvi.yaml's
// settings, 5/32 lane utilisation).
#include <math.h>

#include "zs_common.cuh"

#define SUBPIX_MAX_WIN 7                       // half window; the reference uses 5
#define SUBPIX_WARPS 4
#define SUBPIX_NPX ((2 * SUBPIX_MAX_WIN + 1) * (2 * SUBPIX_MAX_WIN + 1))

struct subpix_args {
    zs_pyr_view v;
    int first;
    float2* xy; const int* count; int cap;
    int win_w, win_h, max_iters;
    double eps2;
    float ex[2 * SUBPIX_MAX_WIN + 1], ey[2 * SUBPIX_MAX_WIN + 1];   // expf(-x*x) tables, computed on the host like OpenCV does
};

__global__ void __launch_bounds__(SUBPIX_WARPS * 32) k_corner_subpix(subpix_args a)
{
    __shared__ float s_sub[SUBPIX_WARPS][(2 * SUBPIX_MAX_WIN + 3) * (2 * SUBPIX_MAX_WIN + 3)];
    __shared__ double s_term[SUBPIX_WARPS][5 * SUBPIX_NPX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.y, i = blockIdx.x * SUBPIX_WARPS + warp;
    if (i >= min(a.count[img], a.cap)) return;
    const int w = a.v.w[0], h = a.v.h[0], pitch = a.v.pitch[0];
    const int slot = zs_slot(a.first, img, a.v.slots);
    const uint8_t* src = a.v.img[0] + (size_t)slot * a.v.slot_stride[0] + (size_t)a.v.pad_y * pitch + a.v.pad_x;
    float* sub = s_sub[warp];
    const int ww = 2 * a.win_w + 1, wh = 2 * a.win_h + 1, pw = ww + 2, ph = wh + 2;
    const float2 cT = a.xy[(size_t)img * a.cap + i];
    float cix = cT.x, ciy = cT.y;
    int iter = 0;
    double err = 0;
    do {
        // ---- patch (getRectSubPix_8u32f)
        const float centerx = __fsub_rn(cix, __fmul_rn((float)(pw - 1), 0.5f)), centery = __fsub_rn(ciy, __fmul_rn((float)(ph - 1), 0.5f));
        const int ipx = __float2int_rd(centerx), ipy = __float2int_rd(centery);
        __syncwarp();
        if (0 <= ipx && ipx + pw < w && 0 <= ipy && ipy + ph < h) {
            float fa = __fsub_rn(centerx, (float)ipx);
            const float fb = __fsub_rn(centery, (float)ipy);
            fa = fmaxf(fa, 0.0001f);
            const float a12 = __fmul_rn(fa, __fsub_rn(1.f, fb)), a22 = __fmul_rn(fa, fb), b1 = __fsub_rn(1.f, fb), b2 = fb;
            const double s = (1. - (double)fa) / (double)fa;
            for (int r = lane; r < ph; r += 32) {
                const uint8_t* p = src + (size_t)(ipy + r) * pitch + ipx;
                float prev = __fmul_rn(__fsub_rn(1.f, fa), __fadd_rn(__fmul_rn(b1, (float)p[0]), __fmul_rn(b2, (float)p[pitch])));
                for (int j = 0; j < pw; ++j) {
                    const float t = __fadd_rn(__fmul_rn(a12, (float)p[j + 1]), __fmul_rn(a22, (float)p[j + 1 + pitch]));
                    sub[r * pw + j] = __fadd_rn(prev, t);
                    prev = (float)((double)t * s);
                }
            }
        } else {
            const float fa = __fsub_rn(centerx, (float)ipx), fb = __fsub_rn(centery, (float)ipy);
            const float na = __fsub_rn(1.f, fa), nb = __fsub_rn(1.f, fb);
            const float a11 = __fmul_rn(na, nb), a12 = __fmul_rn(fa, nb), a21 = __fmul_rn(na, fb), a22 = __fmul_rn(fa, fb);
            for (int e = lane; e < pw * ph; e += 32) {
                const int r = e / pw, j = e - r * pw;
                int y0 = ipy + r, y1 = y0 + 1, x0 = ipx + j, x1 = x0 + 1;
                y0 = min(max(y0, 0), h - 1); y1 = min(max(y1, 0), h - 1);
                const bool xin = x0 >= 0 && x1 <= w - 1;
                x0 = min(max(x0, 0), w - 1); x1 = min(max(x1, 0), w - 1);
                const uint8_t* r0 = src + (size_t)y0 * pitch; const uint8_t* r1 = src + (size_t)y1 * pitch;
                float val;
                if (xin)
                    val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)r0[x0], a11), __fmul_rn((float)r0[x1], a12)),
                                              __fmul_rn((float)r1[x0], a21)), __fmul_rn((float)r1[x1], a22));
                else
                    val = __fadd_rn(__fmul_rn((float)r0[x0], nb), __fmul_rn((float)r1[x0], fb));
                sub[e] = val;
            }
        }
        __syncwarp();
        // ---- normal equations.  Terms: every lane, one window pixel per step; sums: lanes 0..4, raster order
        const int npx = ww * wh;
        double* tm = s_term[warp];
        for (int e = lane; e < npx; e += 32) {
            const int r = e / ww, j = e - r * ww;
            const float* sp = sub + (r + 1) * pw + 1;
            const double m = (double)__fmul_rn(a.ey[r], a.ex[j]);
            const double tgx = (double)__fsub_rn(sp[j + 1], sp[j - 1]);
            const double tgy = (double)__fsub_rn(sp[j + pw], sp[j - pw]);
            const double px = (double)(j - a.win_w), py = (double)(r - a.win_h);
            const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m), gxy = __dmul_rn(__dmul_rn(tgx, tgy), m),
                         gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
            tm[e] = gxx; tm[npx + e] = gxy; tm[2 * npx + e] = gyy;
            tm[3 * npx + e] = __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py));
            tm[4 * npx + e] = __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py));
        }
        __syncwarp();
        double acc = 0;
        if (lane < 5) {
            const double* t = tm + lane * npx;
#pragma unroll 4
            for (int e = 0; e < npx; ++e) acc = __dadd_rn(acc, t[e]);
        }
        const double sa = __shfl_sync(0xffffffffu, acc, 0), sb = __shfl_sync(0xffffffffu, acc, 1), sc = __shfl_sync(0xffffffffu, acc, 2),
                     bb1 = __shfl_sync(0xffffffffu, acc, 3), bb2 = __shfl_sync(0xffffffffu, acc, 4);
        const double det = __dsub_rn(__dmul_rn(sa, sc), __dmul_rn(sb, sb));
        if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
        const double scale = __ddiv_rn(1.0, det);
        const float nx = (float)__dsub_rn(__dadd_rn((double)cix, __dmul_rn(__dmul_rn(sc, scale), bb1)), __dmul_rn(__dmul_rn(sb, scale), bb2));
        const float ny = (float)__dadd_rn(__dsub_rn((double)ciy, __dmul_rn(__dmul_rn(sb, scale), bb1)), __dmul_rn(__dmul_rn(sa, scale), bb2));
        const float ex = __fsub_rn(nx, cix), ey = __fsub_rn(ny, ciy);
        err = (double)__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
        // an update that would leave the image is dropped (cv2 4.13 behaviour, pinned by the oracle fixtures)
        if (nx < 0.f || nx >= (float)w || ny < 0.f || ny >= (float)h) break;
        cix = nx; ciy = ny;
    } while (++iter < a.max_iters && err > a.eps2);
    // too far from the start = poor convergence: keep the input
    if (fabsf(__fsub_rn(cix, cT.x)) > (float)a.win_w || fabsf(__fsub_rn(ciy, cT.y)) > (float)a.win_h) { cix = cT.x; ciy = cT.y; }
    if (lane == 0) a.xy[(size_t)img * a.cap + i] = make_float2(cix, ciy);
}

extern "C" zs_status zs_corner_subpix(zs_context* ctx, const zs_pyramid* p, int first, int count, float* d_xy, const int* d_count,
                                      int cap, int win_w, int win_h, int max_iters, double epsilon)
{
    ZS_REQUIRE(ctx && p && d_xy && d_count, "null argument");
    ZS_REQUIRE(count >= 0 && count <= p->slots && first >= 0 && cap > 0, "bad range");
    ZS_REQUIRE(win_w >= 1 && win_h >= 1 && win_w <= SUBPIX_MAX_WIN && win_h <= SUBPIX_MAX_WIN, "half window must be within 1..7");
    if (count == 0) return ZS_OK;
    subpix_args a;
    a.v = p->v; a.first = first; a.xy = (float2*)d_xy; a.count = d_count; a.cap = cap; a.win_w = win_w; a.win_h = win_h;
    a.max_iters = max_iters < 1 ? 1 : max_iters > 100 ? 100 : max_iters;          // cv: MIN(MAX(maxCount, 1), 100)
    const double eps = epsilon < 0 ? 0 : epsilon;
    a.eps2 = eps * eps;
    // OpenCV: float y = (float)(i - win)/win; float vy = std::exp(-y*y)  (expf on the host)
    for (int i = 0; i < 2 * win_w + 1; ++i) { const float x = (float)(i - win_w) / win_w; a.ex[i] = expf(-x * x); }
    for (int i = 0; i < 2 * win_h + 1; ++i) { const float y = (float)(i - win_h) / win_h; a.ey[i] = expf(-y * y); }
    k_corner_subpix<<<dim3(zs_div_up(cap, SUBPIX_WARPS), count), SUBPIX_WARPS * 32, 0, ctx->stream>>>(a);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}
