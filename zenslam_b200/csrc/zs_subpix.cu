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

// ------------------------------------------------------------------------------------------------------
// v2: six points per warp.  The serial part of the algorithm -- 121 dependent double additions per sum, five sums per
// point and iteration -- is what bounds v1 (one point per warp: five busy lanes, a latency-bound chain; ncu: 41 % issue
// slots used, 13.5 resident warps per SM).  Here lane 5 g + s carries sum s of point slot g (g = 0..5), so the chain phase
// advances thirty sums per instruction, and everything that is order-free is spread over all 32 lanes:
//   * the getRectSubPix patch: inside the image the CPU's row recurrence has no carried dependency -- prev_j is a function
//     of t_{j-1} alone -- so every patch pixel is computed on its own from four image pixels (same operations, same
//     roundings);
//   * the five terms of every window pixel (gxx, gxy, gyy, gxx px + gxy py, gxy px + gyy py -- the very expressions v1
//     evaluates), two window rows of all six points at a time, parked in shared memory (row stride odd, so that the thirty
//     chain lanes spread over the banks); the chain lanes then only load and add, in raster order.
// A point slot that has converged takes the warp's next point at the next iteration (warp-local queue over a contiguous
// range of (image, corner) items), so lanes do not idle behind the slowest of six.  Results are bit-identical to v1.
// ------------------------------------------------------------------------------------------------------
#define SPX_G 6
#define SPX_ROWS 2                              // window rows of terms parked per chunk (the term loop below assumes 2)

// exact u8 -> f32 on the ALU / FMA pipes (I2F issues on the quarter-rate XU pipe): 2^23 + v is exactly representable
__device__ __forceinline__ float spx_u8f(unsigned v) { return __fsub_rn(__uint_as_float(0x4b000000u | v), 8388608.f); }

__device__ __forceinline__ int spx_div(int e, int rcp) { return (int)(((unsigned)e * (unsigned)rcp) >> 16); }   // e / d for e < 1024, d <= 32, rcp = 65536 / d + 1

__global__ void __launch_bounds__(SUBPIX_WARPS * 32) k_corner_subpix_v2(subpix_args a, int total_items, int items_per_warp)
{
    extern __shared__ __align__(16) uint8_t spx_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = a.v.w[0], h = a.v.h[0], pitch = a.v.pitch[0];
    const int ww = 2 * a.win_w + 1, wh = 2 * a.win_h + 1, pw = ww + 2, ph = wh + 2, npatch = pw * ph;
    const int chunk = (SPX_ROWS * ww) | 1;                                  // row stride of the term arrays: terms per point and chunk,
                                                                            // made odd so that the thirty chain lanes spread over all banks
    const size_t warp_bytes = (size_t)SPX_G * (5 * chunk * sizeof(double) + ((npatch * sizeof(float) + 7) & ~(size_t)7));
    double* s_mask = (double*)spx_smem;                                     // [wh * ww]: (double)(ey[r] * ex[j]), shared by the CTA
    const int mask_doubles = (ww * wh + 1) & ~1;
    double* s_pxd = s_mask + mask_doubles;                                  // [ww]: (double)(j - win_w);  [wh]: (double)(r - win_h)
    double* s_pyd = s_pxd + ww;
    const int table_doubles = (mask_doubles + ww + wh + 1) & ~1;
    double* s_term = (double*)(spx_smem + table_doubles * sizeof(double) + warp * warp_bytes);  // [G][5][chunk]
    float* s_sub = (float*)(s_term + SPX_G * 5 * chunk);                    // [G][patch_stride]
    const int patch_stride = (int)(((npatch * sizeof(float) + 7) & ~(size_t)7) / sizeof(float));
    const int rcp_pw = 65536 / pw + 1, rcp_ww = 65536 / ww + 1;
    const int g = lane / 5, s = lane - 5 * g;                               // lanes 30, 31: g = 6, helpers of the 32-wide phases only
    const int leader = g < SPX_G ? 5 * g : 0;
    const uint8_t* base = a.v.img[0] + (size_t)a.v.pad_y * pitch + a.v.pad_x;

    for (int e = threadIdx.x; e < ww * wh; e += SUBPIX_WARPS * 32) {
        const int r = spx_div(e, rcp_ww), j = e - r * ww;
        s_mask[e] = (double)__fmul_rn(a.ey[r], a.ex[j]);
    }
    for (int e = threadIdx.x; e < ww + wh; e += SUBPIX_WARPS * 32) s_pxd[e] = e < ww ? (double)(e - a.win_w) : (double)(e - ww - a.win_h);
    __syncthreads();

    int next = (blockIdx.x * SUBPIX_WARPS + warp) * items_per_warp;
    const int end = min(next + items_per_warp, total_items);
    bool active = false;
    int item = 0, iter = 0, slot = 0;
    float cix = 0.f, ciy = 0.f, c0x = 0.f, c0y = 0.f;

    for (;;) {
        // ---- refill idle point slots from the warp's range (an item beyond its image's corner count is skipped)
        for (int tries = 0; tries < 3; ++tries) {
            const unsigned idle = __ballot_sync(0xffffffffu, !active && s == 0 && g < SPX_G);
            if (!idle || next >= end) break;
            const int mine = next + __popc(idle & ((1u << lane) - 1));
            const int id = __shfl_sync(0xffffffffu, mine, leader);
            next += __popc(idle);
            if (!active && g < SPX_G && id < end) {
                const int img = id / a.cap, i = id - img * a.cap;
                if (i < min(a.count[img], a.cap)) {
                    const float2 c = a.xy[id];
                    item = id; slot = zs_slot(a.first, img, a.v.slots); cix = c0x = c.x; ciy = c0y = c.y; iter = 0; active = true;
                }
            }
        }
        const unsigned act = __ballot_sync(0xffffffffu, active && s == 0);        // bit 5 g = point slot g is iterating
        if (!act) {
            if (next >= end) break;                                              // the range is exhausted too
            continue;                                                            // only skipped items so far: keep refilling
        }

        // ---- patches (getRectSubPix_8u32f).  The per-point part -- integer origin, fractions, the double (1 - a) / a of the row
        // recurrence -- is computed by every point slot for itself at once; the pixels then take all 32 lanes, slot after slot
        const float my_cx = __fsub_rn(cix, __fmul_rn((float)(pw - 1), 0.5f)), my_cy = __fsub_rn(ciy, __fmul_rn((float)(ph - 1), 0.5f));
        const int my_ipx = __float2int_rd(my_cx), my_ipy = __float2int_rd(my_cy);
        const bool my_inside = 0 <= my_ipx && my_ipx + pw < w && 0 <= my_ipy && my_ipy + ph < h;
        float my_fa = __fsub_rn(my_cx, (float)my_ipx);
        const float my_fb = __fsub_rn(my_cy, (float)my_ipy);
        if (my_inside) my_fa = fmaxf(my_fa, 0.0001f);
        const double my_sc = (1. - (double)my_fa) / (double)my_fa;              // used by the inside form only (fa >= 1e-4 there)
        const unsigned inside = __ballot_sync(0xffffffffu, my_inside);
        for (int gg = 0; gg < SPX_G; ++gg) {
            if (!(act >> (5 * gg) & 1)) continue;
            const int ipx = __shfl_sync(0xffffffffu, my_ipx, 5 * gg), ipy = __shfl_sync(0xffffffffu, my_ipy, 5 * gg);
            const float fa = __shfl_sync(0xffffffffu, my_fa, 5 * gg), fb = __shfl_sync(0xffffffffu, my_fb, 5 * gg);
            const uint8_t* src = base + (size_t)__shfl_sync(0xffffffffu, slot, 5 * gg) * a.v.slot_stride[0];
            float* sub = s_sub + gg * patch_stride;
            if (inside >> (5 * gg) & 1) {
                const double sc = __shfl_sync(0xffffffffu, my_sc, 5 * gg);
                const float a12 = __fmul_rn(fa, __fsub_rn(1.f, fb)), a22 = __fmul_rn(fa, fb), b1 = __fsub_rn(1.f, fb), b2 = fb;
                const uint8_t* org = src + (ptrdiff_t)ipy * pitch + ipx;
                for (int e = lane; e < npatch; e += 32) {
                    const int r = spx_div(e, rcp_pw), j = e - r * pw;
                    const uint8_t* p = org + (r * pitch + j);
                    const float p0 = spx_u8f(p[0]), p1 = spx_u8f(p[pitch]);
                    const float t = __fadd_rn(__fmul_rn(a12, spx_u8f(p[1])), __fmul_rn(a22, spx_u8f(p[1 + pitch])));
                    float prev;
                    if (j == 0) prev = __fmul_rn(__fsub_rn(1.f, fa), __fadd_rn(__fmul_rn(b1, p0), __fmul_rn(b2, p1)));
                    else prev = (float)((double)__fadd_rn(__fmul_rn(a12, p0), __fmul_rn(a22, p1)) * sc);
                    sub[e] = __fadd_rn(prev, t);
                }
            } else {
                const float na = __fsub_rn(1.f, fa), nb = __fsub_rn(1.f, fb);
                const float a11 = __fmul_rn(na, nb), a12 = __fmul_rn(fa, nb), a21 = __fmul_rn(na, fb), a22 = __fmul_rn(fa, fb);
                for (int e = lane; e < npatch; e += 32) {
                    const int r = spx_div(e, rcp_pw), j = e - r * pw;
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
        }
        __syncwarp();

        // ---- normal equations: SPX_ROWS window rows at a time -- products by all lanes, then thirty ordered sums in step
        double acc = 0;
        const double* ts = s_term + ((g < SPX_G ? g : 0) * 5 + s) * chunk;                  // this lane's sum: terms of point slot g, row s
        for (int r0 = 0; r0 < wh; r0 += SPX_ROWS) {
            const int rows = min(SPX_ROWS, wh - r0), n = rows * ww;
            // element idx = gg * n + e walks the six slots' terms 32 at a time: (gg, e) advance without divisions
            int gg = spx_div(lane, 65536 / n + 1), e = lane - gg * n;
            for (int idx = lane; idx < SPX_G * n; idx += 32, e += 32) {
                while (e >= n) { e -= n; ++gg; }
                if (!(act >> (5 * gg) & 1)) continue;
                int rr = 0, j = e;
                if (j >= ww) { j -= ww; rr = 1; }                            // SPX_ROWS == 2
                const int r = r0 + rr;
                const float* sp = s_sub + gg * patch_stride + (r + 1) * pw + 1 + j;
                const double m = s_mask[r * ww + j];
                const double tgx = (double)__fsub_rn(sp[1], sp[-1]);
                const double tgy = (double)__fsub_rn(sp[pw], sp[-pw]);
                const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m), gxy = __dmul_rn(__dmul_rn(tgx, tgy), m),
                             gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
                const double px = s_pxd[j], py = s_pyd[r];
                double* tm = s_term + gg * 5 * chunk + e;
                tm[0] = gxx; tm[chunk] = gxy; tm[2 * chunk] = gyy;
                tm[3 * chunk] = __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py));
                tm[4 * chunk] = __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py));
            }
            __syncwarp();
            if (active && g < SPX_G) {
#pragma unroll 4
                for (int e = 0; e < n; ++e) acc = __dadd_rn(acc, ts[e]);
            }
            __syncwarp();
        }

        // ---- 2x2 solve and update, every lane of a point slot on the same numbers
        const double sa = __shfl_sync(0xffffffffu, acc, leader), sb = __shfl_sync(0xffffffffu, acc, leader + 1),
                     scc = __shfl_sync(0xffffffffu, acc, leader + 2), bb1 = __shfl_sync(0xffffffffu, acc, leader + 3),
                     bb2 = __shfl_sync(0xffffffffu, acc, leader + 4);
        if (active && g < SPX_G) {
            bool done = false;
            const double det = __dsub_rn(__dmul_rn(sa, scc), __dmul_rn(sb, sb));
            if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) done = true;
            else {
                const double scale = __ddiv_rn(1.0, det);
                const float nx = (float)__dsub_rn(__dadd_rn((double)cix, __dmul_rn(__dmul_rn(scc, scale), bb1)), __dmul_rn(__dmul_rn(sb, scale), bb2));
                const float ny = (float)__dadd_rn(__dsub_rn((double)ciy, __dmul_rn(__dmul_rn(sb, scale), bb1)), __dmul_rn(__dmul_rn(sa, scale), bb2));
                const float ex = __fsub_rn(nx, cix), ey = __fsub_rn(ny, ciy);
                const double err = (double)__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                // an update that would leave the image is dropped (cv2 4.13 behaviour, pinned by the oracle fixtures)
                if (nx < 0.f || nx >= (float)w || ny < 0.f || ny >= (float)h) done = true;
                else {
                    cix = nx; ciy = ny;
                    if (!(++iter < a.max_iters && err > a.eps2)) done = true;
                }
            }
            if (done) {
                // too far from the start = poor convergence: keep the input
                if (fabsf(__fsub_rn(cix, c0x)) > (float)a.win_w || fabsf(__fsub_rn(ciy, c0y)) > (float)a.win_h) { cix = c0x; ciy = c0y; }
                if (s == 0) a.xy[item] = make_float2(cix, ciy);
                active = false;
            }
        }
    }
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
    if (ctx->sw.subpix_v1) {
        k_corner_subpix<<<dim3(zs_div_up(cap, SUBPIX_WARPS), count), SUBPIX_WARPS * 32, 0, ctx->stream>>>(a);
        ZS_LAUNCH_CHECK(ctx);
        return ZS_OK;
    }
    // six points per warp; a warp's range of (image, corner) items is sized so that one wave of resident warps covers the call
    const long long total = (long long)count * cap;
    ZS_REQUIRE(total < (1LL << 31), "count * cap too large");
    const int ww = 2 * win_w + 1, wh = 2 * win_h + 1, npatch = (ww + 2) * (wh + 2);
    const size_t warp_bytes = (size_t)SPX_G * (5 * ((SPX_ROWS * ww) | 1) * sizeof(double) + ((npatch * sizeof(float) + 7) & ~(size_t)7));
    const size_t smem = warp_bytes * SUBPIX_WARPS + (size_t)(((((ww * wh + 1) & ~1) + ww + wh) + 1) & ~1) * sizeof(double);
    if (smem > 48 * 1024) ZS_CUDA(cudaFuncSetAttribute(k_corner_subpix_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one wave: the kernel is latency-bound per warp (a warp runs its points to convergence), so a second, partly filled
    // wave would cost as much as the first (measured: 822 CTAs on 740 resident slots = 1.20 instead of 1.02 ms)
    int ctas_per_sm = 0;
    ZS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_corner_subpix_v2, SUBPIX_WARPS * 32, smem));
    const long long resident_warps = (long long)ctx->sm_count * (ctas_per_sm > 0 ? ctas_per_sm : 1) * SUBPIX_WARPS;
    int ipw = (int)((total + resident_warps - 1) / resident_warps);
    ipw = ipw < SPX_G ? SPX_G : ipw;
    const int warps = (int)((total + ipw - 1) / ipw);
    k_corner_subpix_v2<<<zs_div_up(warps, SUBPIX_WARPS), SUBPIX_WARPS * 32, smem, ctx->stream>>>(a, (int)total, ipw);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}
