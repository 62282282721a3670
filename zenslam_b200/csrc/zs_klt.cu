// zs_klt.cu -- pyramidal Lucas-Kanade: pyr_lk::calc_optical_flow_pyr_lk == cv::calcOpticalFlowPyrLK
// (zenslam_core/include/zenslam/tracking/pyr_lk.h:15-26, zenslam_core/source/tracking/pyr_lk.cpp:25) and the
// forward-backward gate of keypoint_tracker::track_keypoints (keypoint_tracker.cpp:129-197, 343-434).
// Arithmetic follows SURVEY A.8: 14-bit fixed-point bilinear weights, integer template (I, Ix, Iy) and
// mismatch sums accumulated EXACTLY in integers, one conversion to float32, float32 2x2 solve with the
// reference's operation order (compiled with -fmad=false; every product/sum below is a separate rounding).
//
// One warp per feature; all pyramid levels and all Gauss-Newton iterations run inside the kernel (no
// per-level launch).  The template patch lives in shared memory; the warp sweeps the window with its 32
// lanes, control flow (early exits, iteration counts) is warp-uniform so divergence is across warps only.
#include <math.h>
#include <stdlib.h>

#include "zs_common.cuh"

#define KLT_WARPS 8

struct klt_args {
    zs_pyr_view v;
    const int* prev_slot; const int* next_slot;
    const float2* prev_pts; float2* next_pts; const int* count;
    const int* pts_row;      // optional: row of prev_pts / count each job reads (null: row = job)
    // optional second target per job (template sharing): job j also tracks its points into slot next_slot2[j] and
    // writes those results to the rows of job out_job2[j] (< 0: none).  A job whose prev_slot is < 0 is skipped.
    const int* next_slot2; const int* out_job2;
    const int* job_list;     // optional: blockIdx.y indexes this list of jobs to run (the rest were folded into second targets)
    int cap, win_w, win_h, max_level, max_iters, flags;
    double eps2, min_eig;
    float eps2_lo, eps2_hi;  // float band around eps2 inside which the double comparison is evaluated
    uint8_t* status; float* err;
    int fb; double fb_thr; uint8_t* keep;
    // persistent launch of the tiled kernel: CTAs take (job, point) items from this counter until n_items are gone
    int* work; int n_items, n_jobs_listed;
};

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int lo = __shfl_xor_sync(0xffffffffu, (int)(v & 0xffffffffll), o);
        const int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), o);
        v += ((long long)hi << 32) | (unsigned int)lo;
    }
    return v;
}

__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11)
{
    const float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), 16384.f));
    w11 = 16384 - w00 - w01 - w10;
}

struct lk_result { float x, y; int status; float err; };

// Track one point from plane set (I, dI) to J over all levels.  Warp-uniform.
template <bool BIG>
__device__ __forceinline__ lk_result lk_track_point(const klt_args& a, int slot_i, int slot_j, float2 prev, float2 init,
                                                     bool use_init, short* sI, short2* sD, int lane)
{
    const zs_pyr_view& v = a.v;
    const int ww = a.win_w, wh = a.win_h, npx = ww * wh;
    const float hwx = __fmul_rn((float)(ww - 1), 0.5f), hwy = __fmul_rn((float)(wh - 1), 0.5f);
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    lk_result r; r.status = 1; r.err = 0.f; r.x = 0.f; r.y = 0.f;
    float outx = 0.f, outy = 0.f;
    const int top = min(a.max_level, v.levels - 1);

    for (int level = top; level >= 0; --level) {
        const int cols = v.w[level], rows = v.h[level], pitch = v.pitch[level];
        const size_t org = (size_t)v.pad_y * pitch + v.pad_x;
        const uint8_t* I = v.img[level] + (size_t)slot_i * v.slot_stride[level] + org;
        const short2* dI = v.der[level] + (size_t)slot_i * v.slot_stride[level] + org;
        const uint8_t* J = v.img[level] + (size_t)slot_j * v.slot_stride[level] + org;
        const float scale = 1.f / (float)(1 << level);
        float px = __fmul_rn(prev.x, scale), py = __fmul_rn(prev.y, scale);
        float nx, ny;
        if (level == top) {
            if (use_init) { nx = __fmul_rn(init.x, scale); ny = __fmul_rn(init.y, scale); }
            else { nx = px; ny = py; }
        } else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;

        px = __fsub_rn(px, hwx); py = __fsub_rn(py, hwy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (ipx < -ww || ipx >= cols || ipy < -wh || ipy >= rows) {
            if (level == 0) { r.status = 0; r.err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        // ---- template: I, Ix, Iy over the window; A = sum of gradient products
        long long sA11 = 0, sA12 = 0, sA22 = 0;
        {
            int pA11 = 0, pA12 = 0, pA22 = 0;
            int x = lane, y = 0;
            while (x >= ww) { x -= ww; ++y; }
            const uint8_t* Ib = I + (ptrdiff_t)ipy * pitch + ipx;
            const short2* dIb = dI + (ptrdiff_t)ipy * pitch + ipx;
            for (int idx = lane; idx < npx; idx += 32) {
                const uint8_t* s0 = Ib + y * pitch + x;
                const short2* d0 = dIb + y * pitch + x;
                const int ival = ((int)s0[0] * w00 + (int)s0[1] * w01 + (int)s0[pitch] * w10 + (int)s0[pitch + 1] * w11 + (1 << 8)) >> 9;
                const short2 a00 = d0[0], a01 = d0[1], a10 = d0[pitch], a11 = d0[pitch + 1];
                const int ixv = ((int)a00.x * w00 + (int)a01.x * w01 + (int)a10.x * w10 + (int)a11.x * w11 + (1 << 13)) >> 14;
                const int iyv = ((int)a00.y * w00 + (int)a01.y * w01 + (int)a10.y * w10 + (int)a11.y * w11 + (1 << 13)) >> 14;
                sI[idx] = (short)ival;
                sD[idx] = make_short2((short)ixv, (short)iyv);
                if (BIG) { sA11 += ixv * ixv; sA12 += ixv * iyv; sA22 += iyv * iyv; }
                else { pA11 += ixv * ixv; pA12 += ixv * iyv; pA22 += iyv * iyv; }
                x += 32;
                while (x >= ww) { x -= ww; ++y; }
            }
            if (!BIG) { sA11 = pA11; sA12 = pA12; sA22 = pA22; }
        }
        sA11 = warp_sum_ll(sA11); sA12 = warp_sum_ll(sA12); sA22 = warp_sum_ll(sA22);
        __syncwarp();
        const float A11 = __fmul_rn(__ll2float_rn(sA11), FLT_SCALE), A12 = __fmul_rn(__ll2float_rn(sA12), FLT_SCALE),
                    A22 = __fmul_rn(__ll2float_rn(sA22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * ww * wh));
        if (a.flags & ZS_LK_GET_MIN_EIGENVALS) r.err = minEig;
        if ((double)minEig < a.min_eig || D < 1.1920929e-07f) {
            if (level == 0) r.status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, hwx); ny = __fsub_rn(ny, hwy);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < a.max_iters; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -ww || inx >= cols || iny < -wh || iny >= rows) {
                if (level == 0) r.status = 0;
                break;
            }
            lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            long long sb1 = 0, sb2 = 0;
            {
                int pb1 = 0, pb2 = 0;
                int x = lane, y = 0;
                while (x >= ww) { x -= ww; ++y; }
                const uint8_t* Jb = J + (ptrdiff_t)iny * pitch + inx;
                for (int idx = lane; idx < npx; idx += 32) {
                    const uint8_t* j0 = Jb + y * pitch + x;
                    const int diff = (((int)j0[0] * w00 + (int)j0[1] * w01 + (int)j0[pitch] * w10 + (int)j0[pitch + 1] * w11 + (1 << 8)) >> 9)
                                     - (int)sI[idx];
                    const short2 g = sD[idx];
                    if (BIG) { sb1 += diff * (int)g.x; sb2 += diff * (int)g.y; }
                    else { pb1 += diff * (int)g.x; pb2 += diff * (int)g.y; }
                    x += 32;
                    while (x >= ww) { x -= ww; ++y; }
                }
                if (!BIG) { sb1 = pb1; sb2 = pb2; }
            }
            sb1 = warp_sum_ll(sb1); sb2 = warp_sum_ll(sb2);
            const float b1 = __fmul_rn(__ll2float_rn(sb1), FLT_SCALE), b2 = __fmul_rn(__ll2float_rn(sb2), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, hwx); outy = __fadd_rn(ny, hwy);
            if ((double)dx * (double)dx + (double)dy * (double)dy <= a.eps2) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        __syncwarp();
    }
    r.x = outx; r.y = outy;
    return r;
}


// ------------------------------------------------------------------------------------------------------
// v2: specialised warp-per-feature tracker for windows up to 32 columns wide (compile-time WW x WH).
//   * lane = window column, rows fully unrolled; the gradient template (Ix, Iy) lives in REGISTERS;
//   * bilinear sampling with two dp2a per pixel: weights packed s16x2, the horizontally adjacent u8 pair
//     packed in the low half of a register, the pair of the row below rolled over from the previous row;
//   * the mismatch sums are formed as  sum(J*Ix) - sum(I*Ix)  (the second term is a per-level constant of
//     the template), which is the same integer as sum((J - I)*Ix) and drops I from the inner loop;
//   * exact integer warp reductions with REDUX (split into 16-bit halves so 32 lanes cannot overflow).
// Same arithmetic, same results, ~3.5x fewer issued instructions than the generic kernel below.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp2a_lo_su(int a_s16x2, unsigned b_u8, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_u8), "r"(c));
    return d;
}

__device__ __forceinline__ long long warp_sum_exact(int p)
{
    const int lo = p & 0xffff, hi = p >> 16;
    const int slo = __reduce_add_sync(0xffffffffu, lo);
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    return ((long long)shi << 16) + (long long)slo;
}

__device__ __forceinline__ int pack_s16x2(int lo, int hi) { return (lo & 0xffff) | (hi << 16); }

template <int WW, int WH>
__device__ __forceinline__ lk_result lk_track_point_v2(const klt_args& a, int slot_i, int slot_j, float2 prev, float2 init,
                                                        bool use_init, int lane, int* sG)
{
    const zs_pyr_view& v = a.v;
    const float hwx = (float)(WW - 1) * 0.5f, hwy = (float)(WH - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    lk_result r; r.status = 1; r.err = 0.f; r.x = 0.f; r.y = 0.f;
    float outx = 0.f, outy = 0.f;
    const int top = min(a.max_level, v.levels - 1);
    const bool active = lane < WW;

    for (int level = top; level >= 0; --level) {
        const int cols = v.w[level], rows = v.h[level], pitch = v.pitch[level];
        const size_t org = (size_t)v.pad_y * pitch + v.pad_x;
        const uint8_t* I = v.img[level] + (size_t)slot_i * v.slot_stride[level] + org;
        const short2* dI = v.der[level] + (size_t)slot_i * v.slot_stride[level] + org;
        const uint8_t* J = v.img[level] + (size_t)slot_j * v.slot_stride[level] + org;
        const float scale = 1.f / (float)(1 << level);
        float px = __fmul_rn(prev.x, scale), py = __fmul_rn(prev.y, scale);
        float nx, ny;
        if (level == top) {
            if (use_init) { nx = __fmul_rn(init.x, scale); ny = __fmul_rn(init.y, scale); }
            else { nx = px; ny = py; }
        } else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;

        px = __fsub_rn(px, hwx); py = __fsub_rn(py, hwy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (ipx < -WW || ipx >= cols || ipy < -WH || ipy >= rows) {
            if (level == 0) { r.status = 0; r.err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        // ---- template: rolled row loop (small code), (Ix,Iy) handed over through shared memory as packed
        // s16x2 words, then unpacked once into registers for the fully unrolled iteration loop
        int pA11 = 0, pA12 = 0, pA22 = 0, pc1 = 0, pc2 = 0;
        {
            const int wt = pack_s16x2(w00, w01), wb = pack_s16x2(w10, w11);
            const uint8_t* ip = I + (ptrdiff_t)ipy * pitch + ipx + lane;
            const short2* dp = dI + (ptrdiff_t)ipy * pitch + ipx + lane;
            unsigned pc = (unsigned)ip[0] | ((unsigned)ip[1] << 8);
            short2 c0 = dp[0], c1 = dp[1];
#pragma unroll 1
            for (int y = 0; y < WH; ++y) {
                ip += pitch; dp += pitch;
                const unsigned pn = (unsigned)ip[0] | ((unsigned)ip[1] << 8);
                const short2 n0 = dp[0], n1 = dp[1];
                const int ival = dp2a_lo_su(wt, pc, dp2a_lo_su(wb, pn, 1 << 8)) >> 9;
                int ixv = ((int)c0.x * w00 + (int)c1.x * w01 + (int)n0.x * w10 + (int)n1.x * w11 + (1 << 13)) >> 14;
                int iyv = ((int)c0.y * w00 + (int)c1.y * w01 + (int)n0.y * w10 + (int)n1.y * w11 + (1 << 13)) >> 14;
                if (!active) { ixv = 0; iyv = 0; }
                sG[y * 32 + lane] = pack_s16x2(ixv, iyv);
                pA11 += ixv * ixv; pA12 += ixv * iyv; pA22 += iyv * iyv;
                pc1 += ival * ixv; pc2 += ival * iyv;
                pc = pn; c0 = n0; c1 = n1;
            }
        }
        int Ix[WH], Iy[WH];
#pragma unroll
        for (int y = 0; y < WH; ++y) {
            const int g = sG[y * 32 + lane];       // written by this same lane: no synchronisation needed
            Ix[y] = (int)(short)(g & 0xffff); Iy[y] = g >> 16;
        }
        const long long sA11 = warp_sum_exact(pA11), sA12 = warp_sum_exact(pA12), sA22 = warp_sum_exact(pA22);
        const long long sc1 = warp_sum_exact(pc1), sc2 = warp_sum_exact(pc2);
        const float A11 = __fmul_rn(__ll2float_rn(sA11), FLT_SCALE), A12 = __fmul_rn(__ll2float_rn(sA12), FLT_SCALE),
                    A22 = __fmul_rn(__ll2float_rn(sA22), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * WW * WH));
        if (a.flags & ZS_LK_GET_MIN_EIGENVALS) r.err = minEig;
        if ((double)minEig < a.min_eig || D < 1.1920929e-07f) {
            if (level == 0) r.status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, hwx); ny = __fsub_rn(ny, hwy);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < a.max_iters; ++j) {
            const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
            if (inx < -WW || inx >= cols || iny < -WH || iny >= rows) {
                if (level == 0) r.status = 0;
                break;
            }
            lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            const int wt = pack_s16x2(w00, w01), wb = pack_s16x2(w10, w11);
            const uint8_t* jp = J + (ptrdiff_t)iny * pitch + inx + lane;
            unsigned pc = (unsigned)jp[0] | ((unsigned)jp[1] << 8);
            int pb1 = 0, pb2 = 0;
#pragma unroll
            for (int y = 0; y < WH; ++y) {
                jp += pitch;
                const unsigned pn = (unsigned)jp[0] | ((unsigned)jp[1] << 8);
                const int jval = dp2a_lo_su(wt, pc, dp2a_lo_su(wb, pn, 1 << 8)) >> 9;
                pb1 += jval * Ix[y]; pb2 += jval * Iy[y];
                pc = pn;
            }
            const long long sb1 = warp_sum_exact(pb1) - sc1, sb2 = warp_sum_exact(pb2) - sc2;
            const float b1 = __fmul_rn(__ll2float_rn(sb1), FLT_SCALE), b2 = __fmul_rn(__ll2float_rn(sb2), FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, hwx); outy = __fadd_rn(ny, hwy);
            if ((double)dx * (double)dx + (double)dy * (double)dy <= a.eps2) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
    }
    r.x = outx; r.y = outy;
    return r;
}

#define KLT2_WARPS 4

// grid: (ceil(cap / KLT2_WARPS), jobs); static smem: WH x 32 words per warp
template <int WW, int WH>
__global__ void __launch_bounds__(KLT2_WARPS * 32) k_klt_track_v2(klt_args a)
{
    __shared__ int s_g[KLT2_WARPS][WH * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job = blockIdx.y;
    const int i = blockIdx.x * KLT2_WARPS + warp;
    const int in_row = a.pts_row ? a.pts_row[job] : job;
    if (i >= min(a.count[in_row], a.cap)) return;
    const size_t o = (size_t)job * a.cap + i;
    const float2 p0 = a.prev_pts[(size_t)in_row * a.cap + i];
    bool use_init = (a.flags & ZS_LK_USE_INITIAL_FLOW) != 0;
    float2 init = make_float2(0.f, 0.f);
    if (use_init) init = a.next_pts[o];
    int si = a.prev_slot[job], sj = a.next_slot[job];
    float2 from = p0;
    int fwd_status = 1;
    // pass 0 = forward call; pass 1 (fb only) = backward call from the forward result without initial flow
    // (keypoint_tracker.cpp:156-170).  One inlined instance of the tracker keeps the code inside the I-cache.
#pragma unroll 1
    for (int pass = 0; pass < (a.fb ? 2 : 1); ++pass) {
        const lk_result f = lk_track_point_v2<WW, WH>(a, si, sj, from, init, use_init, lane, s_g[warp]);
        if (pass == 0) {
            if (lane == 0) {
                a.next_pts[o] = make_float2(f.x, f.y);
                a.status[o] = (uint8_t)f.status;
                a.err[o] = f.err;
            }
            fwd_status = f.status;
            from = make_float2(f.x, f.y);
            const int t = si; si = sj; sj = t;
            use_init = false;
        } else if (lane == 0) {
            // cv::norm(Point2f) -> sqrt((double)dx*dx + (double)dy*dy) < klt_threshold (keypoint_tracker.cpp:180)
            const float dx = __fsub_rn(f.x, p0.x), dy = __fsub_rn(f.y, p0.y);
            const double nrm = sqrt((double)dx * (double)dx + (double)dy * (double)dy);
            a.keep[o] = (uint8_t)(fwd_status && f.status && nrm < a.fb_thr);
        }
    }
}


// ------------------------------------------------------------------------------------------------------
// v3/v4: TMA-staged patches.  Per warp, one elected lane asks the TMA unit for the patch of J (and, once per
// level, the patch of I plus the (dx,dy) patch) around the window's integer origin; completion is signalled on
// an mbarrier and no LSU instruction or register is spent on the gather.  TMA box origins must be 16-byte
// aligned (an unaligned origin faults with "illegal instruction" on sm_100a), so the boxes are 48 bytes /
// 36 words wide and start at the aligned column below the window; the residual offset is applied when the
// lanes read shared memory (word offset folded into the address, byte offset folded into the funnel shifts
// that form the (x, x+1) byte pairs anyway).  The J patch is fetched again only when the window leaves the
// 16-byte block or changes row (most Gauss-Newton steps move < 1 px).
//   lane = (k = lane & 3: columns 8k..8k+7, g = lane >> 2: rows g, g+8, g+16, g+24): 4 x 8 pixels per lane.
// Arithmetic and results are identical to v2 / the generic kernel.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int x, int y, int z, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
__device__ __forceinline__ int dp2a_hi_su(int a_s16x2, unsigned b_u8, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_u8), "r"(c));
    return d;
}

// One warp per 32x32 window tile, one feature per CTA: a 31x31 window is one tile (one-warp CTAs, 24 per SM: with 32-thread
// CTAs ptxas fits the tracker in 80 registers instead of 125, so 24 warps instead of 16 hide the serial tail of every
// Gauss-Newton step -- REDUX, 2x2 solve, TMA waits; measured 13.25 -> 12.07 ms per 128-frame batch; 20 CTAs: 12.14 ms; 28:
// 13.3 ms; 32: 14.2 ms (spills)).  A 63x63 window (tumvi.yaml:45) is 2x2 tiles = four warps that each run the same code on
// their own TMA-staged patch and add their exact integer partial sums through shared memory (one CTA barrier per sum set,
// double-buffered); every warp then repeats the scalar tail on identical numbers, so control flow stays CTA-uniform.
#define KLT4_JP 12                   // J / I patch row pitch in words (48-byte TMA box)
#define KLT4_DP 36                   // derivative patch row pitch in words
#define KLT4_ROWS 33                 // TMA box rows: 32 window rows + the bilinear row below
#define KLT4_SJ_BYTES 1664           // 33 rows x 48 B = 1584, rounded to a 128-byte multiple
#define KLT4_SD_BYTES 4864           // 33 rows x 144 B = 4752, rounded to a 128-byte multiple
#define KLT4_WARP_BYTES (KLT4_SJ_BYTES + KLT4_SD_BYTES + 128)

template <int WW, int WH>
struct klt4_cfg {
    static constexpr int TX = (WW + 31) / 32, TY = (WH + 31) / 32, NT = TX * TY;       // tiles = warps per CTA
    static constexpr int LW = WW - 32 * (TX - 1), LH = WH - 32 * (TY - 1);             // size of the last tile column / row
    static constexpr int JB = TY > 1 ? 4 : (WH + 7) / 8;                               // 8-row bands a tile sweeps
};

// the eight (x, x+1) byte pairs of a lane's row: R0 = bytes 0..3, F0 = bytes 1..4, R1 = bytes 4..7, F1 = bytes 5..8
struct row8 { unsigned R0, F0, R1, F1; };

// w points at the word holding byte 0 (byte offset t = 0..3 inside it); sh = 8 t
__device__ __forceinline__ row8 load_row8(const uint32_t* w, int sh)
{
    const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
    row8 r;
    r.R0 = __funnelshift_r(w0, w1, sh); r.F0 = __funnelshift_rc(w0, w1, sh + 8);
    r.R1 = __funnelshift_r(w1, w2, sh); r.F1 = __funnelshift_rc(w1, w2, sh + 8);
    return r;
}

// bilinear sample of pixel P (0..7) from the row pair (top, bottom)
template <int P>
__device__ __forceinline__ int sample8(int wt, int wb, const row8& a, const row8& b)
{
    const unsigned ta = (P < 4) ? ((P & 1) ? a.F0 : a.R0) : ((P & 1) ? a.F1 : a.R1);
    const unsigned tb = (P < 4) ? ((P & 1) ? b.F0 : b.R0) : ((P & 1) ? b.F1 : b.R1);
    int acc;
    if ((P & 3) < 2) acc = dp2a_lo_su(wt, ta, dp2a_lo_su(wb, tb, 1 << 8));
    else acc = dp2a_hi_su(wt, ta, dp2a_hi_su(wb, tb, 1 << 8));
    // (the shift as IMAD.HI by 2^23 would move it from the ALU to the FMA pipe, but IMAD.HI issues far slower than SHF: KLT 11.86
    // -> 13.64 ms at C2, 13.87 -> 15.55 ms at TUMVI; gpurun_out/r5i)
    return acc >> 9;
}

// exact sum of N per-warp totals over the NT warps of the CTA (identity for one-warp CTAs).  red = [2][NT][N] in shared
// memory; the buffer alternates so that one barrier per call suffices (a warp can only overwrite a buffer after every
// warp has passed the barrier of the call in between, i.e. after every warp has read it).
template <int NT, int N>
__device__ __forceinline__ void cta_sum(long long (&v)[N], long long* red, int& phase, int warp, int lane)
{
    if (NT == 1) return;
    long long* buf = red + phase * (NT * 5);
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) buf[warp * 5 + n] = v[n];
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        long long t = 0;
#pragma unroll
        for (int w = 0; w < NT; ++w) t += buf[w * 5 + n];
        v[n] = t;
    }
    phase ^= 1;
}

// Up to two targets per source point: the per-level template (I patch, gradients, A matrix) depends only on the
// source image and point, so jobs that track the SAME keypoints of the SAME image into two different images
// (stereo L->R of frame k and temporal L_k -> L_{k+1}) share it and only the Gauss-Newton loops run twice.
struct lk_state { float outx, outy; int status; };

template <int WW, int WH, int PK>
__device__ __forceinline__ void lk_track_point_v4(const klt_args& a, int slot_i, int slot_j0, int slot_j1, int ntgt,
                                                  float2 prev, float2 init, bool use_init, int warp, int lane, uint8_t* sJ,
                                                  uint8_t* sD, uint32_t bar, uint32_t& parity, long long* red, int& phase,
                                                  lk_state& t0, lk_state& t1, float& err)
{
    typedef klt4_cfg<WW, WH> cfg;
    const zs_pyr_view& v = a.v;
    const float hwx = (float)(WW - 1) * 0.5f, hwy = (float)(WH - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    t0.status = 1; t1.status = 1; t0.outx = t0.outy = t1.outx = t1.outy = 0.f; err = 0.f;
    const int top = min(a.max_level, v.levels - 1);
    const int k = lane & 3, g = lane >> 2;
    // this warp's tile: origin offset inside the window and valid size (compile-time for single-tile windows)
    const int tx = cfg::TX == 1 ? 0 : warp % cfg::TX, ty = cfg::TY == 1 ? 0 : warp / cfg::TX;
    const int tw = (cfg::TX == 1 || tx == cfg::TX - 1) ? cfg::LW : 32, th = (cfg::TY == 1 || ty == cfg::TY - 1) ? cfg::LH : 32;
    const uint32_t sJ_a = smem_u32(sJ), sD_a = smem_u32(sD);
    const char* maps = (const char*)v.tmaps;
    const uint32_t* jbase = (const uint32_t*)sJ + g * KLT4_JP + 2 * k;
    const uint32_t* dbase = (const uint32_t*)sD + g * KLT4_DP + 8 * k;

    for (int level = top; level >= 0; --level) {
        const int cols = v.w[level], rows = v.h[level];
        const float scale = 1.f / (float)(1 << level);
        float px = __fmul_rn(prev.x, scale), py = __fmul_rn(prev.y, scale);
        if (level == top) {
            if (use_init) { t0.outx = __fmul_rn(init.x, scale); t0.outy = __fmul_rn(init.y, scale); }
            else { t0.outx = px; t0.outy = py; }
            t1.outx = px; t1.outy = py;                       // second targets never carry an initial flow
        } else {
            t0.outx = __fmul_rn(t0.outx, 2.f); t0.outy = __fmul_rn(t0.outy, 2.f);
            t1.outx = __fmul_rn(t1.outx, 2.f); t1.outy = __fmul_rn(t1.outy, 2.f);
        }

        px = __fsub_rn(px, hwx); py = __fsub_rn(py, hwy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (ipx < -WW || ipx >= cols || ipy < -WH || ipy >= rows) {
            if (level == 0) { t0.status = 0; t1.status = 0; err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        // ---- fetch the I patch (into the J buffer) and the derivative patch of this warp's tile
        const int gx = v.pad_x + ipx + 32 * tx, gy = v.pad_y + ipy + 32 * ty;      // padded-plane coordinates, gx >= 1
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(bar, (48 + 36 * 4) * KLT4_ROWS);
            tma_load_3d(sJ_a, maps + (size_t)(2 * level) * 128, gx & ~15, gy, slot_i, bar);
            tma_load_3d(sD_a, maps + (size_t)(2 * level + 1) * 128, gx & ~3, gy, slot_i, bar);
        }
        mbar_wait(bar, parity); parity ^= 1;

        // ---- template in registers: JB row bands x 8 pixels per lane
        int Ix[cfg::JB][PK == 1 ? 4 : 8], Iy[cfg::JB][PK == 1 ? 4 : 8];      // PK: s16 x 2 words of pixel pairs (see lk_track_point_v5)
        int pA11 = 0, pA12 = 0, pA22 = 0, pc1 = 0, pc2 = 0;
        {
            const int wt = pack_s16x2(w00, w01), wb = pack_s16x2(w10, w11);
            const uint32_t* jw = jbase + ((gx & 15) >> 2);
            const int sh = (gx & 3) * 8;
            const uint32_t* dw = dbase + (gx & 3);
#pragma unroll
            for (int j = 0; j < cfg::JB; ++j) {
                const row8 ra = load_row8(jw + j * 8 * KLT4_JP, sh), rb = load_row8(jw + (j * 8 + 1) * KLT4_JP, sh);
                int tx_[9], ty_[9], bx[9], by[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const unsigned dt = dw[j * 8 * KLT4_DP + q], db = dw[(j * 8 + 1) * KLT4_DP + q];
                    tx_[q] = (int)(short)(dt & 0xffff); ty_[q] = (int)dt >> 16;
                    bx[q] = (int)(short)(db & 0xffff); by[q] = (int)db >> 16;
                }
                const bool row_masked = (8 * j + 7 >= cfg::LH) && (8 * j + g >= th);
                int iv[8];
                iv[0] = sample8<0>(wt, wb, ra, rb); iv[1] = sample8<1>(wt, wb, ra, rb); iv[2] = sample8<2>(wt, wb, ra, rb);
                iv[3] = sample8<3>(wt, wb, ra, rb); iv[4] = sample8<4>(wt, wb, ra, rb); iv[5] = sample8<5>(wt, wb, ra, rb);
                iv[6] = sample8<6>(wt, wb, ra, rb); iv[7] = sample8<7>(wt, wb, ra, rb);
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) {
                    int ixv = (tx_[pp] * w00 + tx_[pp + 1] * w01 + bx[pp] * w10 + bx[pp + 1] * w11 + (1 << 13)) >> 14;
                    int iyv = (ty_[pp] * w00 + ty_[pp + 1] * w01 + by[pp] * w10 + by[pp + 1] * w11 + (1 << 13)) >> 14;
                    if ((8 * j + 7 >= cfg::LH) || (24 + pp >= cfg::LW)) {      // compile-time: only rows / columns some tile masks
                        bool masked = row_masked;
                        if (24 + pp >= cfg::LW) masked = masked || (8 * k + pp >= tw);
                        if (masked) { ixv = 0; iyv = 0; }
                    }
                    if constexpr (PK == 1) {
                        if (pp & 1) { Ix[j][pp >> 1] |= ixv << 16; Iy[j][pp >> 1] |= iyv << 16; }
                        else { Ix[j][pp >> 1] = ixv & 0xffff; Iy[j][pp >> 1] = iyv & 0xffff; }
                    } else { Ix[j][pp] = ixv; Iy[j][pp] = iyv; }
                    pA11 += ixv * ixv; pA12 += ixv * iyv; pA22 += iyv * iyv;
                    pc1 += iv[pp] * ixv; pc2 += iv[pp] * iyv;
                }
            }
        }
        long long sT[5] = { warp_sum_exact(pA11), warp_sum_exact(pA12), warp_sum_exact(pA22), warp_sum_exact(pc1), warp_sum_exact(pc2) };
        cta_sum<cfg::NT, 5>(sT, red, phase, warp, lane);
        const long long sc1 = sT[3], sc2 = sT[4];
        const float A11 = __fmul_rn(__ll2float_rn(sT[0]), FLT_SCALE), A12 = __fmul_rn(__ll2float_rn(sT[1]), FLT_SCALE),
                    A22 = __fmul_rn(__ll2float_rn(sT[2]), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * WW * WH));
        if (a.flags & ZS_LK_GET_MIN_EIGENVALS) err = minEig;
        if ((double)minEig < a.min_eig || D < 1.1920929e-07f) {
            if (level == 0) { t0.status = 0; t1.status = 0; }
            continue;
        }
        D = __fdiv_rn(1.f, D);

        // ---- Gauss-Newton loops, one per target (a single copy of the loop body: the state is swapped in and out)
#pragma unroll 1
        for (int t = 0; t < ntgt; ++t) {
            const int slot_j = t ? slot_j1 : slot_j0;
            float outx = t ? t1.outx : t0.outx, outy = t ? t1.outy : t0.outy;
            int status = 1;
            float nx = __fsub_rn(outx, hwx), ny = __fsub_rn(outy, hwy);
            float pdx = 0.f, pdy = 0.f;
            int cur_x0 = 0x7fffffff, cur_y = 0x7fffffff;       // origin of the J patch now in shared memory
            for (int it = 0; it < a.max_iters; ++it) {
                const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
                if (inx < -WW || inx >= cols || iny < -WH || iny >= rows) {
                    status = 0;
                    break;
                }
                const int jx = v.pad_x + inx + 32 * tx, jy = v.pad_y + iny + 32 * ty;
                if ((jx & ~15) != cur_x0 || jy != cur_y) {
                    __syncwarp();                                  // every lane is done with the previous patch
                    cur_x0 = jx & ~15; cur_y = jy;
                    if (lane == 0) {
                        mbar_expect_tx(bar, 48 * KLT4_ROWS);
                        tma_load_3d(sJ_a, maps + (size_t)(2 * level) * 128, cur_x0, cur_y, slot_j, bar);
                    }
                    mbar_wait(bar, parity); parity ^= 1;
                }
                lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
                const int wt = pack_s16x2(w00, w01), wb = pack_s16x2(w10, w11);
                const uint32_t* jw = jbase + ((jx & 15) >> 2);
                const int sh = (jx & 3) * 8;
                int pb1 = 0, pb2 = 0;
                if constexpr (PK == 1) {
                    int xl = 0, xh = 0, yl = 0, yh = 0;
#pragma unroll
                    for (int j = 0; j < cfg::JB; ++j) {
                        const row8 ra = load_row8(jw + j * 8 * KLT4_JP, sh), rb = load_row8(jw + (j * 8 + 1) * KLT4_JP, sh);
                        unsigned bp;
                        bp = __byte_perm(sample8<0>(wt, wb, ra, rb), sample8<1>(wt, wb, ra, rb), 0x5140);
                        xl = dp2a_lo_su(Ix[j][0], bp, xl); xh = dp2a_hi_su(Ix[j][0], bp, xh);
                        yl = dp2a_lo_su(Iy[j][0], bp, yl); yh = dp2a_hi_su(Iy[j][0], bp, yh);
                        bp = __byte_perm(sample8<2>(wt, wb, ra, rb), sample8<3>(wt, wb, ra, rb), 0x5140);
                        xl = dp2a_lo_su(Ix[j][1], bp, xl); xh = dp2a_hi_su(Ix[j][1], bp, xh);
                        yl = dp2a_lo_su(Iy[j][1], bp, yl); yh = dp2a_hi_su(Iy[j][1], bp, yh);
                        bp = __byte_perm(sample8<4>(wt, wb, ra, rb), sample8<5>(wt, wb, ra, rb), 0x5140);
                        xl = dp2a_lo_su(Ix[j][2], bp, xl); xh = dp2a_hi_su(Ix[j][2], bp, xh);
                        yl = dp2a_lo_su(Iy[j][2], bp, yl); yh = dp2a_hi_su(Iy[j][2], bp, yh);
                        bp = __byte_perm(sample8<6>(wt, wb, ra, rb), sample8<7>(wt, wb, ra, rb), 0x5140);
                        xl = dp2a_lo_su(Ix[j][3], bp, xl); xh = dp2a_hi_su(Ix[j][3], bp, xh);
                        yl = dp2a_lo_su(Iy[j][3], bp, yl); yh = dp2a_hi_su(Iy[j][3], bp, yh);
                    }
                    pb1 = xl + (xh << 8); pb2 = yl + (yh << 8);
                } else {
#pragma unroll
                for (int j = 0; j < cfg::JB; ++j) {
                    const row8 ra = load_row8(jw + j * 8 * KLT4_JP, sh), rb = load_row8(jw + (j * 8 + 1) * KLT4_JP, sh);
                    int jv;
                    jv = sample8<0>(wt, wb, ra, rb); pb1 += jv * Ix[j][0]; pb2 += jv * Iy[j][0];
                    jv = sample8<1>(wt, wb, ra, rb); pb1 += jv * Ix[j][1]; pb2 += jv * Iy[j][1];
                    jv = sample8<2>(wt, wb, ra, rb); pb1 += jv * Ix[j][2]; pb2 += jv * Iy[j][2];
                    jv = sample8<3>(wt, wb, ra, rb); pb1 += jv * Ix[j][3]; pb2 += jv * Iy[j][3];
                    jv = sample8<4>(wt, wb, ra, rb); pb1 += jv * Ix[j][4]; pb2 += jv * Iy[j][4];
                    jv = sample8<5>(wt, wb, ra, rb); pb1 += jv * Ix[j][5]; pb2 += jv * Iy[j][5];
                    jv = sample8<6>(wt, wb, ra, rb); pb1 += jv * Ix[j][6]; pb2 += jv * Iy[j][6];
                    jv = sample8<7>(wt, wb, ra, rb); pb1 += jv * Ix[j][7]; pb2 += jv * Iy[j][7];
                }
                }
                long long sB[2] = { warp_sum_exact(pb1), warp_sum_exact(pb2) };
                cta_sum<cfg::NT, 2>(sB, red, phase, warp, lane);
                const long long sb1 = sB[0] - sc1, sb2 = sB[1] - sc2;
                const float b1 = __fmul_rn(__ll2float_rn(sb1), FLT_SCALE), b2 = __fmul_rn(__ll2float_rn(sb2), FLT_SCALE);
                const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
                const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
                nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
                outx = __fadd_rn(nx, hwx); outy = __fadd_rn(ny, hwy);
                // OpenCV: (double)dx*dx + (double)dy*dy <= eps^2.  A float estimate decides unless it falls inside a
                // +-1e-6 relative band around eps^2 (its own error is < 2e-7), where the exact double form is evaluated.
                const float ss = fmaf(dx, dx, __fmul_rn(dy, dy));
                bool conv = ss <= a.eps2_lo;
                if (!conv && ss < a.eps2_hi) conv = (double)dx * (double)dx + (double)dy * (double)dy <= a.eps2;
                if (conv) break;
                // OpenCV: fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01 in double; for a float v, v < 0.01 <=> v <= 0.01f
                // (0.01f = 0.00999999978 is the largest float below 0.01)
                if (it > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                    outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (t) { t1.outx = outx; t1.outy = outy; if (level == 0 && !status) t1.status = 0; }
            else { t0.outx = outx; t0.outy = outy; if (level == 0 && !status) t0.status = 0; }
        }
    }
}

// grid: (cap, jobs); block = one warp per window tile; dynamic smem = tiles * KLT4_WARP_BYTES
// passes: 0 = forward (one or two targets), 1 = backward of target 0, 2 = backward of target 1 (fb only); one
// inlined instance of the tracker serves all passes (keeps the code inside the instruction cache).
// ------------------------------------------------------------------------------------------------------
// v5: TWO window tiles per warp (63x63 windows: 2 x 2 tiles of 32 x 32 = two warps per feature, warp w owning the tile
// row w).  In v4 every one of the four warps of a feature repeats the scalar part of each Gauss-Newton step -- bounds and
// refetch tests, weights, reductions, the 2x2 solve, the convergence tests: 38 % of an iteration's instructions -- on
// identical numbers.  With two tiles per warp that part runs twice per feature instead of four times, the two tiles'
// pixel loops are independent instruction streams the scheduler can interleave, and the two tiles' lane partials share
// their REDUX (the 16-bit halves of two partials still cannot overflow).  The price is registers: both tiles' gradient
// templates live in registers (128 of them), so a CTA of two warps needs ~200 registers per thread and five CTAs fit an SM.
// Arithmetic and results are identical to v4.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_sum_exact2(int pa, int pb)
{
    const int lo = (pa & 0xffff) + (pb & 0xffff), hi = (pa >> 16) + (pb >> 16);
    const int slo = __reduce_add_sync(0xffffffffu, lo);
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    return ((long long)shi << 16) + (long long)slo;
}

// PK = packed template: the gradients of a lane's pixel PAIRS as s16 x 2 words (half the template registers).  The products
// then come from IDP.2A as well: jv < 2^13 is split into its two bytes, a PRMT puts (lo_p, lo_q, hi_p, hi_q) of a pixel pair
// into one word, and IDP.2A.LO / .HI accumulate sum(lo * I) and sum(hi * I) -- sum(jv * I) = lo-sum + 256 hi-sum exactly.
// 4 IDP + 1 PRMT per pixel pair instead of 4 IMAD: 4 more instructions per 8-pixel band, 64 registers fewer per thread.
template <int WW, int WH, int PK>
__device__ __forceinline__ void lk_track_point_v5(const klt_args& a, int slot_i, int slot_j0, int slot_j1, int ntgt,
                                                  float2 prev, float2 init, bool use_init, int warp, int lane, uint8_t* tiles,
                                                  uint32_t bar, uint32_t& parity, long long* red, int& phase,
                                                  lk_state& t0, lk_state& t1, float& err)
{
    typedef klt4_cfg<WW, WH> cfg;
    static_assert(cfg::TX == 2 && cfg::TY == 2, "two tiles per warp: 2 x 2 tile windows");
    constexpr int TPW = 2, NW = 2;                           // tiles per warp (the tile row of warp `warp`), warps per CTA
    constexpr int TILE_BYTES = KLT4_SJ_BYTES + KLT4_SD_BYTES;
    const zs_pyr_view& v = a.v;
    const float hwx = (float)(WW - 1) * 0.5f, hwy = (float)(WH - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    t0.status = 1; t1.status = 1; t0.outx = t0.outy = t1.outx = t1.outy = 0.f; err = 0.f;
    const int top = min(a.max_level, v.levels - 1);
    const int k = lane & 3, g = lane >> 2;
    const int ty = warp;                                     // tile row; tile column = t (compile-time in the unrolled loops)
    const int th = (ty == cfg::TY - 1) ? cfg::LH : 32;
    const uint32_t tiles_a = smem_u32(tiles);
    const char* maps = (const char*)v.tmaps;
    const uint32_t* jbase = (const uint32_t*)tiles + g * KLT4_JP + 2 * k;
    const uint32_t* dbase = (const uint32_t*)(tiles + KLT4_SJ_BYTES) + g * KLT4_DP + 8 * k;

    for (int level = top; level >= 0; --level) {
        const int cols = v.w[level], rows = v.h[level];
        const float scale = 1.f / (float)(1 << level);
        float px = __fmul_rn(prev.x, scale), py = __fmul_rn(prev.y, scale);
        if (level == top) {
            if (use_init) { t0.outx = __fmul_rn(init.x, scale); t0.outy = __fmul_rn(init.y, scale); }
            else { t0.outx = px; t0.outy = py; }
            t1.outx = px; t1.outy = py;
        } else {
            t0.outx = __fmul_rn(t0.outx, 2.f); t0.outy = __fmul_rn(t0.outy, 2.f);
            t1.outx = __fmul_rn(t1.outx, 2.f); t1.outy = __fmul_rn(t1.outy, 2.f);
        }
        px = __fsub_rn(px, hwx); py = __fsub_rn(py, hwy);
        const int ipx = __float2int_rd(px), ipy = __float2int_rd(py);
        if (ipx < -WW || ipx >= cols || ipy < -WH || ipy >= rows) {
            if (level == 0) { t0.status = 0; t1.status = 0; err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);

        // ---- fetch the I patches (into the J buffers) and the derivative patches of this warp's two tiles
        const int gx = v.pad_x + ipx, gy = v.pad_y + ipy + 32 * ty;          // tile column t adds 32 t to gx: same alignment residue
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(bar, TPW * (48 + 36 * 4) * KLT4_ROWS);
#pragma unroll
            for (int t = 0; t < TPW; ++t) {
                tma_load_3d(tiles_a + t * TILE_BYTES, maps + (size_t)(2 * level) * 128, (gx + 32 * t) & ~15, gy, slot_i, bar);
                tma_load_3d(tiles_a + t * TILE_BYTES + KLT4_SJ_BYTES, maps + (size_t)(2 * level + 1) * 128, (gx + 32 * t) & ~3, gy, slot_i, bar);
            }
        }
        mbar_wait(bar, parity); parity ^= 1;

        // ---- templates in registers: 2 tiles x 4 row bands x 8 pixels per lane
        int Ix[TPW][4][PK != 0 ? 4 : 8], Iy[TPW][4][PK != 0 ? 4 : 8];
        long long sT[5];
        if constexpr (PK >= 2) {
            // ROLLED template pass (round 2, late).  Unrolled over 2 tiles x 4 bands the template section is 1 920 instructions
            // (30 KB) of straight-line code that every warp streams through once per level: with sixteen warps per SM at
            // different places in it the instruction cache thrashes (ncu: `no_instruction` 1.25 warps per issue, 65 % of those
            // samples in this section).  Here ONE band body runs eight times; a band's packed gradients (8 words per lane) are
            // parked in the derivative rows the band has just consumed (rows 8j .. 8j+7 of its tile, dead from then on; row
            // 8j+8 is still needed by the next band and stays) and loaded back into registers, where the Gauss-Newton loops
            // want them, by sixteen 128-bit loads afterwards.  Both tiles' lane partials share one 32-bit word: 64 pixels x
            // 4080 x 8160 < 2^31.
            int pA11 = 0, pA12 = 0, pA22 = 0, pc1 = 0, pc2 = 0;
            const int wt = pack_s16x2(w00, w01), wb = pack_s16x2(w10, w11);
            const int sh = (gx & 3) * 8;
            const uint32_t* jw = jbase + ((gx & 15) >> 2);
            const uint32_t* dw = dbase + (gx & 3);
            uint32_t* park = (uint32_t*)(tiles + KLT4_SJ_BYTES) + 4 * lane;
#pragma unroll 1
            for (int tj = 0; tj < TPW * 4; ++tj) {
                const int t = tj >> 2, j = tj & 3;
                const row8 ra = load_row8(jw, sh), rb = load_row8(jw + KLT4_JP, sh);
                int tx_[9], ty_[9], bx[9], by[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const unsigned dt = dw[q], db = dw[KLT4_DP + q];
                    tx_[q] = (int)(short)(dt & 0xffff); ty_[q] = (int)dt >> 16;
                    bx[q] = (int)(short)(db & 0xffff); by[q] = (int)db >> 16;
                }
                const bool row_masked = (8 * j + 7 >= cfg::LH) && (8 * j + g >= th);
                int iv[8];
                iv[0] = sample8<0>(wt, wb, ra, rb); iv[1] = sample8<1>(wt, wb, ra, rb); iv[2] = sample8<2>(wt, wb, ra, rb);
                iv[3] = sample8<3>(wt, wb, ra, rb); iv[4] = sample8<4>(wt, wb, ra, rb); iv[5] = sample8<5>(wt, wb, ra, rb);
                iv[6] = sample8<6>(wt, wb, ra, rb); iv[7] = sample8<7>(wt, wb, ra, rb);
                unsigned gxp[4], gyp[4];
#pragma unroll
                for (int pp = 0; pp < 8; ++pp) {
                    int ixv = (tx_[pp] * w00 + tx_[pp + 1] * w01 + bx[pp] * w10 + bx[pp + 1] * w11 + (1 << 13)) >> 14;
                    int iyv = (ty_[pp] * w00 + ty_[pp + 1] * w01 + by[pp] * w10 + by[pp + 1] * w11 + (1 << 13)) >> 14;
                    bool masked = row_masked;
                    if (24 + pp >= cfg::LW) masked = masked || (t == cfg::TX - 1 && 8 * k + pp >= cfg::LW);
                    if (masked) { ixv = 0; iyv = 0; }
                    if (pp & 1) { gxp[pp >> 1] |= (unsigned)ixv << 16; gyp[pp >> 1] |= (unsigned)iyv << 16; }
                    else { gxp[pp >> 1] = ixv & 0xffff; gyp[pp >> 1] = iyv & 0xffff; }
                    pA11 += ixv * ixv; pA12 += ixv * iyv; pA22 += iyv * iyv;
                    pc1 += iv[pp] * ixv; pc2 += iv[pp] * iyv;
                }
                __syncwarp();                                  // every lane has read this band's derivative rows
                uint32_t* st = park + t * (TILE_BYTES / 4) + j * (8 * KLT4_DP);
                *(uint4*)st = make_uint4(gxp[0], gxp[1], gxp[2], gxp[3]);
                *(uint4*)(st + 128) = make_uint4(gyp[0], gyp[1], gyp[2], gyp[3]);
                jw += 8 * KLT4_JP; dw += 8 * KLT4_DP;
                if (j == 3) { jw += TILE_BYTES / 4 - 32 * KLT4_JP; dw += TILE_BYTES / 4 - 32 * KLT4_DP; }
            }
#pragma unroll
            for (int t = 0; t < TPW; ++t)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t* st = park + t * (TILE_BYTES / 4) + j * (8 * KLT4_DP);
                    const uint4 gx4 = *(const uint4*)st, gy4 = *(const uint4*)(st + 128);
                    Ix[t][j][0] = gx4.x; Ix[t][j][1] = gx4.y; Ix[t][j][2] = gx4.z; Ix[t][j][3] = gx4.w;
                    Iy[t][j][0] = gy4.x; Iy[t][j][1] = gy4.y; Iy[t][j][2] = gy4.z; Iy[t][j][3] = gy4.w;
                }
            sT[0] = warp_sum_exact(pA11); sT[1] = warp_sum_exact(pA12); sT[2] = warp_sum_exact(pA22);
            sT[3] = warp_sum_exact(pc1); sT[4] = warp_sum_exact(pc2);
        } else {
            int pA11[TPW], pA12[TPW], pA22[TPW], pc1[TPW], pc2[TPW];
            const int wt = pack_s16x2(w00, w01), wb = pack_s16x2(w10, w11);
            const int sh = (gx & 3) * 8;
#pragma unroll
            for (int t = 0; t < TPW; ++t) {
                pA11[t] = pA12[t] = pA22[t] = pc1[t] = pc2[t] = 0;
                const uint32_t* jw = jbase + t * (TILE_BYTES / 4) + ((gx & 15) >> 2);
                const uint32_t* dw = dbase + t * (TILE_BYTES / 4) + (gx & 3);
                constexpr int dummy = 0; (void)dummy;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const row8 ra = load_row8(jw + j * 8 * KLT4_JP, sh), rb = load_row8(jw + (j * 8 + 1) * KLT4_JP, sh);
                    int tx_[9], ty_[9], bx[9], by[9];
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const unsigned dt = dw[j * 8 * KLT4_DP + q], db = dw[(j * 8 + 1) * KLT4_DP + q];
                        tx_[q] = (int)(short)(dt & 0xffff); ty_[q] = (int)dt >> 16;
                        bx[q] = (int)(short)(db & 0xffff); by[q] = (int)db >> 16;
                    }
                    const bool row_masked = (8 * j + 7 >= cfg::LH) && (8 * j + g >= th);
                    int iv[8];
                    iv[0] = sample8<0>(wt, wb, ra, rb); iv[1] = sample8<1>(wt, wb, ra, rb); iv[2] = sample8<2>(wt, wb, ra, rb);
                    iv[3] = sample8<3>(wt, wb, ra, rb); iv[4] = sample8<4>(wt, wb, ra, rb); iv[5] = sample8<5>(wt, wb, ra, rb);
                    iv[6] = sample8<6>(wt, wb, ra, rb); iv[7] = sample8<7>(wt, wb, ra, rb);
#pragma unroll
                    for (int pp = 0; pp < 8; ++pp) {
                        int ixv = (tx_[pp] * w00 + tx_[pp + 1] * w01 + bx[pp] * w10 + bx[pp + 1] * w11 + (1 << 13)) >> 14;
                        int iyv = (ty_[pp] * w00 + ty_[pp + 1] * w01 + by[pp] * w10 + by[pp + 1] * w11 + (1 << 13)) >> 14;
                        // tile column t is the last one for t == 1: its column 31 (pixel 7 of k == 3) lies outside a 63-wide window
                        const bool col_masked = (t == cfg::TX - 1) && (24 + pp >= cfg::LW) && (8 * k + pp >= cfg::LW);
                        if ((8 * j + 7 >= cfg::LH) || ((t == cfg::TX - 1) && (24 + pp >= cfg::LW))) {
                            if (row_masked || col_masked) { ixv = 0; iyv = 0; }
                        }
                        if constexpr (PK != 0) {
                            if (pp & 1) { Ix[t][j][pp >> 1] |= ixv << 16; Iy[t][j][pp >> 1] |= iyv << 16; }
                            else { Ix[t][j][pp >> 1] = ixv & 0xffff; Iy[t][j][pp >> 1] = iyv & 0xffff; }
                        } else { Ix[t][j][pp] = ixv; Iy[t][j][pp] = iyv; }
                        pA11[t] += ixv * ixv; pA12[t] += ixv * iyv; pA22[t] += iyv * iyv;
                        pc1[t] += iv[pp] * ixv; pc2[t] += iv[pp] * iyv;
                    }
                }
            }
            sT[0] = warp_sum_exact2(pA11[0], pA11[1]); sT[1] = warp_sum_exact2(pA12[0], pA12[1]); sT[2] = warp_sum_exact2(pA22[0], pA22[1]);
            sT[3] = warp_sum_exact2(pc1[0], pc1[1]); sT[4] = warp_sum_exact2(pc2[0], pc2[1]);
        }
        cta_sum<NW, 5>(sT, red, phase, warp, lane);
        const long long sc1 = sT[3], sc2 = sT[4];
        const float A11 = __fmul_rn(__ll2float_rn(sT[0]), FLT_SCALE), A12 = __fmul_rn(__ll2float_rn(sT[1]), FLT_SCALE),
                    A22 = __fmul_rn(__ll2float_rn(sT[2]), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * WW * WH));
        if (a.flags & ZS_LK_GET_MIN_EIGENVALS) err = minEig;
        if ((double)minEig < a.min_eig || D < 1.1920929e-07f) {
            if (level == 0) { t0.status = 0; t1.status = 0; }
            continue;
        }
        D = __fdiv_rn(1.f, D);

#pragma unroll 1
        for (int tg = 0; tg < ntgt; ++tg) {
            const int slot_j = tg ? slot_j1 : slot_j0;
            float outx = tg ? t1.outx : t0.outx, outy = tg ? t1.outy : t0.outy;
            int status = 1;
            float nx = __fsub_rn(outx, hwx), ny = __fsub_rn(outy, hwy);
            float pdx = 0.f, pdy = 0.f;
            int cur_x0 = 0x7fffffff, cur_y = 0x7fffffff;       // origin of the J patches now in shared memory (tile column 0)
            for (int it = 0; it < a.max_iters; ++it) {
                const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
                if (inx < -WW || inx >= cols || iny < -WH || iny >= rows) {
                    status = 0;
                    break;
                }
                const int jx = v.pad_x + inx, jy = v.pad_y + iny + 32 * ty;
                if ((jx & ~15) != cur_x0 || jy != cur_y) {
                    __syncwarp();                                  // every lane is done with the previous patches
                    cur_x0 = jx & ~15; cur_y = jy;
                    if (lane == 0) {
                        mbar_expect_tx(bar, TPW * 48 * KLT4_ROWS);
#pragma unroll
                        for (int t = 0; t < TPW; ++t)
                            tma_load_3d(tiles_a + t * TILE_BYTES, maps + (size_t)(2 * level) * 128, cur_x0 + 32 * t, cur_y, slot_j, bar);
                    }
                    mbar_wait(bar, parity); parity ^= 1;
                }
                lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
                const int wt = pack_s16x2(w00, w01), wb = pack_s16x2(w10, w11);
                const int sh = (jx & 3) * 8;
                int pb1[TPW], pb2[TPW];
#pragma unroll
                for (int t = 0; t < TPW; ++t) {
                    pb1[t] = 0; pb2[t] = 0;
                    const uint32_t* jw = jbase + t * (TILE_BYTES / 4) + ((jx & 15) >> 2);
                    if constexpr (PK != 0) {
                        int xl = 0, xh = 0, yl = 0, yh = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const row8 ra = load_row8(jw + j * 8 * KLT4_JP, sh), rb = load_row8(jw + (j * 8 + 1) * KLT4_JP, sh);
                            unsigned bp;
                            bp = __byte_perm(sample8<0>(wt, wb, ra, rb), sample8<1>(wt, wb, ra, rb), 0x5140);
                            xl = dp2a_lo_su(Ix[t][j][0], bp, xl); xh = dp2a_hi_su(Ix[t][j][0], bp, xh);
                            yl = dp2a_lo_su(Iy[t][j][0], bp, yl); yh = dp2a_hi_su(Iy[t][j][0], bp, yh);
                            bp = __byte_perm(sample8<2>(wt, wb, ra, rb), sample8<3>(wt, wb, ra, rb), 0x5140);
                            xl = dp2a_lo_su(Ix[t][j][1], bp, xl); xh = dp2a_hi_su(Ix[t][j][1], bp, xh);
                            yl = dp2a_lo_su(Iy[t][j][1], bp, yl); yh = dp2a_hi_su(Iy[t][j][1], bp, yh);
                            bp = __byte_perm(sample8<4>(wt, wb, ra, rb), sample8<5>(wt, wb, ra, rb), 0x5140);
                            xl = dp2a_lo_su(Ix[t][j][2], bp, xl); xh = dp2a_hi_su(Ix[t][j][2], bp, xh);
                            yl = dp2a_lo_su(Iy[t][j][2], bp, yl); yh = dp2a_hi_su(Iy[t][j][2], bp, yh);
                            bp = __byte_perm(sample8<6>(wt, wb, ra, rb), sample8<7>(wt, wb, ra, rb), 0x5140);
                            xl = dp2a_lo_su(Ix[t][j][3], bp, xl); xh = dp2a_hi_su(Ix[t][j][3], bp, xh);
                            yl = dp2a_lo_su(Iy[t][j][3], bp, yl); yh = dp2a_hi_su(Iy[t][j][3], bp, yh);
                        }
                        pb1[t] = xl + (xh << 8); pb2[t] = yl + (yh << 8);
                    } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const row8 ra = load_row8(jw + j * 8 * KLT4_JP, sh), rb = load_row8(jw + (j * 8 + 1) * KLT4_JP, sh);
                        int jv;
                        jv = sample8<0>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][0]; pb2[t] += jv * Iy[t][j][0];
                        jv = sample8<1>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][1]; pb2[t] += jv * Iy[t][j][1];
                        jv = sample8<2>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][2]; pb2[t] += jv * Iy[t][j][2];
                        jv = sample8<3>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][3]; pb2[t] += jv * Iy[t][j][3];
                        jv = sample8<4>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][4]; pb2[t] += jv * Iy[t][j][4];
                        jv = sample8<5>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][5]; pb2[t] += jv * Iy[t][j][5];
                        jv = sample8<6>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][6]; pb2[t] += jv * Iy[t][j][6];
                        jv = sample8<7>(wt, wb, ra, rb); pb1[t] += jv * Ix[t][j][7]; pb2[t] += jv * Iy[t][j][7];
                    }
                    }
                }
                long long sB[2] = { warp_sum_exact2(pb1[0], pb1[1]), warp_sum_exact2(pb2[0], pb2[1]) };
                cta_sum<NW, 2>(sB, red, phase, warp, lane);
                const long long sb1 = sB[0] - sc1, sb2 = sB[1] - sc2;
                const float b1 = __fmul_rn(__ll2float_rn(sb1), FLT_SCALE), b2 = __fmul_rn(__ll2float_rn(sb2), FLT_SCALE);
                const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
                const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
                nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
                outx = __fadd_rn(nx, hwx); outy = __fadd_rn(ny, hwy);
                const float ss = fmaf(dx, dx, __fmul_rn(dy, dy));
                bool conv = ss <= a.eps2_lo;
                if (!conv && ss < a.eps2_hi) conv = (double)dx * (double)dx + (double)dy * (double)dy <= a.eps2;
                if (conv) break;
                if (it > 0 && fabsf(__fadd_rn(dx, pdx)) <= 0.01f && fabsf(__fadd_rn(dy, pdy)) <= 0.01f) {
                    outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f)); outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (tg) { t1.outx = outx; t1.outy = outy; if (level == 0 && !status) t1.status = 0; }
            else { t0.outx = outx; t0.outy = outy; if (level == 0 && !status) t0.status = 0; }
        }
    }
}


// TPW = window tiles per warp: 1 (one warp per tile, lk_track_point_v4) or 2 (lk_track_point_v5); sJ = the warp's first tile
// buffer, scratch = its 128-byte bookkeeping line (which also holds the mbarrier)
template <int WW, int WH, int TPW, int PK>
__device__ __forceinline__ void klt4_item(const klt_args& a, int job, int i, int warp, int lane, uint8_t* sJ, uint8_t* scratch, uint32_t bar,
                                          uint32_t& parity, long long* s_red, int& phase)
{
    typedef klt4_cfg<WW, WH> cfg;
    uint8_t* sD = sJ + KLT4_SJ_BYTES;
    const int slot_src = a.prev_slot[job];
    if (slot_src < 0) return;                                  // job folded into another job's second target
    const int in_row = a.pts_row ? a.pts_row[job] : job;
    if (i >= min(a.count[in_row], a.cap)) return;
    // Per-warp bookkeeping lives in the 128-byte scratch line behind the barrier (warp-uniform values, written by
    // every lane with the same data, read back as broadcasts): it would otherwise sit in registers across the
    // whole inlined tracker and push the kernel past 128 registers (4 CTAs per SM).
    //   w[4] slot_src  w[5] slot_t0  w[6] slot_t1  w[7] ntgt  w[8..9] p0  w[10..12] f0  w[13..15] f1  w[16] job1
    volatile int* w = (volatile int*)scratch;
    volatile float* wf = (volatile float*)w;
    __syncwarp();
    {
        const int job1 = a.out_job2 ? a.out_job2[job] : -1;
        w[4] = slot_src; w[5] = a.next_slot[job]; w[6] = job1 >= 0 ? a.next_slot2[job] : -1; w[7] = job1 >= 0 ? 2 : 1; w[16] = job1;
        const float2 p0 = a.prev_pts[(size_t)in_row * a.cap + i];
        wf[8] = p0.x; wf[9] = p0.y;
    }
    __syncwarp();
    const bool use_init = (a.flags & ZS_LK_USE_INITIAL_FLOW) != 0;
    const bool writer = lane == 0 && (cfg::NT / TPW == 1 || warp == 0);
    const int passes = a.fb ? 1 + w[7] : 1;
#pragma unroll 1
    for (int pass = 0; pass < passes; ++pass) {
        const int job1 = w[16];
        const size_t o0 = (size_t)job * a.cap + i, o1 = (size_t)(job1 >= 0 ? job1 : job) * a.cap + i;
        // a point whose forward track failed is dropped whatever the backward call returns (keypoint_tracker.cpp:180:
        // status && status_back && ...), and nothing else of the backward call is visible: skip it
        if (pass > 0 && !w[9 + 3 * pass]) {
            if (writer) a.keep[pass == 1 ? o0 : o1] = 0;
            continue;
        }
        // pass 0: forward from p0 (one or two targets); pass 1 / 2: backward of target 0 / 1 from its forward result,
        // images swapped, no initial flow (keypoint_tracker.cpp:156-170)
        const int si = pass == 0 ? w[4] : w[4 + pass], sj0 = pass == 0 ? w[5] : w[4], sj1 = w[6];
        const float2 from = make_float2(wf[pass == 0 ? 8 : 7 + 3 * pass], wf[pass == 0 ? 9 : 8 + 3 * pass]);
        float2 init = make_float2(0.f, 0.f);
        const bool ui = use_init && pass == 0;
        if (ui) init = a.next_pts[(size_t)job * a.cap + i];
        lk_state r0, r1; float err;
        if constexpr (TPW == 1)
            lk_track_point_v4<WW, WH, PK>(a, si, sj0, sj1, pass == 0 ? w[7] : 1, from, init, ui, warp, lane, sJ, sD, bar, parity, s_red,
                                      phase, r0, r1, err);
        else
            lk_track_point_v5<WW, WH, PK>(a, si, sj0, sj1, pass == 0 ? w[7] : 1, from, init, ui, warp, lane, sJ, bar, parity, s_red,
                                      phase, r0, r1, err);
        if (pass == 0) {
            __syncwarp();
            wf[10] = r0.outx; wf[11] = r0.outy; w[12] = r0.status; wf[13] = r1.outx; wf[14] = r1.outy; w[15] = r1.status;
            __syncwarp();
            if (writer) {
                a.next_pts[o0] = make_float2(r0.outx, r0.outy); a.status[o0] = (uint8_t)r0.status; a.err[o0] = err;
                if (job1 >= 0) { a.next_pts[o1] = make_float2(r1.outx, r1.outy); a.status[o1] = (uint8_t)r1.status; a.err[o1] = err; }
            }
        } else if (writer) {
            // cv::norm(Point2f) -> sqrt((double)dx*dx + (double)dy*dy) < klt_threshold (keypoint_tracker.cpp:180)
            const float dx = __fsub_rn(r0.outx, wf[8]), dy = __fsub_rn(r0.outy, wf[9]);
            const double nrm = sqrt((double)dx * (double)dx + (double)dy * (double)dy);
            a.keep[pass == 1 ? o0 : o1] = (uint8_t)(r0.status && nrm < a.fb_thr);
        }
    }
}

// block = one warp per window tile; dynamic smem = tiles * KLT4_WARP_BYTES
// passes: 0 = forward (one or two targets), 1 = backward of target 0, 2 = backward of target 1 (fb only); one
// inlined instance of the tracker serves all passes (keeps the code inside the instruction cache).
// Launch forms: grid (cap, jobs) with one (point, job) item per CTA (a.work == nullptr), or a PERSISTENT grid of
// sm_count x MINB CTAs that take items from an atomic counter, point index fastest (neighbouring CTAs then work on the same
// image pair): no CTA relaunch gap between the ~160 items a CTA slot serves per 128-frame batch, no empty CTAs for the
// points beyond an image's corner count.
template <int WW, int WH, int MINB, int TPW = 1, int PK = 0>
__global__ void __launch_bounds__(klt4_cfg<WW, WH>::NT / TPW * 32, MINB) k_klt_track_v4(klt_args a)
{
    typedef klt4_cfg<WW, WH> cfg;
    constexpr int NW = cfg::NT / TPW;                          // warps per CTA
    constexpr int WARP_BYTES = TPW * (KLT4_SJ_BYTES + KLT4_SD_BYTES) + 128;
    extern __shared__ __align__(128) uint8_t smem3[];
    __shared__ long long s_red[NW > 1 ? 2 * cfg::NT * 5 : 1];
    __shared__ int s_item;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sJ = smem3 + (size_t)warp * WARP_BYTES;
    uint8_t* scratch = sJ + TPW * (KLT4_SJ_BYTES + KLT4_SD_BYTES);
    const uint32_t bar = smem_u32(scratch);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0;
    int phase = 0;
    const bool persistent = a.work != nullptr;
    for (;;) {                                                 // ONE inlined instance of the tracker serves both launch forms
        int jl = blockIdx.y, i = blockIdx.x;
        if (persistent) {
            int item;
            if (NW == 1) {
                item = 0;
                if (lane == 0) item = atomicAdd(a.work, 1);
                item = __shfl_sync(0xffffffffu, item, 0);
            } else {
                __syncthreads();                               // every warp has read the previous item
                if (threadIdx.x == 0) s_item = atomicAdd(a.work, 1);
                __syncthreads();
                item = s_item;
            }
            if (item >= a.n_items) break;
            jl = item / a.cap; i = item - jl * a.cap;
        }
        klt4_item<WW, WH, TPW, PK>(a, a.job_list ? a.job_list[jl] : jl, i, warp, lane, sJ, scratch, bar, parity, s_red, phase);
        if (!persistent) break;
    }
}

// grid: (ceil(cap / KLT_WARPS), jobs); dynamic smem = KLT_WARPS * per-warp template bytes
template <bool BIG>
__global__ void __launch_bounds__(KLT_WARPS * 32) k_klt_track(klt_args a)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job = blockIdx.y;
    const int i = blockIdx.x * KLT_WARPS + warp;
    const int in_row = a.pts_row ? a.pts_row[job] : job;
    if (i >= min(a.count[in_row], a.cap)) return;
    const int npx = a.win_w * a.win_h;
    const size_t per_warp = (size_t)((npx + 1) & ~1) * 2 + (size_t)npx * 4;
    short* sI = (short*)(smem + warp * per_warp);
    short2* sD = (short2*)(smem + warp * per_warp + (size_t)((npx + 1) & ~1) * 2);
    const size_t o = (size_t)job * a.cap + i;
    const float2 p0 = a.prev_pts[(size_t)in_row * a.cap + i];
    const bool use_init = (a.flags & ZS_LK_USE_INITIAL_FLOW) != 0;
    float2 init = make_float2(0.f, 0.f);
    if (use_init) init = a.next_pts[o];
    const int si = a.prev_slot[job], sj = a.next_slot[job];
    const lk_result f = lk_track_point<BIG>(a, si, sj, p0, init, use_init, sI, sD, lane);
    if (lane == 0) {
        a.next_pts[o] = make_float2(f.x, f.y);
        a.status[o] = (uint8_t)f.status;
        a.err[o] = f.err;
    }
    if (a.fb) {
        // backward call (keypoint_tracker.cpp:156-170): from the forward result, no initial flow
        const lk_result b = lk_track_point<BIG>(a, sj, si, make_float2(f.x, f.y), init, false, sI, sD, lane);
        if (lane == 0) {
            // cv::norm(Point2f) -> sqrt((double)dx*dx + (double)dy*dy) < klt_threshold (keypoint_tracker.cpp:180)
            const float dx = __fsub_rn(b.x, p0.x), dy = __fsub_rn(b.y, p0.y);
            const double nrm = sqrt((double)dx * (double)dx + (double)dy * (double)dy);
            a.keep[o] = (uint8_t)(f.status && b.status && nrm < a.fb_thr);
        }
    }
}

// windows the TMA-staged tiled kernel is instantiated for (everything else runs the generic shared-memory kernel)
bool zs_klt_tiled_window(const zs_context* ctx, const zs_pyramid* p, int win_w, int win_h)
{
    if (!p->v.tmaps || ctx->sw.klt_no_tma || win_w != win_h) return false;
    return win_w == 15 || win_w == 21 || win_w == 31 || win_w == 63;
}

zs_status zs_klt_launch(zs_context* ctx, const zs_pyramid* p, const int* d_prev_slot, const int* d_next_slot,
                        const float* d_prev_pts, float* d_next_pts, const int* d_count, const int* d_pts_row, int jobs, int cap,
                        const zs_lk_params* prm, uint8_t* d_status, float* d_err, int fb, double fb_thr, uint8_t* d_keep,
                        const int* d_next_slot2, const int* d_out_job2, const int* d_job_list, int n_list)
{
    ZS_REQUIRE(ctx && p && d_prev_slot && d_next_slot && d_prev_pts && d_next_pts && d_count && prm && d_status && d_err,
               "null argument");
    ZS_REQUIRE(jobs >= 0 && cap > 0, "bad sizes");
    ZS_REQUIRE(prm->win_w == p->win_w && prm->win_h == p->win_h,
               "LK window must equal the window the pyramid was padded for (reference: both are klt_window_size)");
    ZS_REQUIRE(prm->flags & ZS_LK_GET_MIN_EIGENVALS, "only OPTFLOW_LK_GET_MIN_EIGENVALS mode is implemented (the reference's)");
    ZS_REQUIRE(prm->max_level >= 0, "max_level < 0");
    if (jobs == 0) return ZS_OK;
    klt_args a;
    a.v = p->v; a.prev_slot = d_prev_slot; a.next_slot = d_next_slot;
    a.prev_pts = (const float2*)d_prev_pts; a.next_pts = (float2*)d_next_pts; a.count = d_count; a.pts_row = d_pts_row;
    a.cap = cap; a.win_w = prm->win_w; a.win_h = prm->win_h; a.max_level = prm->max_level;
    int mi = prm->max_iters; mi = mi < 0 ? 0 : mi > 100 ? 100 : mi;          // cv: clamp(maxCount, 0, 100)
    double eps = prm->epsilon; eps = eps < 0 ? 0 : eps > 10. ? 10. : eps;   // cv: clamp(epsilon, 0, 10)
    a.max_iters = mi; a.eps2 = eps * eps; a.flags = prm->flags; a.min_eig = prm->min_eig_threshold;
    a.status = d_status; a.err = d_err; a.fb = fb; a.fb_thr = fb_thr; a.keep = d_keep;
    a.next_slot2 = d_next_slot2; a.out_job2 = d_out_job2; a.job_list = nullptr;
    if (a.eps2 < 1e-30) { a.eps2_lo = -1.f; a.eps2_hi = 3.0e38f; }         // always take the exact comparison
    else {
        a.eps2_lo = nextafterf((float)(a.eps2 * (1.0 - 1e-6)), 0.f);
        a.eps2_hi = nextafterf((float)(a.eps2 * (1.0 + 1e-6)), 3.0e38f);
    }
    // TMA-staged tiled kernel: the reference's default 31x31, the shipped 63x63 (tumvi.yaml:45), 21x21 (OpenCV's default), 15x15
    if (zs_klt_tiled_window(ctx, p, a.win_w, a.win_h)) {
        // with a job list only the jobs that still own work get blocks (folded jobs would launch cap empty CTAs each)
        if (d_job_list && n_list > 0) a.job_list = d_job_list;
        const int listed = a.job_list ? n_list : jobs;
        const long long items = (long long)listed * cap;
        a.work = nullptr; a.n_items = 0; a.n_jobs_listed = listed;
        // persistent form when there are many waves of CTAs to run (ZS_KLT_NO_PERSIST: always one CTA per item).  Measured at
        // C2, 128 frames per batch (160 items per CTA slot): KLT 12.01 -> 11.85 ms; at 1 - 16 frames per batch no difference,
        // so small launches keep the plain grid and skip the counter reset
        const long long persist_min = ctx->sw.klt_persist_min > 0 ? ctx->sw.klt_persist_min : 4LL * ctx->sm_count * 24 + 1;
        const bool persist = !ctx->sw.klt_no_persist && items < (1LL << 31) && items >= persist_min;
        if (persist) {
            a.work = ctx->d_klt_work;                          // a device int of the context, zeroed in stream order before the launch
            a.n_items = (int)items;
            ZS_CUDA(cudaMemsetAsync(a.work, 0, sizeof(int), ctx->stream));
        }
#define KLT4_LAUNCH_T(W_, H_, MB_, TPW_) KLT4_LAUNCH_P(W_, H_, MB_, TPW_, 0)
#define KLT4_LAUNCH_P(W_, H_, MB_, TPW_, PK_)                                                                                     \
        do {                                                                                                                   \
            const int nw = klt4_cfg<W_, H_>::NT / TPW_;                                                                         \
            const size_t sm = (size_t)nw * (TPW_ * (KLT4_SJ_BYTES + KLT4_SD_BYTES) + 128);                                     \
            const long long resident = (long long)ctx->sm_count * MB_;                                                         \
            const dim3 grid = persist ? dim3((unsigned)(items < resident ? items : resident), 1) : dim3(cap, listed);          \
            if (!persist) a.work = nullptr;                                                                                    \
            if (sm > 48 * 1024) ZS_CUDA(cudaFuncSetAttribute(k_klt_track_v4<W_, H_, MB_, TPW_, PK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
            k_klt_track_v4<W_, H_, MB_, TPW_, PK_><<<grid, nw * 32, sm, ctx->stream>>>(a);                                      \
        } while (0)
#define KLT4_LAUNCH(W_, H_, MB_) KLT4_LAUNCH_T(W_, H_, MB_, 1)
        // (packed template, ZS_KLT31_PACKED = CTAs per SM: 78 / 72 registers and no spills, but 390 instead of 378 instructions per
        // iteration, all of the extra ones on the ALU pipe: 24 CTAs 11.86 ms, 28: 11.93 ms, 32 (64 registers): 12.22 ms against
        // 11.85 ms -- the 31x31 kernel is bound by issue slots, not by resident warps; profiles/r2_klt_packed_template.txt)
        if (a.win_w == 31 && ctx->sw.klt31_packed == 24) KLT4_LAUNCH_P(31, 31, 24, 1, 1);
        else if (a.win_w == 31 && ctx->sw.klt31_packed == 28) KLT4_LAUNCH_P(31, 31, 28, 1, 1);
        else if (a.win_w == 31) KLT4_LAUNCH(31, 31, 24);
        // (four-warp CTAs: 4 per SM = 128 registers: 16.4 - 16.5 ms; 5 per SM (96 registers, 240 B of spills): 16.6 ms; 6 (80
        // registers, 324 B): 20.1 ms)
        // 63x63 (tumvi.yaml:45): two tiles per warp, two-warp CTAs, six per SM (lk_track_point_v5; 168 registers, 240 B of
        // spills): 15.8 ms per 128-frame TUMVI batch (1024x1024, 225 points per image).  Measured alternatives: 4 CTAs per SM
        // (224 registers, no spills) 16.3 ms, 5: 17.5 ms, 7 / 8 (128 registers, 468 B of spills): 20.5 / 19.1 ms; the four-warp
        // form (ZS_KLT63_FOUR_WARPS; one tile per warp, 128 registers, 4 CTAs per SM): 16.5 ms
        // Round 2, late: the PACKED template (gradients of pixel pairs as s16 x 2, products by IDP.2A on the bytes of jv) halves
        // the template registers: 128 registers with 36 B of spills at EIGHT CTAs per SM (the limit of shared memory) instead of
        // 168 at six: 15.89 -> 14.95 ms (6 CTAs, 158 registers, no spills: 15.48; 7: 15.56).  ZS_KLT63_UNPACKED restores the
        // register-per-pixel template, ZS_KLT63_PACKED = 6 / 7 the other occupancies.
        else if (a.win_w == 63 && !ctx->sw.klt63_four_warps && !ctx->sw.klt63_unpacked && ctx->sw.klt63_packed == 6) KLT4_LAUNCH_P(63, 63, 6, 2, 1);
        else if (a.win_w == 63 && !ctx->sw.klt63_four_warps && !ctx->sw.klt63_unpacked && ctx->sw.klt63_packed == 7) KLT4_LAUNCH_P(63, 63, 7, 2, 1);
        else if (a.win_w == 63 && !ctx->sw.klt63_four_warps && !ctx->sw.klt63_unpacked && ctx->sw.klt63_packed == 8) KLT4_LAUNCH_P(63, 63, 8, 2, 1);
        else if (a.win_w == 63 && !ctx->sw.klt63_four_warps && !ctx->sw.klt63_unpacked) KLT4_LAUNCH_P(63, 63, 8, 2, 2);
        else if (a.win_w == 63 && !ctx->sw.klt63_four_warps) KLT4_LAUNCH_T(63, 63, 6, 2);
        else if (a.win_w == 63) KLT4_LAUNCH(63, 63, 4);
        else if (a.win_w == 21) KLT4_LAUNCH(21, 21, 24);
        else KLT4_LAUNCH(15, 15, 24);
#undef KLT4_LAUNCH
#undef KLT4_LAUNCH_P
#undef KLT4_LAUNCH_T
        ZS_LAUNCH_CHECK(ctx);
        return ZS_OK;
    }
    ZS_REQUIRE(!d_out_job2, "second targets need the TMA-staged tiled kernel");
    // specialised register-template kernels for the common window sizes
    {
        const dim3 grid2(zs_div_up(cap, KLT2_WARPS), jobs);
        bool done = true;
        if (a.win_w == 31 && a.win_h == 31) k_klt_track_v2<31, 31><<<grid2, KLT2_WARPS * 32, 0, ctx->stream>>>(a);
        else if (a.win_w == 21 && a.win_h == 21) k_klt_track_v2<21, 21><<<grid2, KLT2_WARPS * 32, 0, ctx->stream>>>(a);
        else if (a.win_w == 15 && a.win_h == 15) k_klt_track_v2<15, 15><<<grid2, KLT2_WARPS * 32, 0, ctx->stream>>>(a);
        else done = false;
        if (done) { ZS_LAUNCH_CHECK(ctx); return ZS_OK; }
    }
    const int npx = a.win_w * a.win_h;
    const size_t per_warp = (size_t)((npx + 1) & ~1) * 2 + (size_t)npx * 4;
    const size_t smem = per_warp * KLT_WARPS;
    ZS_REQUIRE(smem <= 227 * 1024, "window too large for the shared-memory template");
    const bool big = (npx + 31) / 32 > 60;      // int32 lane partials hold 64 pixels (|diff*Ix| < 2^25)
    const dim3 grid(zs_div_up(cap, KLT_WARPS), jobs);
    if (big) {
        if (smem > 48 * 1024) ZS_CUDA(cudaFuncSetAttribute(k_klt_track<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_klt_track<true><<<grid, KLT_WARPS * 32, smem, ctx->stream>>>(a);
    } else {
        if (smem > 48 * 1024) ZS_CUDA(cudaFuncSetAttribute(k_klt_track<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_klt_track<false><<<grid, KLT_WARPS * 32, smem, ctx->stream>>>(a);
    }
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

extern "C" zs_status zs_klt_track(zs_context* ctx, const zs_pyramid* p, const int* d_prev_slot, const int* d_next_slot,
                                  const float* d_prev_pts, float* d_next_pts, const int* d_count, int jobs, int cap,
                                  const zs_lk_params* params, uint8_t* d_status, float* d_err)
{
    return zs_klt_launch(ctx, p, d_prev_slot, d_next_slot, d_prev_pts, d_next_pts, d_count, nullptr, jobs, cap, params,
                         d_status, d_err, 0, 0.0, nullptr, nullptr, nullptr, nullptr, 0);
}

extern "C" zs_status zs_klt_track_fb(zs_context* ctx, const zs_pyramid* p, const int* d_prev_slot, const int* d_next_slot,
                                     const float* d_prev_pts, float* d_next_pts, const int* d_count, int jobs, int cap,
                                     const zs_lk_params* params, double klt_threshold, uint8_t* d_status, float* d_err,
                                     uint8_t* d_keep)
{
    ZS_REQUIRE(d_keep, "d_keep is null");
    return zs_klt_launch(ctx, p, d_prev_slot, d_next_slot, d_prev_pts, d_next_pts, d_count, nullptr, jobs, cap, params,
                         d_status, d_err, 1, klt_threshold, d_keep, nullptr, nullptr, nullptr, 0);
}
