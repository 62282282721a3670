// zs_triangulate.cu -- stereo triangulation of matched keypoints with the reference's gates, one thread per pair.
// Reference: triangulator::triangulate_keypoints (zenslam_core/source/mapping/triangulator.cpp:39-132), its
// filter_epipolar (:152-188), utils::triangulate_points = cv::triangulatePoints (mapping/triangulation_utils.cpp:135-160),
// utils::project (utils/utils_opencv.cpp:443-480), the epipolar angle (triangulator.cpp:14-29).  SURVEY 8(f3).
//
// Everything is double precision like the reference (B200 keeps a real FP64 pipe).  The 4x4 DLT system is solved the way
// OpenCV does it: one-sided Jacobi rotations on the columns of A (JacobiSVDImpl_, eps = 10*DBL_EPSILON, <= 30 sweeps),
// singular values sorted descending, X = last row of V^T rounded to FLOAT (cv::triangulatePoints returns CV_32F for
// Point2f inputs), then X/W in double.  Embarrassingly parallel; RANSAC / map bookkeeping stay on the host.
#include <float.h>
#include <math.h>

#include "zs_common.cuh"

struct tri_args {
    double P[2][12];
    double F[9];
    double t[3];
    int use_f;
    double epipolar_threshold, reprojection_threshold, min_depth, max_depth;
    const float2* pts0; const float2* pts1; int n;
    double* xyz; uint8_t* keep; double* diag;
};

__device__ __forceinline__ void jacobi_svd4_last_vt(double At[4][4], double last[4])
{
    double W[4], Vt[4][4];
    const double eps = DBL_EPSILON * 10;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) sd = __dadd_rn(sd, __dmul_rn(At[i][k], At[i][k]));
        W[i] = sd;
#pragma unroll
        for (int k = 0; k < 4; ++k) Vt[i][k] = (i == k) ? 1.0 : 0.0;
    }
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = i + 1; j < 4; ++j) {
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < 4; ++k) p = __dadd_rn(p, __dmul_rn(At[i][k], At[j][k]));
                if (fabs(p) <= __dmul_rn(eps, sqrt(__dmul_rn(a, b)))) continue;
                p = __dmul_rn(p, 2.0);
                const double beta = __dsub_rn(a, b), gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = __dmul_rn(__dsub_rn(gamma, beta), 0.5);
                    s = sqrt(__ddiv_rn(delta, gamma));
                    c = __ddiv_rn(p, __dmul_rn(__dmul_rn(gamma, s), 2.0));
                } else {
                    c = sqrt(__ddiv_rn(__dadd_rn(gamma, beta), __dmul_rn(gamma, 2.0)));
                    s = __ddiv_rn(p, __dmul_rn(__dmul_rn(gamma, c), 2.0));
                }
                a = b = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double t0 = __dadd_rn(__dmul_rn(c, At[i][k]), __dmul_rn(s, At[j][k]));
                    const double t1 = __dadd_rn(__dmul_rn(-s, At[i][k]), __dmul_rn(c, At[j][k]));
                    At[i][k] = t0; At[j][k] = t1;
                    a = __dadd_rn(a, __dmul_rn(t0, t0)); b = __dadd_rn(b, __dmul_rn(t1, t1));
                }
                W[i] = a; W[j] = b;
                changed = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double t0 = __dadd_rn(__dmul_rn(c, Vt[i][k]), __dmul_rn(s, Vt[j][k]));
                    const double t1 = __dadd_rn(__dmul_rn(-s, Vt[i][k]), __dmul_rn(c, Vt[j][k]));
                    Vt[i][k] = t0; Vt[j][k] = t1;
                }
            }
        if (!changed) break;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) sd = __dadd_rn(sd, __dmul_rn(At[i][k], At[i][k]));
        W[i] = sqrt(sd);
    }
    // the descending selection sort of OpenCV leaves, in the last row, what this scan picks: the row that a stable
    // "swap the first maximum forward" sort ends with.  Reproduce the sort itself to keep its tie behaviour.
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        int j = i;
#pragma unroll
        for (int k = i + 1; k < 4; ++k) if (W[j] < W[k]) j = k;
        if (i != j) {
            double tmp = W[i]; W[i] = W[j]; W[j] = tmp;
#pragma unroll
            for (int k = 0; k < 4; ++k) { tmp = Vt[i][k]; Vt[i][k] = Vt[j][k]; Vt[j][k] = tmp; }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) last[k] = Vt[3][k];
}

__global__ void __launch_bounds__(128) k_triangulate(tri_args a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const float2 q0 = a.pts0[i], q1 = a.pts1[i];
    const double px[2] = { (double)q0.x, (double)q1.x }, py[2] = { (double)q0.y, (double)q1.y };
    double epi = 0; bool ok = true;
    if (a.use_f) {
        const double p1[3] = { px[1], py[1], 1.0 }, p0[3] = { px[0], py[0], 1.0 };
        double s = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double r = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) r = __dadd_rn(r, __dmul_rn(p1[k], a.F[3 * k + j]));
            s = __dadd_rn(s, __dmul_rn(r, p0[j]));
        }
        epi = s;
        ok = fabs(epi) < a.epipolar_threshold;
    }
    double At[4][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            At[k][2 * j + 0] = __dsub_rn(__dmul_rn(px[j], a.P[j][8 + k]), a.P[j][k]);
            At[k][2 * j + 1] = __dsub_rn(__dmul_rn(py[j], a.P[j][8 + k]), a.P[j][4 + k]);
        }
    double v[4];
    jacobi_svd4_last_vt(At, v);
    const float X[4] = { (float)v[0], (float)v[1], (float)v[2], (float)v[3] };
    double p3[3] = { 0, 0, 0 };
    if (fabs((double)X[3]) > 1E-9) {
        p3[0] = __ddiv_rn((double)X[0], (double)X[3]); p3[1] = __ddiv_rn((double)X[1], (double)X[3]); p3[2] = __ddiv_rn((double)X[2], (double)X[3]);
    }
    a.xyz[3 * (size_t)i] = p3[0]; a.xyz[3 * (size_t)i + 1] = p3[1]; a.xyz[3 * (size_t)i + 2] = p3[2];
    double err[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        double q[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            q[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.P[j][4 * r], p3[0]), __dmul_rn(a.P[j][4 * r + 1], p3[1])), __dmul_rn(a.P[j][4 * r + 2], p3[2])),
                             a.P[j][4 * r + 3]);
        double u = 0, w = 0;
        if (fabs(q[2]) > 1E-9) { u = __ddiv_rn(q[0], q[2]); w = __ddiv_rn(q[1], q[2]); }
        const double dx = __dsub_rn(u, px[j]), dy = __dsub_rn(w, py[j]);
        err[j] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    }
    const double v1[3] = { __dsub_rn(p3[0], a.t[0]), __dsub_rn(p3[1], a.t[1]), __dsub_rn(p3[2], a.t[2]) };
    const double n0 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(p3[0], p3[0]), __dmul_rn(p3[1], p3[1])), __dmul_rn(p3[2], p3[2])));
    const double n1 = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v1[0], v1[0]), __dmul_rn(v1[1], v1[1])), __dmul_rn(v1[2], v1[2])));
    const double nn = __dmul_rn(n0, n1);
    double ang = 0.0;
    if (!(nn < 1e-12)) {
        double cs = __ddiv_rn(__dadd_rn(__dadd_rn(__dmul_rn(p3[0], v1[0]), __dmul_rn(p3[1], v1[1])), __dmul_rn(p3[2], v1[2])), nn);
        cs = cs < -1.0 ? -1.0 : cs > 1.0 ? 1.0 : cs;
        ang = __ddiv_rn(__dmul_rn(fabs(acos(cs)), 180.0), 3.1415926535897932384626433832795);
    }
    a.keep[i] = (uint8_t)(ok && p3[2] > 0 && n0 > a.min_depth && n0 < a.max_depth && err[0] < a.reprojection_threshold &&
                          err[1] < a.reprojection_threshold && ang > 0.25 && ang < 180 - 0.25);
    if (a.diag) { a.diag[4 * (size_t)i] = epi; a.diag[4 * (size_t)i + 1] = err[0]; a.diag[4 * (size_t)i + 2] = err[1]; a.diag[4 * (size_t)i + 3] = ang; }
}

extern "C" zs_status zs_triangulate_keypoints(zs_context* ctx, const double* P0, const double* P1, const double* F, const double* t,
                                              const float* d_pts0, const float* d_pts1, int n, const zs_triangulation_params* prm,
                                              double* d_xyz, uint8_t* d_keep, double* d_diag)
{
    ZS_REQUIRE(ctx && P0 && P1 && t && prm, "null argument");
    ZS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return ZS_OK;
    ZS_REQUIRE(d_pts0 && d_pts1 && d_xyz && d_keep, "null argument");
    tri_args a;
    memcpy(a.P[0], P0, sizeof(double) * 12); memcpy(a.P[1], P1, sizeof(double) * 12); memcpy(a.t, t, sizeof(double) * 3);
    a.use_f = (F != nullptr && prm->filter_epipolar) ? 1 : 0;
    if (a.use_f) memcpy(a.F, F, sizeof(double) * 9); else memset(a.F, 0, sizeof(a.F));
    a.epipolar_threshold = prm->epipolar_threshold; a.reprojection_threshold = prm->reprojection_threshold;
    a.min_depth = prm->min_depth; a.max_depth = prm->max_depth;
    a.pts0 = (const float2*)d_pts0; a.pts1 = (const float2*)d_pts1; a.n = n; a.xyz = d_xyz; a.keep = d_keep; a.diag = d_diag;
    k_triangulate<<<zs_div_up(n, 128), 128, 0, ctx->stream>>>(a);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

extern "C" zs_status zs_triangulate_keypoints_host(zs_context* ctx, const double* P0, const double* P1, const double* F, const double* t,
                                                   const float* pts0, const float* pts1, int n, const zs_triangulation_params* prm,
                                                   double* xyz, uint8_t* keep, double* diag)
{
    ZS_REQUIRE(ctx && prm, "null argument");
    if (n <= 0) return ZS_OK;
    ZS_REQUIRE(pts0 && pts1 && xyz && keep, "null argument");
    ZS_CUDA(cudaSetDevice(ctx->device));
    const size_t o_p0 = 0, o_p1 = o_p0 + ((size_t)n * 8 + 255) / 256 * 256, o_xyz = o_p1 + ((size_t)n * 8 + 255) / 256 * 256,
                 o_diag = o_xyz + ((size_t)n * 24 + 255) / 256 * 256, o_keep = o_diag + ((size_t)n * 32 + 255) / 256 * 256, total = o_keep + n;
    zs_async_buffer buf(ctx->stream);
    ZS_CUDA(cudaMallocAsync((void**)&buf.p, total, ctx->stream));
    uint8_t* base = buf.p;
    cudaError_t e = cudaMemcpyAsync(base + o_p0, pts0, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(base + o_p1, pts1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    zs_status st = ZS_OK;
    if (e == cudaSuccess)
        st = zs_triangulate_keypoints(ctx, P0, P1, F, t, (const float*)(base + o_p0), (const float*)(base + o_p1), n, prm, (double*)(base + o_xyz),
                                      base + o_keep, diag ? (double*)(base + o_diag) : nullptr);
    if (e == cudaSuccess && st == ZS_OK) e = cudaMemcpyAsync(xyz, base + o_xyz, (size_t)n * 24, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && st == ZS_OK) e = cudaMemcpyAsync(keep, base + o_keep, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && st == ZS_OK && diag) e = cudaMemcpyAsync(diag, base + o_diag, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return zs_cuda_fail(e, "zs_triangulate_keypoints_host", __FILE__, __LINE__);
    return st;
}
