// zs_landmarks.cu -- landmark association around the hot path:
//   * the radius pre-filter of keypoint_tracker::assign_landmark_indices and utils::match_keypoints3d
//     (zenslam_core/source/tracking/keypoint_tracker.cpp:213, zenslam_core/source/matching/matching_utils.cpp:153,242 ->
//     point3d_cloud::radius_search, types/point3d_cloud.cpp:52-67),
//   * utils::match_keypoints3d itself (matching_utils.cpp:132-343): camera-frame transform, depth / frustum filter,
//     cross-checked Hamming match (the kernels of zs_match.cu), reprojection gate.
// The radius search reproduces the reference as written: nanoflann counts the landmarks with squared distance < radius^2
// (L2_Simple_Adaptor: differences squared and summed x, y, z in double) and radius_search then returns the FIRST `count`
// landmarks of the cloud in insertion order -- `this->operator()(i)`, not `matches[i].first` (point3d_cloud.cpp:61-64).
#include <math.h>

#include <vector>

#include "zs_common.cuh"

#define LM_THREADS 1024

__device__ __forceinline__ int lm_block_sum(int v, int* warp_sums)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) warp_sums[warp] = v;
    __syncthreads();
    int t = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += warp_sums[k];
    __syncthreads();
    return t;
}

// grid: sequences; xyz [S][cap][3], n [S], center [S][3] -> nt[seq] = radius > 0 ? #{squared distance < r2} : n[seq]
__global__ void __launch_bounds__(LM_THREADS) k_lm_radius_count(const double* __restrict__ xyz, const int* __restrict__ n, int cap,
                                                                const double* __restrict__ center, double radius, double r2,
                                                                int* __restrict__ nt)
{
    __shared__ int warp_sums[32];
    const int seq = blockIdx.x;
    const int m = min(n[seq], cap);
    if (!(radius > 0.0)) {                                  // keypoint_tracker.cpp:211: the search only runs for match_radius > 0
        if (threadIdx.x == 0) nt[seq] = m;
        return;
    }
    const double* p = xyz + (size_t)seq * cap * 3;
    const double cx = center[3 * seq], cy = center[3 * seq + 1], cz = center[3 * seq + 2];
    int c = 0;
    for (int i = threadIdx.x; i < m; i += LM_THREADS) {
        const double dx = __dsub_rn(cx, p[3 * i]), dy = __dsub_rn(cy, p[3 * i + 1]), dz = __dsub_rn(cz, p[3 * i + 2]);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        c += d2 < r2 ? 1 : 0;
    }
    c = lm_block_sum(c, warp_sums);
    if (threadIdx.x == 0) nt[seq] = c;
}

zs_status zs_lm_radius_count(zs_context* ctx, const double* d_xyz, const int* d_n, int cap, const double* d_center, double radius,
                             int sequences, int* d_nt)
{
    k_lm_radius_count<<<sequences, LM_THREADS, 0, ctx->stream>>>(d_xyz, d_n, cap, d_center, radius, radius * radius, d_nt);
    ZS_LAUNCH_CHECK(ctx);
    return ZS_OK;
}

struct lm_cam_args {
    const double* xyz; const uint8_t* desc; const int* count;          // landmarks in insertion order, *count candidates
    double Ri[9], ti[3], P[12];                                        // inverse pose (camera <- world), projection
    int frustum, width, height; double margin;
    double* cam; uint8_t* desc_out; int* row_out; int* n_out;          // compacted: camera-frame xyz, descriptor, source row
};

__device__ __forceinline__ bool lm_project(const double* P, double x, double y, double z, double& u, double& v, double& w)
{
    const double a = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[0], x), __dmul_rn(P[1], y)), __dmul_rn(P[2], z)), P[3]);
    const double b = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[4], x), __dmul_rn(P[5], y)), __dmul_rn(P[6], z)), P[7]);
    w = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[8], x), __dmul_rn(P[9], y)), __dmul_rn(P[10], z)), P[11]);
    const bool ok = fabs(w) > 1e-9;
    u = ok ? __ddiv_rn(a, w) : 0.0; v = ok ? __ddiv_rn(b, w) : 0.0;      // utils::project: (0, 0) when |w| <= 1e-9
    return ok;
}

// one block: candidates -> camera frame, z > 0 (+ frustum), ordered compaction (matching_utils.cpp:152-157, 240-275)
__global__ void __launch_bounds__(LM_THREADS) k_lm_to_camera(lm_cam_args a)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = *a.count;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += LM_THREADS) {
        const int i = base + threadIdx.x;
        bool f = false;
        double x = 0, y = 0, z = 0;
        if (i < m) {
            const double px = a.xyz[3 * i], py = a.xyz[3 * i + 1], pz = a.xyz[3 * i + 2];
            // cv::Affine3d * cv::Point3d: m0 x + m1 y + m2 z + m3, left to right
            x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.Ri[0], px), __dmul_rn(a.Ri[1], py)), __dmul_rn(a.Ri[2], pz)), a.ti[0]);
            y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.Ri[3], px), __dmul_rn(a.Ri[4], py)), __dmul_rn(a.Ri[5], pz)), a.ti[1]);
            z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a.Ri[6], px), __dmul_rn(a.Ri[7], py)), __dmul_rn(a.Ri[8], pz)), a.ti[2]);
            f = z > 0.0;
            if (f && a.frustum) {                            // is_in_frustum (matching_utils.cpp:105-130)
                double u, v, w;
                lm_project(a.P, x, y, z, u, v, w);
                f = !(fabs(w) < 1e-9) && u >= -a.margin && u < (double)a.width + a.margin && v >= -a.margin && v < (double)a.height + a.margin;
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_sums[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < LM_THREADS / 32; ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
        if (f) {
            const int o = carry + woff + __popc(bal & ((1u << lane) - 1));
            a.cam[3 * o] = x; a.cam[3 * o + 1] = y; a.cam[3 * o + 2] = z;
            a.row_out[o] = i;
            const uint4* s = (const uint4*)(a.desc + (size_t)i * 32);
            uint4* d = (uint4*)(a.desc_out + (size_t)o * 32);
            d[0] = s[0]; d[1] = s[1];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.n_out = carry;
}

struct lm_gate_args {
    const double* cam; const int* row; const int* n_sel; const int* match; const float* kp_xy;
    const int* lm_index; const int* kp_index;
    double P[12]; double threshold;
    int* out_lm; int* out_kp; float* out_err; int* n_out;
};

// one block: reprojection gate on the cross-checked matches, ordered (matching_utils.cpp:188-213, 306-331)
__global__ void __launch_bounds__(LM_THREADS) k_lm_reproject_gate(lm_gate_args a)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = *a.n_sel;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += LM_THREADS) {
        const int j = base + threadIdx.x;
        bool f = false;
        double err = 0;
        int t = -1;
        if (j < m && (t = a.match[j]) >= 0) {
            double u, v, w;
            lm_project(a.P, a.cam[3 * j], a.cam[3 * j + 1], a.cam[3 * j + 2], u, v, w);
            const double ex = __dsub_rn(u, (double)a.kp_xy[2 * t]), ey = __dsub_rn(v, (double)a.kp_xy[2 * t + 1]);
            err = sqrt(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));          // cv::norm(Point2d)
            f = err < a.threshold;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_sums[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
        for (int k = 0; k < LM_THREADS / 32; ++k) { const int s = warp_sums[k]; if (k < warp) woff += s; total += s; }
        if (f) {
            const int o = carry + woff + __popc(bal & ((1u << lane) - 1));
            a.out_lm[o] = a.lm_index[a.row[j]]; a.out_kp[o] = a.kp_index[t]; a.out_err[o] = (float)err;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.n_out = carry;
}

static inline size_t lm_al(size_t v) { return (v + 255) / 256 * 256; }

extern "C" zs_status zs_match_keypoints3d_host(zs_context* ctx, const int* landmark_index, const double* landmark_xyz,
                                               const uint8_t* landmark_desc, int n_landmarks, const int* keypoint_index,
                                               const float* keypoint_xy, const uint8_t* keypoint_desc, int n_keypoints,
                                               const double* R, const double* t, const double* P, double radius, double threshold,
                                               int image_width, int image_height, double frustum_margin, int* out_landmark_index,
                                               int* out_keypoint_index, float* out_error, int* n_out)
{
    ZS_REQUIRE(ctx && n_out && R && t && P, "null argument");
    ZS_REQUIRE(n_landmarks >= 0 && n_keypoints >= 0, "negative count");
    *n_out = 0;
    if (n_landmarks == 0 || n_keypoints == 0) return ZS_OK;          // matching_utils.cpp:140-141, 147-148
    ZS_REQUIRE(landmark_index && landmark_xyz && landmark_desc && keypoint_index && keypoint_xy && keypoint_desc && out_landmark_index &&
               out_keypoint_index && out_error, "null array");
    ZS_CUDA(cudaSetDevice(ctx->device));
    const size_t M = n_landmarks, N = n_keypoints;
    const size_t o_xyz = 0, o_desc = o_xyz + lm_al(M * 24), o_lmi = o_desc + lm_al(M * 32), o_kd = o_lmi + lm_al(M * 4),
                 o_kxy = o_kd + lm_al(N * 32), o_kpi = o_kxy + lm_al(N * 8), o_cam = o_kpi + lm_al(N * 4), o_dc = o_cam + lm_al(M * 24),
                 o_row = o_dc + lm_al(M * 32), o_match = o_row + lm_al(M * 4), o_dist = o_match + lm_al(M * 4),
                 o_olm = o_dist + lm_al(M * 4), o_okp = o_olm + lm_al(M * 4), o_oerr = o_okp + lm_al(M * 4), o_cnt = o_oerr + lm_al(M * 4),
                 total = o_cnt + 256;
    zs_async_buffer buf(ctx->stream);
    ZS_CUDA(cudaMallocAsync((void**)&buf.p, total, ctx->stream));
    uint8_t* b = buf.p;
    int* cnt = (int*)(b + o_cnt);        // [0] landmarks, [1] candidates after the radius search, [2] after the filters, [3] keypoints, [4] matches out
    const int h_cnt[5] = { n_landmarks, 0, 0, n_keypoints, 0 };
    double center[3] = { t[0], t[1], t[2] };
    ZS_CUDA(cudaMemcpyAsync(b + o_xyz, landmark_xyz, M * 24, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(b + o_desc, landmark_desc, M * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(b + o_lmi, landmark_index, M * 4, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(b + o_kd, keypoint_desc, N * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(b + o_kxy, keypoint_xy, N * 8, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(b + o_kpi, keypoint_index, N * 4, cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(cnt, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, ctx->stream));
    ZS_CUDA(cudaMemcpyAsync(cnt + 8, center, sizeof(center), cudaMemcpyHostToDevice, ctx->stream));      // doubles at byte 32
    // radius_search around the camera: radius <= 0 finds nothing (squared distance < radius^2 is never true for radius 0;
    // a negative radius squares to a positive bound, like in the reference)
    k_lm_radius_count<<<1, LM_THREADS, 0, ctx->stream>>>((const double*)(b + o_xyz), cnt, n_landmarks, (const double*)(cnt + 8), 1.0,
                                                         radius * radius, cnt + 1);
    ZS_LAUNCH_CHECK(ctx);
    lm_cam_args ca;
    ca.xyz = (const double*)(b + o_xyz); ca.desc = b + o_desc; ca.count = cnt + 1;
    // pose_of_camera0_in_world.inv(): (R^T, -R^T t).  The reference inverts the 4x4 matrix numerically (cv::Affine3d::inv ->
    // Matx::inv, DECOMP_SVD), so its camera-frame coordinates agree with these to rounding, not bit for bit.
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) ca.Ri[3 * r + c] = R[3 * c + r];
        ca.ti[r] = -(R[0 * 3 + r] * t[0] + R[1 * 3 + r] * t[1] + R[2 * 3 + r] * t[2]);
    }
    for (int i = 0; i < 12; ++i) ca.P[i] = P[i];
    ca.frustum = image_width > 0 && image_height > 0; ca.width = image_width; ca.height = image_height; ca.margin = frustum_margin;
    ca.cam = (double*)(b + o_cam); ca.desc_out = b + o_dc; ca.row_out = (int*)(b + o_row); ca.n_out = cnt + 2;
    k_lm_to_camera<<<1, LM_THREADS, 0, ctx->stream>>>(ca);
    ZS_LAUNCH_CHECK(ctx);
    // cv::BFMatcher(NORM_HAMMING, true).match(descriptors3d, descriptors2d): landmarks are the query side
    zs_status st = zs_match_hamming_cross(ctx, b + o_dc, cnt + 2, M * 32, b + o_kd, cnt + 3, N * 32, 1, n_landmarks, n_keypoints,
                                          (int*)(b + o_match), (float*)(b + o_dist));
    if (st != ZS_OK) return st;
    lm_gate_args ga;
    ga.cam = ca.cam; ga.row = ca.row_out; ga.n_sel = cnt + 2; ga.match = (const int*)(b + o_match); ga.kp_xy = (const float*)(b + o_kxy);
    ga.lm_index = (const int*)(b + o_lmi); ga.kp_index = (const int*)(b + o_kpi);
    for (int i = 0; i < 12; ++i) ga.P[i] = P[i];
    ga.threshold = threshold;
    ga.out_lm = (int*)(b + o_olm); ga.out_kp = (int*)(b + o_okp); ga.out_err = (float*)(b + o_oerr); ga.n_out = cnt + 4;
    k_lm_reproject_gate<<<1, LM_THREADS, 0, ctx->stream>>>(ga);
    ZS_LAUNCH_CHECK(ctx);
    int n = 0;
    ZS_CUDA(cudaMemcpyAsync(&n, cnt + 4, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n > 0) {
        ZS_CUDA(cudaMemcpyAsync(out_landmark_index, ga.out_lm, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(out_keypoint_index, ga.out_kp, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaMemcpyAsync(out_error, ga.out_err, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
        ZS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *n_out = n;
    return ZS_OK;
}
