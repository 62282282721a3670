/*
 * zenslam_cuda.h -- C ABI of libzenslam_cuda.so, the B200 (sm_100a) backend for ZenSLAM's
 * per-frame stereo front-end: optical-flow pyramids, FAST/grid detection, ORB description,
 * brute-force matching and pyramidal KLT.
 *
 * The reference (vinodkhare/zenslam) has no C ABI: its seams are C++ virtual interfaces over
 * OpenCV types.  Every entry point below names the reference interface it stands behind
 * (file:line relative to the reference root); the C++ adapter in zenslam_cuda/ converts
 * cv::Mat / std::vector to these calls, exactly as zenslam_metal/ does for Metal
 * (zenslam_metal/source/pyr_lk_factory.cpp:7-49).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns zs_status (0 = ok, < 0 = error),
 *     never throws, never allocates caller-visible memory;
 *   - `_host` functions take HOST pointers, perform the H2D/D2H copies themselves and return
 *     when the results are in the caller's buffers (drop-in semantics of the reference calls);
 *   - all other functions take DEVICE pointers, enqueue work on the context's stream and
 *     return without synchronising;
 *   - there is no CPU fallback: without a usable sm_100 device zs_context_create fails.
 */
#ifndef ZENSLAM_CUDA_H
#define ZENSLAM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZS_API __attribute__((visibility("default")))

typedef int zs_status;
enum {
    ZS_OK = 0,
    ZS_ERR_NO_DEVICE = -1,     /* no CUDA device / not an sm_100 part */
    ZS_ERR_INVALID = -2,       /* bad argument */
    ZS_ERR_CUDA = -3,          /* CUDA runtime error; see zs_last_error_string */
    ZS_ERR_CAPACITY = -4,      /* caller buffer too small */
    ZS_ERR_UNSUPPORTED = -5    /* valid in the reference but outside this backend (e.g. non-integer L2 input) */
};

/* cv::OPTFLOW_* values the reference passes (keypoint_tracker.cpp:153,390) */
#define ZS_LK_USE_INITIAL_FLOW 4
#define ZS_LK_GET_MIN_EIGENVALS 8

typedef struct zs_context zs_context;    /* device + stream + scratch */
typedef struct zs_pyramid zs_pyramid;    /* S image slots, each a padded optical-flow pyramid in HBM */
typedef struct zs_frontend zs_frontend;  /* batched stereo front-end (native per-frame pipeline) */

/* ---- availability / context ---------------------------------------------------------------
 * cf. zenslam::metal::is_available() (zenslam_metal/include/zenslam_metal/pyr_lk.h:9) and the
 * factory that returns an empty pointer when the backend is absent (pyr_lk_factory.cpp:41-49). */
ZS_API int zs_is_available(void);
ZS_API const char* zs_version(void);
ZS_API const char* zs_status_string(zs_status s);
ZS_API const char* zs_last_error_string(void);

/* stream: a cudaStream_t to enqueue on, or NULL to let the context create its own. */
ZS_API zs_status zs_context_create(int device, void* stream, zs_context** out);
ZS_API void zs_context_destroy(zs_context* ctx);
ZS_API zs_status zs_context_synchronize(zs_context* ctx);
/* errors that kernels of stream-asynchronous entries found in their DATA (zs_match_l2_*: descriptors that are not integers
 * in 0..255): waits for the stream, returns ZS_ERR_UNSUPPORTED with a message if one was raised since the last call, and
 * clears it */
ZS_API zs_status zs_context_async_error(zs_context* ctx);
ZS_API void* zs_context_stream(zs_context* ctx);
/* number of kernels this library has launched on the context since creation */
ZS_API uint64_t zs_context_launch_count(const zs_context* ctx);
/* The ZS_* A/B switches are environment variables read once, when the context is created, never on a
 * launch path; this re-reads them (measurement / test hook; every variant computes the same results). */
ZS_API zs_status zs_context_reload_switches(zs_context* ctx);

/* ---- pre-processing: processor::process (zenslam_core/source/processor.cpp:25-55), SURVEY 8(f1) -----------
 * utils::convert_color(BGR2GRAY) -> optional CLAHE(4.0, 8x8) (processor.h:38) -> utils::rectify = cv::remap(INTER_LINEAR)
 * with the CV_32FC1 maps of cv::initUndistortRectifyMap (utils.cpp:119-124, calibration.cpp:60-70).  Device pointers;
 * `count` images `stride` bytes apart (stride 0 = the same image / the same maps for every item); bit-exact vs cv2. */
ZS_API zs_status zs_cvt_bgr2gray(zs_context* ctx, const uint8_t* d_bgr, size_t pitch, size_t stride, int width, int height,
                                 int count, uint8_t* d_gray, size_t gray_pitch, size_t gray_stride);
ZS_API zs_status zs_clahe(zs_context* ctx, const uint8_t* d_src, size_t pitch, size_t stride, int width, int height,
                          int count, double clip_limit, int tiles_x, int tiles_y, uint8_t* d_dst, size_t dst_pitch,
                          size_t dst_stride);
/* map_pitch / map_stride are in floats; taps outside the source read 0 (BORDER_CONSTANT) */
ZS_API zs_status zs_remap_linear(zs_context* ctx, const uint8_t* d_src, size_t pitch, size_t stride, int src_width,
                                 int src_height, int count, const float* d_map_x, const float* d_map_y, size_t map_pitch,
                                 size_t map_stride, int dst_width, int dst_height, uint8_t* d_dst, size_t dst_pitch,
                                 size_t dst_stride);
/* the image path of processor::process for one HOST image: channels 3 (BGR) or 1; map_x / map_y may be NULL (no
 * rectification); `undistorted` receives width*height bytes (frame::processed::undistorted) */
ZS_API zs_status zs_process_image_host(zs_context* ctx, const uint8_t* image, int channels, int width, int height,
                                       size_t pitch, int clahe_enabled, double clahe_clip_limit, const float* map_x,
                                       const float* map_y, uint8_t* undistorted);

/* ---- pyramids: cv::buildOpticalFlowPyramid(img, pyr, win, maxLevel, withDerivatives=true) ----
 * utils::pyramid (zenslam_core/source/utils/utils_opencv.cpp:525-530; called processor.cpp:37,53).
 * One zs_pyramid holds `slots` images of identical size.  Level l of slot s is an image plane
 * padded REFLECT_101 by the window plus a zero-padded (dx,dy) int16 Scharr plane. */
ZS_API zs_status zs_pyramid_create(zs_context* ctx, int width, int height, int slots,
                                   int win_w, int win_h, int max_level, zs_pyramid** out);
ZS_API void zs_pyramid_destroy(zs_pyramid* p);
ZS_API int zs_pyramid_levels(const zs_pyramid* p);
ZS_API zs_status zs_pyramid_level_size(const zs_pyramid* p, int level, int* width, int* height);
/* copy `count` images into level 0 of slots first..first+count-1 (slot index taken modulo slots).
 * src_is_host selects cudaMemcpy kind; pitch = bytes per row, stride = bytes per image. */
ZS_API zs_status zs_pyramid_upload(zs_context* ctx, zs_pyramid* p, const uint8_t* src, size_t pitch,
                                   size_t stride, int first, int count, int src_is_host);
/* build levels 1.. and all derivative planes for slots first..first+count-1 (modulo slots) */
ZS_API zs_status zs_pyramid_build(zs_context* ctx, zs_pyramid* p, int first, int count);
/* device address of the level-0 interior of a slot (pitch / slot stride in bytes): lets the pre-processing kernels
 * above produce a frame directly where the pyramid kernels read it */
ZS_API zs_status zs_pyramid_level0(const zs_pyramid* p, int slot, uint8_t** d_ptr, size_t* pitch, size_t* slot_stride);
/* test access: copy one un-padded level image (u8, w*h) / derivative plane (int16, w*h*2) to host */
ZS_API zs_status zs_pyramid_download_image(zs_context* ctx, const zs_pyramid* p, int slot, int level, uint8_t* dst);
ZS_API zs_status zs_pyramid_download_deriv(zs_context* ctx, const zs_pyramid* p, int slot, int level, int16_t* dst);

/* ---- detection: keypoint_detector_grid::detect_keypoints -----------------------------------
 * (zenslam_core/source/detection/keypoint_detector_grid.cpp:39-150; interface keypoint_detector.h:8-14)
 * Per unoccupied cell: cv::FAST(threshold, NMS, TYPE_9_16) on the cell ROI, first maximum-response
 * corner, cell row-major order.  Works on level 0 of slots first..first+count-1.
 * d_occupied: [count][grid_h*grid_w] bytes or NULL.  Outputs are [count][cap] device arrays
 * (cap >= grid_w*grid_h; d_xy is [count][cap][2], x,y interleaved like cv::KeyPoint::pt) plus d_count[count]. */
ZS_API zs_status zs_fast_grid_detect(zs_context* ctx, const zs_pyramid* p, int first, int count,
                                     int cell_w, int cell_h, int threshold, const uint8_t* d_occupied,
                                     float* d_xy, float* d_response, int* d_count, int cap);

/* full-frame cv::FAST(img, threshold, true) for keypoint_detector_simple
 * (zenslam_core/source/detection/keypoint_detector_simple.cpp:38-63): raster order, optional mask
 * d_mask [count][height][width] (0 = rejected), cap entries per image; d_count receives the true
 * number found (may exceed cap, in which case only the first cap are stored). */
ZS_API zs_status zs_fast_detect(zs_context* ctx, const zs_pyramid* p, int first, int count, int threshold,
                                const uint8_t* d_mask, float* d_xy, float* d_response,
                                int* d_count, int cap);

/* ---- sub-pixel refinement: cv::cornerSubPix(image, pts, win, Size(-1,-1), {EPS+COUNT, max_iters, epsilon}) ----
 * keypoint_detector_parallel::detect_keypoints refines the grid corners with win (5,5), 30 iterations, eps 0.01
 * (zenslam_core/source/detection/keypoint_detector_parallel.cpp:160-170).  d_xy is [count][cap][2], refined in place
 * on level 0 of slots first..first+count-1; win_w / win_h are HALF sizes (1..7), like cv::Size(5,5).  Bit-identical to
 * OpenCV built without IPP. */
ZS_API zs_status zs_corner_subpix(zs_context* ctx, const zs_pyramid* p, int first, int count, float* d_xy,
                                  const int* d_count, int cap, int win_w, int win_h, int max_iters, double epsilon);

/* ---- description: cv::ORB::create()->compute(image, keypoints, descriptors) ----------------
 * (keypoint_detector_grid.cpp:138).  Drops keypoints outside the 31-px border keeping order,
 * blurs (7x7 sigma 2, float32), 256 rBRIEF tests rotated by the keypoint angle (degrees; NULL = -1,
 * what FAST keypoints carry).  In/out arrays are [count][cap]; d_src_index (optional) receives the
 * input row each surviving keypoint came from. */
ZS_API zs_status zs_orb_compute(zs_context* ctx, const zs_pyramid* p, int first, int count,
                                const float* d_xy_in, const float* d_resp_in,
                                const float* d_angle_in, const int* d_count_in, int cap,
                                float* d_xy, float* d_resp, int* d_src_index, int* d_count,
                                uint8_t* d_desc /* [count][cap][32] */);
/* test access: the blurred level-0 image ORB samples from (u8, w*h) */
ZS_API zs_status zs_orb_download_blur(zs_context* ctx, const zs_pyramid* p, int slot, uint8_t* dst);

/* ---- multi-scale ORB detector: `feature: ORB` ----------------------------------------------
 * cv::ORB::create(nfeatures, scale_factor, nlevels, edge_threshold, 0, 2, HARRIS_SCORE, patch_size, fast_threshold)
 *   ->detect(image, keypoints, mask), then (optional) cv::ORB::create()->compute(image, keypoints, descriptors)
 * (zenslam_core/source/detection/keypoint_detector_simple.cpp:17,27,49,54; the reference passes 500, 1.2f, 8, 31, 31 and
 * detection.fast_threshold).  Images [count][height][pitch] u8 on the device, optional masks of the same shape (0 =
 * rejected).  Outputs are [count][cap] device arrays, cap = zs_orb_detector_capacity(); d_xy is (x, y) interleaved;
 * d_count receives the true number per image (ZS_ERR_CAPACITY is never raised: at most cap are stored).
 * The keypoint set, sizes, angles, Harris responses, octaves and descriptors equal OpenCV's; the order is canonical
 * (octave ascending, then y, then x) because OpenCV's own order comes out of std::nth_element.  d_desc may be NULL
 * (detect only); describing needs edge_threshold >= 31 and patch_size 31 (cv::ORB::create() defaults on the compute side). */
typedef struct zs_orb_detector zs_orb_detector;
ZS_API zs_status zs_orb_detector_create(zs_context* ctx, int width, int height, int max_images, int nfeatures,
                                        float scale_factor, int nlevels, int edge_threshold, int patch_size,
                                        int fast_threshold, zs_orb_detector** out);
ZS_API void zs_orb_detector_destroy(zs_orb_detector* det);
ZS_API int zs_orb_detector_capacity(const zs_orb_detector* det);
ZS_API zs_status zs_orb_detector_level(const zs_orb_detector* det, int level, int* width, int* height, float* scale,
                                       int* nfeatures);
ZS_API zs_status zs_orb_detect_and_compute(zs_context* ctx, zs_orb_detector* det, const uint8_t* d_img, size_t pitch,
                                           size_t stride, const uint8_t* d_mask, size_t mask_pitch, size_t mask_stride,
                                           int count, float* d_xy, float* d_size, float* d_angle, float* d_response,
                                           int* d_octave, int* d_count, uint8_t* d_desc /* [count][cap][32] or NULL */);
/* test access: which = 0 pyramid image, 1 mask, 2 blurred image of `level` of image `image` (u8, level w*h) */
ZS_API zs_status zs_orb_detector_download_level(zs_context* ctx, const zs_orb_detector* det, int image, int level,
                                                int which, uint8_t* dst);
/* host mirror of keypoint_detector_simple::detect_keypoints with `feature: ORB` (mask may be NULL; outputs sized cap;
 * *n_out = keypoints found, ZS_ERR_CAPACITY if that exceeds cap) */
ZS_API zs_status zs_detect_keypoints_orb_host(zs_context* ctx, const uint8_t* img, int width, int height, size_t pitch,
                                              const uint8_t* mask, size_t mask_pitch, int nfeatures, float scale_factor,
                                              int nlevels, int edge_threshold, int patch_size, int fast_threshold,
                                              float* x, float* y, float* size, float* angle, float* response,
                                              int* octave, uint8_t* desc /* or NULL */, int cap, int* n_out);

/* ---- matching: cv::BFMatcher as used by zenslam::matcher --------------------------------------
 * (zenslam_core/source/matching/matcher.cpp:60-80; utils::create_matcher matching_utils.cpp:63-95)
 * `pairs` independent problems; problem k matches queries d_q + k*q_stride (nq[k] rows) against
 * train d_t + k*t_stride (nt[k] rows); strides are in elements of the pointer type (bytes for Hamming,
 * floats for L2).  Descriptors are 32-byte rows (Hamming) or `dim` floats (L2).
 * knn2: per query the two nearest (ties -> smaller train index): d_idx [pairs][cap_q][2] (-1 = none),
 * d_dist [pairs][cap_q][2] (Hamming: popcount as float; L2: sqrtf of the exact integer distance),
 * d_pass [pairs][cap_q]: Lowe ratio gate of matcher.cpp:70 ((double)d0 < ratio*(double)d1, two neighbours).
 * cross: BFMatcher(crossCheck=true).match: d_idx [pairs][cap_q] = train index or -1, d_dist likewise.
 * Hamming calls of at least 12 M distances (pairs * cap_q * cap_t) run on the tcgen05 tensor cores too: the descriptors are
 * expanded to one 0 / 1 byte per bit, for which the squared L2 distance IS the Hamming distance, and go through the kind::i8
 * kernel of the L2 matcher (same exact integers, same tie rule); smaller calls use the CUDA-core kernel. */
ZS_API zs_status zs_match_hamming_knn2(zs_context* ctx, const uint8_t* d_q, const int* d_nq, size_t q_stride,
                                       const uint8_t* d_t, const int* d_nt, size_t t_stride, int pairs,
                                       int cap_q, int cap_t, double ratio,
                                       int* d_idx, float* d_dist, uint8_t* d_pass);
ZS_API zs_status zs_match_hamming_cross(zs_context* ctx, const uint8_t* d_q, const int* d_nq, size_t q_stride,
                                        const uint8_t* d_t, const int* d_nt, size_t t_stride, int pairs,
                                        int cap_q, int cap_t, int* d_idx, float* d_dist);
/* L2 on integer-valued float descriptors (cv::SIFT, values 0..255): exact.  128-d rows run the -2 A.B^T
 * contraction on the tcgen05 tensor cores (kind::i8 on the u8-narrowed rows); other dims (multiples of 4, <= 128) a dp4a
 * kernel.  The calls are stream-asynchronous and never wait for the device: rows holding anything but integers in 0..255
 * make EVERY match of that call come back as -1 (d_pass 0) and raise a flag that zs_context_async_error() -- and the host
 * entries zs_match_host / zs_knn_match_host -- report as ZS_ERR_UNSUPPORTED. */
ZS_API zs_status zs_match_l2_knn2(zs_context* ctx, const float* d_q, const int* d_nq, size_t q_stride,
                                  const float* d_t, const int* d_nt, size_t t_stride, int pairs,
                                  int cap_q, int cap_t, int dim, double ratio,
                                  int* d_idx, float* d_dist, uint8_t* d_pass);
ZS_API zs_status zs_match_l2_cross(zs_context* ctx, const float* d_q, const int* d_nq, size_t q_stride,
                                   const float* d_t, const int* d_nt, size_t t_stride, int pairs,
                                   int cap_q, int cap_t, int dim, int* d_idx, float* d_dist);
/* The same two matchers on u8 rows (cv::SIFT::create(..., descriptorType = CV_8U) output, or float rows narrowed where they
 * are produced): dense [pairs][cap][dim] bytes, 16-byte aligned; no conversion pass and a quarter of the bytes to move.
 * Same results as the float entries on the same values (cv::BFMatcher(NORM_L2) on CV_8U rows). */
ZS_API zs_status zs_match_l2_knn2_u8(zs_context* ctx, const uint8_t* d_q, const int* d_nq, const uint8_t* d_t,
                                     const int* d_nt, int pairs, int cap_q, int cap_t, int dim, double ratio,
                                     int* d_idx, float* d_dist, uint8_t* d_pass);
ZS_API zs_status zs_match_l2_cross_u8(zs_context* ctx, const uint8_t* d_q, const int* d_nq, const uint8_t* d_t,
                                      const int* d_nt, int pairs, int cap_q, int cap_t, int dim,
                                      int* d_idx, float* d_dist);

/* ---- KLT: pyr_lk::calc_optical_flow_pyr_lk == cv::calcOpticalFlowPyrLK ----------------------
 * (zenslam_core/include/zenslam/tracking/pyr_lk.h:15-26; zenslam_core/source/tracking/pyr_lk.cpp:25)
 * `jobs` independent calls: job j tracks d_count[j] points from slot d_prev_slot[j] to slot
 * d_next_slot[j] of the same zs_pyramid.  Point arrays are [jobs][cap] float2 (x,y interleaved);
 * d_next_pts is read when flags has ZS_LK_USE_INITIAL_FLOW.  flags must include
 * ZS_LK_GET_MIN_EIGENVALS (the only mode the reference uses), so err = min-eigenvalue. */
typedef struct {
    int win_w, win_h;        /* tracking.klt_window_size */
    int max_level;           /* tracking.klt_max_level */
    int max_iters;           /* TermCriteria.maxCount (99) */
    double epsilon;          /* TermCriteria.epsilon (0.001) */
    int flags;
    double min_eig_threshold;/* 1e-4 */
} zs_lk_params;

ZS_API zs_status zs_klt_track(zs_context* ctx, const zs_pyramid* p, const int* d_prev_slot, const int* d_next_slot,
                              const float* d_prev_pts, float* d_next_pts, const int* d_count, int jobs, int cap,
                              const zs_lk_params* params, uint8_t* d_status, float* d_err);

/* forward + backward LK and the forward-backward gate of keypoint_tracker::track_keypoints
 * (zenslam_core/source/tracking/keypoint_tracker.cpp:129-197, 343-434): d_keep[j][i] =
 * status && status_back && ||p0_back - p0|| < klt_threshold.  d_next_pts holds the forward result. */
ZS_API zs_status zs_klt_track_fb(zs_context* ctx, const zs_pyramid* p, const int* d_prev_slot, const int* d_next_slot,
                                 const float* d_prev_pts, float* d_next_pts, const int* d_count, int jobs, int cap,
                                 const zs_lk_params* params, double klt_threshold,
                                 uint8_t* d_status, float* d_err, uint8_t* d_keep);

/* ---- host-pointer mirrors of the reference calls (single image, synchronous) ----------------- */

/* pyr_lk::calc_optical_flow_pyr_lk with level-0 images (pyramids are rebuilt on the device:
 * identical to what the reference's processor builds, utils_opencv.cpp:525-530). */
ZS_API zs_status zs_calc_optical_flow_pyr_lk_host(zs_context* ctx,
                                                  const uint8_t* prev_img, const uint8_t* next_img,
                                                  int width, int height, size_t pitch,
                                                  const float* prev_pts, float* next_pts, int n,
                                                  uint8_t* status, float* err, const zs_lk_params* params);
/* The two LK host entries keep the device pyramids of the last few frames (keyed by a 64-bit hash of the frame's bytes,
 * least-recently-used replacement), because the reference's eight LK calls per stereo frame see only four distinct images
 * and two of those were already seen by the previous frame's calls.  A hit skips the upload and the pyramid build; results
 * are unaffected.  ZS_LK_NO_CACHE=1 in the environment disables it.  This returns the hit / miss counts (test access). */
ZS_API zs_status zs_lk_cache_stats(const zs_context* ctx, uint64_t* hits, uint64_t* misses);
/* keypoint_tracker::track_keypoints as ONE call (keypoint_tracker.cpp:129-197, :343-434): forward LK from points_0
 * (initial flow predicted_1, or NULL for none), backward LK from the forward results, keep[i] = both statuses set and
 * ||p0_back - p0|| < klt_threshold.  Same results as two zs_calc_optical_flow_pyr_lk_host calls + the reference's gate,
 * with one upload / pyramid build per frame instead of two and both passes in one kernel.  points_1 receives the
 * forward results for every point; status / err (forward pass) may be NULL. */
ZS_API zs_status zs_track_keypoints_host(zs_context* ctx, const uint8_t* img_0, const uint8_t* img_1,
                                         int width, int height, size_t pitch, const float* points_0,
                                         const float* predicted_1, int n, const zs_lk_params* params,
                                         double klt_threshold, float* points_1, uint8_t* status, float* err,
                                         uint8_t* keep);
/* keypoint_detector_grid::detect_keypoints: occupied [grid_h*grid_w] or NULL; outputs sized
 * grid_w*grid_h (x, y, response) and *32 (desc); *n_out = keypoints that survive ORB's border filter.
 * Limit: the reference sends a free cell where FAST finds nothing through _describer->detect (the 8-level ORB detector,
 * keypoint_detector_grid.cpp:92-95).  That can never yield a keypoint for cells up to 62 px (31-px edge threshold) -- every
 * configuration of BASELINE.json -- and is not implemented: with cells >= 63 px this call returns ZS_ERR_UNSUPPORTED when
 * such a cell occurs instead of diverging silently.  PARALLEL_GRID (tumvi.yaml) has no such step in the reference. */
ZS_API zs_status zs_detect_keypoints_grid_host(zs_context* ctx, const uint8_t* img, int width, int height,
                                               size_t pitch, int cell_w, int cell_h, int threshold,
                                               const uint8_t* occupied, float* x, float* y, float* response,
                                               uint8_t* desc, int* n_out);
/* keypoint_detector_parallel::detect_keypoints (keypoint_detector_parallel.cpp:40-193): the same cells and first
 * strongest corner per cell as the grid detector, then cv::cornerSubPix (5x5 half window, 30 iterations, eps 0.01),
 * then ORB::compute at cvRound(pt).  Same arguments and outputs as zs_detect_keypoints_grid_host; x / y come back
 * sub-pixel. */
ZS_API zs_status zs_detect_keypoints_parallel_host(zs_context* ctx, const uint8_t* img, int width, int height,
                                                   size_t pitch, int cell_w, int cell_h, int threshold,
                                                   const uint8_t* occupied, float* x, float* y, float* response,
                                                   uint8_t* desc, int* n_out);
/* keypoint_detector_simple::detect_keypoints with `feature: FAST` (keypoint_detector_simple.cpp:38-63): full-frame
 * cv::FAST(threshold, NMS) + mask (host, [height][mask_pitch], 0 = rejected, may be NULL), then ORB::compute.  Outputs
 * sized cap; ZS_ERR_CAPACITY (with *n_out = corners found) when the frame holds more corners than cap. */
ZS_API zs_status zs_detect_keypoints_simple_host(zs_context* ctx, const uint8_t* img, int width, int height,
                                                 size_t pitch, const uint8_t* mask, size_t mask_pitch, int threshold,
                                                 float* x, float* y, float* response, uint8_t* desc, int cap,
                                                 int* n_out);
/* matcher::match_keypoints descriptor stage: mode 0 = KNN (ratio), 1 = BRUTE (cross-check);
 * norm 0 = Hamming (32-byte rows), 1 = L2 (dim floats), 2 = L2 on u8 rows (dim bytes).  Outputs sized nq; *n_out matches. */
ZS_API zs_status zs_match_host(zs_context* ctx, const void* q, int nq, const void* t, int nt, int dim,
                               int norm, int mode, double ratio, int* query_idx, int* train_idx, float* distance,
                               int* n_out);

/* cv::BFMatcher::knnMatch(query, train, k) / match() without the reference's gates: what a
 * cv::DescriptorMatcher subclass needs (zenslam_cuda/source/bf_matcher.cpp) so zenslam::matcher's own code
 * (matcher.cpp:60-80, 172-190) runs unchanged on top.  k = 1 or 2; cross_check requires k = 1.
 * idx/dist are [nq][k]; idx = -1 where no neighbour exists (nt < k, or rejected by the cross check). */
ZS_API zs_status zs_knn_match_host(zs_context* ctx, const void* q, int nq, const void* t, int nt, int dim,
                                   int norm, int k, int cross_check, int* idx, float* dist);

/* keypoint_tracker::assign_landmark_indices, descriptor stage (zenslam_core/source/tracking/keypoint_tracker.cpp:262-287;
 * SURVEY 8 f2): cv::BFMatcher(NORM_HAMMING, crossCheck=true).match(new keypoints, landmarks) gated by
 * distance <= max_descriptor_distance.  n keypoint rows against m landmark rows (m can be 10^4..10^5: the train side is
 * split across blocks).  landmark_row[i] = matched landmark row or -1; the radius search that selects the landmarks
 * and the index bookkeeping stay on the host. */
ZS_API zs_status zs_assign_landmarks_host(zs_context* ctx, const uint8_t* keypoint_desc, int n, const uint8_t* landmark_desc,
                                          int m, double max_descriptor_distance, int* landmark_row, float* distance);

/* utils::match_keypoints3d (zenslam_core/source/matching/matching_utils.cpp:132-216, and the overload with frustum culling
 * :218-343), SURVEY 8 a9 / f2: the landmarks within `radius` of the camera (same radius_search quirk as above: the first
 * `count` landmarks in insertion order), moved into the camera frame (pose_of_camera0_in_world.inv(): R row-major 3x3, t[3]),
 * kept when z > 0 (and, with image_width > 0, when they project inside the image +- frustum_margin), matched by
 * cv::BFMatcher(NORM_HAMMING, crossCheck = true).match(landmarks, keypoints) and gated by the reprojection error
 * ||project(P, X) - kp.pt|| < threshold (P row-major 3x4).  keypoints: the caller passes map::values_unmatched(points3d_world)
 * -- the keypoints whose index is not a landmark index -- in key order.  Outputs (sized min(n_landmarks, n_keypoints)):
 * landmark_index / keypoint_index = DMatch::queryIdx / trainIdx as the reference re-keys them, error = DMatch::distance. */
ZS_API zs_status zs_match_keypoints3d_host(zs_context* ctx, const int* landmark_index, const double* landmark_xyz,
                                           const uint8_t* landmark_desc, int n_landmarks, const int* keypoint_index,
                                           const float* keypoint_xy, const uint8_t* keypoint_desc, int n_keypoints,
                                           const double* R, const double* t, const double* P, double radius, double threshold,
                                           int image_width, int image_height, double frustum_margin, int* out_landmark_index,
                                           int* out_keypoint_index, float* out_error, int* n_out);

/* ---- stereo triangulation with its gates: triangulator::triangulate_keypoints ----------------------------------
 * (zenslam_core/source/mapping/triangulator.cpp:39-132, filter_epipolar :152-188; cv::triangulatePoints behind
 * utils::triangulate_points, mapping/triangulation_utils.cpp:135-160).  SURVEY 8(f3).  n matched pairs (same keypoint
 * index in both cameras; the matching by index is host bookkeeping).  P0 / P1: 3x4 projection matrices, F: 3x3
 * fundamental matrix (calibration.fundamental_matrix[0]; NULL = no epipolar filter), t: translation of camera 1 in
 * camera 0 -- all HOST pointers, row-major doubles.  Outputs: xyz [n][3] doubles (every pair, like points3d_all),
 * keep [n] (the pairs that survive every gate), optional diag [n][4] = epipolar error, reprojection errors, angle.
 * Parity is TOLERANCE-based here, unlike the integer rows of the path: FP64 one-sided Jacobi SVD on the device vs OpenCV's on
 * the host -- points within 1e-5 relative of cv::triangulatePoints (>= 98 % identical floats), keep flags identical except
 * within 1e-6 of a gate threshold (tests/test_gpu_frontend.py::test_triangulator_mirror_tolerance_1e5_relative). */
typedef struct {
    int filter_epipolar;             /* triangulation.filter_epipolar */
    double epipolar_threshold;       /* triangulation.epipolar_threshold (0.01) */
    double reprojection_threshold;   /* triangulation.reprojection_threshold (1.0) */
    double min_depth, max_depth;     /* triangulation.min_depth / max_depth (1.0 / 50.0) */
} zs_triangulation_params;
ZS_API zs_status zs_triangulate_keypoints(zs_context* ctx, const double* P0, const double* P1, const double* F, const double* t,
                                          const float* d_pts0, const float* d_pts1, int n, const zs_triangulation_params* params,
                                          double* d_xyz, uint8_t* d_keep, double* d_diag);
ZS_API zs_status zs_triangulate_keypoints_host(zs_context* ctx, const double* P0, const double* P1, const double* F,
                                               const double* t, const float* pts0, const float* pts1, int n,
                                               const zs_triangulation_params* params, double* xyz, uint8_t* keep, double* diag);

/* ---- per-frame stereo tracker: keypoint_tracker::track -------------------------------------------------------
 * (zenslam_core/source/tracking/keypoint_tracker.cpp:41-105) for S independent stereo sequences in lock-step (S = 1: the
 * reference's own use), one call per stereo frame, state
 * (previous pyramids and the two index-keyed keypoint maps) kept on the device: temporal forward-backward KLT of both
 * cameras, grid detection behind the occupancy of the tracked keypoints, stereo tracks L -> R / R -> L of the keypoints
 * the other camera lacks, sequential keypoint indices (keypoint::index_next).  Algorithm GRID or PARALLEL_GRID, feature
 * FAST, descriptor ORB.  Host-side pieces of the reference's flow: the landmark projection behind the initial flow (its
 * result comes in through zs_tracker_set_predictions) and the RANSAC that estimates F for filter_epipolar (the gate itself
 * is zs_tracker_filter_epipolar).  assign_landmark_indices (keypoint_tracker.cpp:55,71,199-291) runs inside the step, between
 * each detection and its map add, against a landmark store kept on the device (zs_tracker_landmarks_add_host; empty store =
 * the reference's early return).  Results: both maps in key (index) order. */
typedef struct zs_tracker zs_tracker;
typedef struct {
    int width, height;
    int cell_w, cell_h, fast_threshold;            /* detection.cell_size, detection.fast_threshold */
    int klt_win_w, klt_win_h, klt_max_level;       /* tracking.klt_window_size, tracking.klt_max_level */
    double klt_threshold;                          /* tracking.klt_threshold */
    int capacity;                                  /* keypoints per camera; 0 = 4 x cells + 64 */
    int first_index;                               /* keypoint::index_next when the sequence starts */
    int sequences;                                 /* independent stereo sequences tracked in lock-step; 0 = 1 */
    int parallel_grid;                             /* detection.algorithm PARALLEL_GRID: cornerSubPix on the new corners */
    int landmark_capacity;                         /* landmarks per sequence the device store can hold; 0 = no landmark association */
    double landmark_match_radius;                  /* tracking.landmark_match_radius (50.0); <= 0: every landmark is a candidate */
    double landmark_match_distance;                /* tracking.landmark_match_distance (32.0) */
} zs_tracker_options;
typedef struct {                                   /* HOST pointers, S = sequences; any may be NULL */
    int cap;                                       /* row length of the arrays below, >= zs_tracker_capacity() */
    int* n;                                        /* [S][2] keypoints per sequence and camera */
    int* index[2]; float* xy[2]; float* response[2]; uint8_t* desc[2];   /* per camera: [S][cap], [S][cap][2], [S][cap], [S][cap][32] */
    int* next_index;                               /* [S] keypoint::index_next of every sequence after this frame */
} zs_tracker_results;
ZS_API zs_status zs_tracker_create(zs_context* ctx, const zs_tracker_options* opt, zs_tracker** out);
ZS_API void zs_tracker_destroy(zs_tracker* t);
ZS_API int zs_tracker_capacity(const zs_tracker* t);
ZS_API int zs_tracker_sequences(const zs_tracker* t);
/* optional, before a track call: predicted positions in the NEXT frame for some of camera's current keypoints, by
 * keypoint index (strictly ascending) -- the landmark projections keypoint_tracker.cpp:361-373 feeds to
 * OPTFLOW_USE_INITIAL_FLOW.  Keypoints without a prediction start from their own position.  Consumed by the next call. */
ZS_API zs_status zs_tracker_set_predictions(zs_tracker* t, int sequence, int camera, const int* index, const float* xy,
                                            int n);
/* The landmark map of `sequence` (system.points3d) as assign_landmark_indices reads it: `system.points3d += pose * points3d`
 * (slam_thread.cpp:210; map::operator+=(const map&), types/map.h:222-236): landmarks whose index the store already holds are
 * skipped, the others are appended in the order given (the reference iterates a std::map: pass them in ascending index order).
 * xyz = world coordinates [n][3], desc [n][32].  Returns ZS_ERR_CAPACITY when the store would exceed landmark_capacity.
 * zs_tracker_set_camera_center: frame_0.pose.translation() for the next step's radius search (keypoint_tracker.cpp:56,72,213);
 * default (0, 0, 0).  One reference quirk is reproduced: point3d_cloud::radius_search counts the landmarks within the radius
 * and then returns the FIRST `count` landmarks in insertion order (types/point3d_cloud.cpp:56-64), so those are the candidates. */
ZS_API zs_status zs_tracker_landmarks_add_host(zs_tracker* t, int sequence, const int* index, const double* xyz,
                                               const uint8_t* desc, int n, int* n_added);
ZS_API int zs_tracker_landmarks_size(const zs_tracker* t, int sequence);
ZS_API zs_status zs_tracker_set_camera_center(zs_tracker* t, int sequence, const double* center);
/* left / right: the new frame of every sequence, [S][height][pitch] with `stride` bytes between sequences (0 = pitch * height) */
ZS_API zs_status zs_tracker_track_host(zs_tracker* t, const uint8_t* left, const uint8_t* right, size_t pitch,
                                       size_t stride, const zs_tracker_results* res);
/* the same step with the frames already on the device (enqueued on the context's stream, no host synchronisation), and the
 * copy of the current maps to the host as a separate call -- for callers that keep frames resident or pipeline transfers */
ZS_API zs_status zs_tracker_track(zs_tracker* t, const uint8_t* d_left, const uint8_t* d_right, size_t pitch, size_t stride);
ZS_API zs_status zs_tracker_download(zs_tracker* t, const zs_tracker_results* res);
/* Capacity: a step in which a map would exceed zs_tracker_capacity() keeps the first `capacity` keypoints of that map (the
 * rest are dropped; keypoint::index_next still advances past them) and the download / wait that returns THAT step reports
 * ZS_ERR_CAPACITY after filling the results -- the step counts as consumed.  Later steps report it again only if they
 * overflow themselves. */
/* filter_epipolar (keypoint_tracker.cpp:293-341) on the maps of the last step of `sequence`, F (row-major 3x3) from the
 * caller -- the reference estimates it with cv::findFundamentalMat RANSAC on the matched points, which stays on the CPU:
 * both maps keep only the keypoints present in both whose |pt0^T F pt1| < threshold.  The filtered maps are what the next
 * step tracks from; read them back with zs_tracker_download. */
ZS_API zs_status zs_tracker_filter_epipolar(zs_tracker* t, int sequence, const double* F, double threshold);
/* pipelined host path: up to two steps in flight (copy-in of step k+1 | step k | copy-out of step k-1 on three streams).
 * `res` names where the maps of THIS step go (whole rows are copied; entries past n are unspecified); the arrays and the
 * frames must stay valid until the zs_tracker_wait that returns this step, and should be pinned for the copies to overlap.
 * zs_tracker_wait returns the oldest step in flight (filling its n / next_index). */
ZS_API zs_status zs_tracker_submit_host(zs_tracker* t, const uint8_t* left, const uint8_t* right, size_t pitch,
                                        size_t stride, const zs_tracker_results* res);
ZS_API zs_status zs_tracker_wait(zs_tracker* t);
ZS_API int zs_tracker_in_flight(const zs_tracker* t);

/* ---- batched stereo front-end ------------------------------------------------------------------
 * The per-frame call pattern of keypoint_tracker::track (keypoint_tracker.cpp:41-105) restated for
 * a batch of B consecutive stereo frames: per frame 2 pyramids, 2 grid detections + ORB, 1 stereo
 * kNN-ratio match, 4 forward+backward KLT pairs (temporal L, temporal R from the previous frame's
 * keypoints; stereo L->R, R->L).  The previous frame of the first batch element is carried from
 * the preceding call.  Field names follow options.yaml. */
typedef struct {
    int width, height, batch;
    int cell_w, cell_h;            /* slam.detection.cell_size */
    int fast_threshold;            /* slam.detection.fast_threshold */
    int klt_win_w, klt_win_h;      /* slam.tracking.klt_window_size */
    int klt_max_level;             /* slam.tracking.klt_max_level */
    double klt_threshold;          /* slam.tracking.klt_threshold */
    double matcher_ratio;          /* slam.matcher_ratio */
    int max_iters; double epsilon; double min_eig_threshold;
    int parallel_grid;             /* slam.detection.algorithm == PARALLEL_GRID: cv::cornerSubPix on the grid corners
                                      (keypoint_detector_parallel.cpp:160-170); 0 = GRID */
} zs_frontend_options;

typedef struct {               /* per-frame result views, all HOST pointers, [batch][cap] rows */
    int cap;                   /* = grid_w*grid_h */
    int* n_left; int* n_right;                 /* [batch] keypoints after ORB's border filter */
    float* kp_left; float* kp_right;           /* [batch][cap][2] x,y */
    float* resp_left; float* resp_right;       /* [batch][cap] */
    uint8_t* desc_left; uint8_t* desc_right;   /* [batch][cap][32] */
    int* match_idx; float* match_dist; uint8_t* match_pass;  /* stereo kNN: [batch][cap][2], [batch][cap][2], [batch][cap] */
    /* KLT pairs, kind 0 temporal-L, 1 temporal-R, 2 stereo L->R, 3 stereo R->L */
    float* track_pts;          /* [4][batch][cap][2] forward result */
    uint8_t* track_keep;       /* [4][batch][cap] FB gate */
    int* track_n;              /* [4][batch] number of points tracked */
} zs_frontend_results;

ZS_API zs_status zs_frontend_create(zs_context* ctx, const zs_frontend_options* opt, zs_frontend** out);
ZS_API void zs_frontend_destroy(zs_frontend* fe);
ZS_API int zs_frontend_capacity(const zs_frontend* fe);
ZS_API size_t zs_frontend_h2d_bytes(const zs_frontend* fe);
ZS_API size_t zs_frontend_d2h_bytes(const zs_frontend* fe);
/* stage B stereo frames (left/right [batch][height][pitch]) into the device ring; host or device source */
ZS_API zs_status zs_frontend_upload(zs_frontend* fe, const uint8_t* left, const uint8_t* right, size_t pitch,
                                    size_t stride, int src_is_host);
/* run the whole hot path on the staged batch (asynchronous) */
ZS_API zs_status zs_frontend_run(zs_frontend* fe);
/* copy results to host (synchronises) */
ZS_API zs_status zs_frontend_download(zs_frontend* fe, const zs_frontend_results* res);
/* per-stage device timing with CUDA events on the context's stream: stages are
 * 0 pyramid, 1 FAST/grid, 2 ORB, 3 match, 4 KLT, 5 carry.  enable(1) resets; collect() synchronises and
 * returns summed milliseconds per stage over the (at most 64 most recent) runs since enable. */
#define ZS_FRONTEND_STAGES 6
ZS_API zs_status zs_frontend_timing_enable(zs_frontend* fe, int on);
ZS_API zs_status zs_frontend_timing_collect(zs_frontend* fe, float* stage_ms_sum, int* runs);
/* upload + run + download with host buffers: the end-to-end call */
ZS_API zs_status zs_frontend_process_host(zs_frontend* fe, const uint8_t* left, const uint8_t* right,
                                          size_t pitch, size_t stride, const zs_frontend_results* res);

/* pipelined end-to-end path: submit returns at once; up to two batches are in flight on three streams (H2D of
 * batch k+1, kernels of batch k, D2H of batch k-1 overlap).  `res` buffers (ideally pinned) must stay valid until
 * the matching zs_frontend_wait returns; waits complete in submission order. */
ZS_API zs_status zs_frontend_submit_host(zs_frontend* fe, const uint8_t* left, const uint8_t* right,
                                         size_t pitch, size_t stride, const zs_frontend_results* res);
ZS_API zs_status zs_frontend_wait(zs_frontend* fe);
/* raw camera frames in: run processor::process's image path (processor.cpp:25-55) on the device in front of the pyramid
 * build -- channels 3 = BGR (COLOR_BGR2GRAY), optional CLAHE(clip), optional rectification with the calibration's
 * CV_32FC1 maps (HOST pointers, width*height floats each; all four or none).  Affects zs_frontend_submit_host and
 * zs_frontend_process_host, whose left / right buffers then hold `channels` bytes per pixel. */
ZS_API zs_status zs_frontend_set_preprocess(zs_frontend* fe, int channels, int clahe_enabled, double clahe_clip_limit,
                                            const float* map_x_left, const float* map_y_left, const float* map_x_right,
                                            const float* map_y_right);
ZS_API int zs_frontend_in_flight(const zs_frontend* fe);

#ifdef __cplusplus
}
#endif
#endif /* ZENSLAM_CUDA_H */
