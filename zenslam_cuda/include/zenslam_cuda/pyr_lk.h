#pragma once

// CUDA counterpart of zenslam_metal/include/zenslam_metal/pyr_lk.h: same free functions, same signature.

#include <vector>

#include <opencv2/core.hpp>

namespace zenslam::cuda
{
    /** true when libzenslam_cuda.so found an sm_100 device (cf. zenslam::metal::is_available) */
    auto is_available() -> bool;

    /** cv::calcOpticalFlowPyrLK semantics on the GPU.  `prev_pyramid` / `next_pyramid` are what
     *  utils::pyramid returns (cv::buildOpticalFlowPyramid output); only level 0 is read -- the device
     *  rebuilds the identical pyramid and Scharr planes itself.  Throws cv::Exception on failure, like OpenCV. */
    void calc_optical_flow_pyr_lk(
        const std::vector<cv::Mat>& prev_pyramid,
        const std::vector<cv::Mat>& next_pyramid,
        const std::vector<cv::Point2f>& prev_points,
        std::vector<cv::Point2f>& next_points,
        std::vector<uchar>& status,
        std::vector<float>& err,
        cv::Size win_size,
        int max_level,
        cv::TermCriteria criteria,
        int flags,
        double min_eig_threshold = 1e-4);

    /** keypoint_tracker::track_keypoints' two pyr_lk calls and its forward-backward gate (keypoint_tracker.cpp:129-197,
     *  :343-434) as one device call: forward LK from `points_0` (with `predicted_1` as initial flow when it is not empty),
     *  backward LK from the results, keep[i] = both statuses set and ||p0_back - p0|| < klt_threshold.  Identical results;
     *  each frame is uploaded and its pyramid built once instead of twice.  Optional: the pyr_lk seam alone is enough for a
     *  drop-in, this is the shortcut a maintainer can take inside track_keypoints. */
    void track_keypoints_fb(
        const std::vector<cv::Mat>& pyramid_0,
        const std::vector<cv::Mat>& pyramid_1,
        const std::vector<cv::Point2f>& points_0,
        const std::vector<cv::Point2f>& predicted_1,
        std::vector<cv::Point2f>& points_1,
        std::vector<uchar>& keep,
        cv::Size win_size,
        int max_level,
        double klt_threshold,
        cv::TermCriteria criteria = { cv::TermCriteria::COUNT | cv::TermCriteria::EPS, 99, 0.001 },
        double min_eig_threshold = 1e-4);
}
