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
}
