#pragma once

#include "zenslam/detection/detection_options.h"
#include "zenslam/detection/keypoint_detector.h"

namespace zenslam::cuda
{
    /** GPU twin of the reference's three detectors, selected by detection_options::algorithm like
     *  keypoint_tracker.cpp:27-38 does:
     *    GRID          keypoint_detector_grid.cpp:39-150 -- occupancy grid from the existing keypoints, FAST-9-16 + NMS
     *                  per free cell, first strongest corner per cell, ORB descriptors
     *    PARALLEL_GRID keypoint_detector_parallel.cpp:40-193 -- the same cells + cv::cornerSubPix
     *    SIMPLE        keypoint_detector_simple.cpp:38-63 -- full-frame detector behind a disc mask; `feature: FAST`
     *                  or `feature: ORB` (cv::ORB::create(500, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31, fast_threshold))
     *  Descriptor ORB only; sequential keypoint::index_next indices.  Anything else throws std::invalid_argument at
     *  construction.  With `feature: ORB` the keypoints come in canonical order (octave, y, x): OpenCV's own order is
     *  whatever std::nth_element leaves. */
    class keypoint_detector_cuda final : public keypoint_detector
    {
    public:
        explicit keypoint_detector_cuda(const detection_options& options);

        [[nodiscard]] std::vector<keypoint> detect_keypoints(const cv::Mat& image, const map<keypoint>& keypoints_existing) const override;

    private:
        [[nodiscard]] std::vector<keypoint> detect_simple(const cv::Mat& image, const map<keypoint>& keypoints_existing) const;

        detection_options _options = { };
    };
}
