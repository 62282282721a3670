#pragma once

#include "zenslam/detection/detection_options.h"
#include "zenslam/detection/keypoint_detector.h"

namespace zenslam::cuda
{
    /** GPU twin of keypoint_detector_grid (zenslam_core/source/detection/keypoint_detector_grid.cpp:39-150):
     *  occupancy grid from the existing keypoints, FAST-9-16 + NMS per free cell, first strongest corner per
     *  cell, ORB descriptors, sequential keypoint::index_next indices.  Supports feature FAST + descriptor ORB
     *  (the default options); anything else throws std::invalid_argument at construction. */
    class keypoint_detector_cuda final : public keypoint_detector
    {
    public:
        explicit keypoint_detector_cuda(const detection_options& options);

        [[nodiscard]] std::vector<keypoint> detect_keypoints(const cv::Mat& image, const map<keypoint>& keypoints_existing) const override;

    private:
        detection_options _options = { };
    };
}
