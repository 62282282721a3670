#pragma once

// The image path of zenslam::processor::process (zenslam_core/source/processor.cpp:25-55) and the numeric core of
// zenslam::triangulator::triangulate_keypoints (zenslam_core/source/mapping/triangulator.cpp:39-132) on the GPU, as free
// functions over OpenCV types only -- the reference classes around them also carry the IMU integrator, the calibration
// object and the point-cloud containers, which stay where they are.

#include <vector>

#include <opencv2/core.hpp>

namespace zenslam::cuda
{
    /** utils::convert_color(BGR2GRAY) -> optional utils::apply_clahe (cv::createCLAHE(clip, 8x8)) -> utils::rectify
     *  (cv::remap, INTER_LINEAR, constant border) of one camera image (processor.cpp:29-36 / :45-52): `image` CV_8UC3 (BGR) or
     *  CV_8UC1, `map_x` / `map_y` the calibration's CV_32FC1 maps (empty: no rectification).  Returns
     *  frame::processed::undistorted[camera]; bit-identical to the OpenCV calls. */
    auto process_image(const cv::Mat& image, bool clahe_enabled, double clahe_clip_limit, const cv::Mat& map_x, const cv::Mat& map_y) -> cv::Mat;

    /** slam_options::triangulation keys the gates read (all_options.h:35-45) */
    struct triangulation_gates
    {
        bool   filter_epipolar        = true;
        double epipolar_threshold     = 0.01;
        double reprojection_threshold = 1.0;
        double min_depth              = 1.0;
        double max_depth              = 50.0;
    };

    /** What triangulate_keypoints does with the matched pairs (same keypoint index in both cameras, ascending): the epipolar
     *  filter (triangulator.cpp:152-188), cv::triangulatePoints behind utils::triangulate_points
     *  (mapping/triangulation_utils.cpp:135-160) and the reprojection / depth / parallax gates (triangulator.cpp:60-128).
     *  projection_0 / projection_1: calibration.projection_matrix[0 / 1]; fundamental: calibration.fundamental_matrix[0]
     *  (nullptr: no epipolar filter); translation: cameras[1].pose_in_cam0.translation().  points3d receives one point per pair
     *  (points3d_all), keep[i] != 0 marks the pairs that pass every gate.  Tolerance-based parity (FP64 SVD): see
     *  include/zenslam_cuda.h. */
    void triangulate_points(const cv::Matx34d& projection_0, const cv::Matx34d& projection_1, const cv::Matx33d* fundamental,
                            const cv::Vec3d& translation, const std::vector<cv::Point2f>& points_0, const std::vector<cv::Point2f>& points_1,
                            const triangulation_gates& gates, std::vector<cv::Point3d>& points3d, std::vector<uchar>& keep);
}
