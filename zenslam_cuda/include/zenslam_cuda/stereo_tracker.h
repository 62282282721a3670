#pragma once

#include <array>
#include <map>
#include <memory>
#include <utility>

#include <opencv2/core.hpp>

#include "zenslam/detection/detection_options.h"
#include "zenslam/tracking_options.h"
#include "zenslam/types/keypoint.h"
#include "zenslam/types/map.h"

struct zs_tracker;

namespace zenslam::cuda
{
    /** keypoint_tracker::track (zenslam_core/source/tracking/keypoint_tracker.cpp:41-105) with its frame-to-frame state
     *  -- the previous pyramids, both keypoint maps, keypoint::index_next -- kept on the GPU: one device call per stereo
     *  frame.  Algorithm GRID, feature FAST, descriptor ORB (anything else throws at construction).  What the reference
     *  does around the KLT / detection calls with CPU-only inputs stays with the caller: the projection behind the
     *  pose-predicted initial flow (its result comes in through set_predictions) and the RANSAC of filter_epipolar
     *  (cv::findFundamentalMat; the gate itself is filter_epipolar below).  assign_landmark_indices
     *  (keypoint_tracker.cpp:55,71,199-291) runs inside track() against the landmarks handed over with add_landmarks.
     *  keypoint_tracker::track is not a virtual seam, so this class is an opt-in replacement of its body. */
    class stereo_tracker final
    {
    public:
        /** `detection` / `tracking` = slam_options::detection / ::tracking (all_options.h:111-137) */
        /** landmark_capacity: landmarks the device store can hold (0: no landmark association, i.e. system.points3d empty) */
        stereo_tracker(const detection_options& detection, const tracking_options& tracking, cv::Size image_size, int landmark_capacity = 0);
        ~stereo_tracker();

        stereo_tracker(const stereo_tracker&)            = delete;
        stereo_tracker& operator=(const stereo_tracker&) = delete;

        /** optional, before track(): predicted positions in the next frame for some keypoints of `camera`, keyed by
         *  keypoint index -- the landmark projections keypoint_tracker.cpp:361-373 uses as initial flow */
        void set_predictions(int camera, const std::map<size_t, cv::Point2f>& predictions);

        /** `system.points3d += ...` (slam_thread.cpp:210) as assign_landmark_indices reads it: landmark index -> (world position,
         *  1 x 32 ORB descriptor); indices the store already holds are skipped, the others appended in key order.  Returns the
         *  number added.  set_camera_center: frame_0.pose.translation() for the radius search of the next track()
         *  (keypoint_tracker.cpp:56,72,213). */
        auto add_landmarks(const std::map<size_t, std::pair<cv::Point3d, cv::Mat>>& landmarks) -> int;
        void set_camera_center(const cv::Point3d& center);

        /** keypoint_tracker::filter_epipolar's gate with F from the caller (cv::findFundamentalMat on the matched points of
         *  the maps track() returned): the device-side maps keep only keypoints present in both cameras with
         *  |pt0^T F pt1| < threshold; returns the filtered maps, which are also what the next track() starts from */
        [[nodiscard]] auto filter_epipolar(const cv::Matx33d& fundamental, double threshold) -> std::array<map<keypoint>, 2>;

        /** undistorted grayscale images of the new stereo frame -> the two keypoint maps of that frame */
        [[nodiscard]] auto track(const cv::Mat& undistorted_0, const cv::Mat& undistorted_1) -> std::array<map<keypoint>, 2>;

    private:
        [[nodiscard]] auto download() -> std::array<map<keypoint>, 2>;

        zs_tracker* _tracker  = nullptr;
        int         _capacity = 0;
    };
}
