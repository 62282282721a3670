#include "zenslam_cuda/pyr_lk.h"

#include "context.h"

auto zenslam::cuda::is_available() -> bool
{
    return detail::context() != nullptr;
}

void zenslam::cuda::calc_optical_flow_pyr_lk(
    const std::vector<cv::Mat>& prev_pyramid,
    const std::vector<cv::Mat>& next_pyramid,
    const std::vector<cv::Point2f>& prev_points,
    std::vector<cv::Point2f>& next_points,
    std::vector<uchar>& status,
    std::vector<float>& err,
    const cv::Size win_size,
    const int max_level,
    const cv::TermCriteria criteria,
    const int flags,
    const double min_eig_threshold)
{
    const auto count = prev_points.size();

    status.assign(count, 0);
    err.assign(count, 0.0f);

    // OPTFLOW_USE_INITIAL_FLOW: next_points is in/out and must already hold one guess per point
    // (keypoint_tracker.cpp:361-373, 390); otherwise it is an output the callee sizes
    const bool use_initial = (flags & ZS_LK_USE_INITIAL_FLOW) != 0;

    if (use_initial)
    {
        CV_Assert(next_points.size() == count);
    }
    else
    {
        next_points.assign(count, cv::Point2f { });
    }

    if (count == 0)
        return;

    CV_Assert(!prev_pyramid.empty() && !next_pyramid.empty());

    // utils::pyramid output (utils_opencv.cpp:525-530): element 0 is the level-0 image, an ROI into the
    // padded buffer -> pass its data pointer and row step
    const cv::Mat& prev = prev_pyramid.front();
    const cv::Mat& next = next_pyramid.front();

    CV_Assert(prev.type() == CV_8UC1 && next.type() == CV_8UC1 && prev.size() == next.size());

    // the C entry point takes one row pitch for both images; ROIs of two same-sized padded buffers share it,
    // anything else is made continuous first
    cv::Mat prev_level0 = prev;
    cv::Mat next_level0 = next;

    if (prev_level0.step != next_level0.step)
    {
        prev_level0 = prev.clone();
        next_level0 = next.clone();
    }

    zs_lk_params params { };
    params.win_w             = win_size.width;
    params.win_h             = win_size.height;
    params.max_level         = max_level;
    params.max_iters         = (criteria.type & cv::TermCriteria::COUNT) ? criteria.maxCount : 30;
    params.epsilon           = (criteria.type & cv::TermCriteria::EPS) ? criteria.epsilon : 0.01;
    params.flags             = flags;
    params.min_eig_threshold = min_eig_threshold;

    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float));

    std::scoped_lock lock { detail::context_mutex() };

    detail::check
    (
        zs_calc_optical_flow_pyr_lk_host
        (
            detail::context(),
            prev_level0.data,
            next_level0.data,
            prev_level0.cols,
            prev_level0.rows,
            prev_level0.step,
            reinterpret_cast<const float*>(prev_points.data()),
            reinterpret_cast<float*>(next_points.data()),
            static_cast<int>(count),
            status.data(),
            err.data(),
            &params
        ),
        "zs_calc_optical_flow_pyr_lk_host"
    );
}

void zenslam::cuda::track_keypoints_fb(
    const std::vector<cv::Mat>& pyramid_0,
    const std::vector<cv::Mat>& pyramid_1,
    const std::vector<cv::Point2f>& points_0,
    const std::vector<cv::Point2f>& predicted_1,
    std::vector<cv::Point2f>& points_1,
    std::vector<uchar>& keep,
    const cv::Size win_size,
    const int max_level,
    const double klt_threshold,
    const cv::TermCriteria criteria,
    const double min_eig_threshold)
{
    const auto count = points_0.size();

    points_1.assign(count, cv::Point2f { });
    keep.assign(count, 0);

    if (count == 0)
        return;

    CV_Assert(predicted_1.empty() || predicted_1.size() == count);
    CV_Assert(!pyramid_0.empty() && !pyramid_1.empty());

    cv::Mat level0_0 = pyramid_0.front();
    cv::Mat level0_1 = pyramid_1.front();

    CV_Assert(level0_0.type() == CV_8UC1 && level0_1.type() == CV_8UC1 && level0_0.size() == level0_1.size());

    if (level0_0.step != level0_1.step)
    {
        level0_0 = level0_0.clone();
        level0_1 = level0_1.clone();
    }

    zs_lk_params params { };
    params.win_w             = win_size.width;
    params.win_h             = win_size.height;
    params.max_level         = max_level;
    params.max_iters         = (criteria.type & cv::TermCriteria::COUNT) ? criteria.maxCount : 30;
    params.epsilon           = (criteria.type & cv::TermCriteria::EPS) ? criteria.epsilon : 0.01;
    params.flags             = ZS_LK_GET_MIN_EIGENVALS;     // what the reference always passes (keypoint_tracker.cpp:153,390)
    params.min_eig_threshold = min_eig_threshold;

    std::scoped_lock lock { detail::context_mutex() };

    detail::check
    (
        zs_track_keypoints_host
        (
            detail::context(),
            level0_0.data,
            level0_1.data,
            level0_0.cols,
            level0_0.rows,
            level0_0.step,
            reinterpret_cast<const float*>(points_0.data()),
            predicted_1.empty() ? nullptr : reinterpret_cast<const float*>(predicted_1.data()),
            static_cast<int>(count),
            &params,
            klt_threshold,
            reinterpret_cast<float*>(points_1.data()),
            nullptr,
            nullptr,
            keep.data()
        ),
        "zs_track_keypoints_host"
    );
}
