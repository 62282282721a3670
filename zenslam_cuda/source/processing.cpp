#include "zenslam_cuda/processing.h"

#include "context.h"

auto zenslam::cuda::process_image(const cv::Mat& image, const bool clahe_enabled, const double clahe_clip_limit, const cv::Mat& map_x, const cv::Mat& map_y) -> cv::Mat
{
    CV_Assert(image.type() == CV_8UC3 || image.type() == CV_8UC1);
    CV_Assert(map_x.empty() == map_y.empty());

    const auto channels = image.type() == CV_8UC3 ? 3 : 1;

    cv::Mat mx = map_x;
    cv::Mat my = map_y;

    if (!mx.empty())
    {
        // utils::rectify remaps into an image of the maps' size, which the calibration makes the image's own (calibration.cpp:60-70)
        CV_Assert(mx.type() == CV_32FC1 && my.type() == CV_32FC1 && mx.size() == image.size() && my.size() == image.size());

        if (!mx.isContinuous()) mx = mx.clone();
        if (!my.isContinuous()) my = my.clone();
    }

    cv::Mat undistorted(image.rows, image.cols, CV_8UC1);

    // processor::process runs this for both cameras on two threads at once: each call takes its own context
    const detail::preprocessing_lease lease { };

    if (lease.get() == nullptr)
        CV_Error(cv::Error::StsError, "zenslam::cuda::process_image: no sm_100 device (there is no CPU fallback in this backend)");

    detail::check
    (
        zs_process_image_host
        (
            lease.get(),
            image.data,
            channels,
            image.cols,
            image.rows,
            image.step,
            clahe_enabled ? 1 : 0,
            clahe_clip_limit,
            mx.empty() ? nullptr : mx.ptr<float>(0),
            my.empty() ? nullptr : my.ptr<float>(0),
            undistorted.data
        ),
        "zs_process_image_host"
    );

    return undistorted;
}

void zenslam::cuda::triangulate_points(
    const cv::Matx34d& projection_0,
    const cv::Matx34d& projection_1,
    const cv::Matx33d* fundamental,
    const cv::Vec3d& translation,
    const std::vector<cv::Point2f>& points_0,
    const std::vector<cv::Point2f>& points_1,
    const triangulation_gates& gates,
    std::vector<cv::Point3d>& points3d,
    std::vector<uchar>& keep)
{
    CV_Assert(points_0.size() == points_1.size());

    const auto count = points_0.size();

    points3d.assign(count, cv::Point3d { });
    keep.assign(count, 0);

    if (count == 0)
        return;

    zs_triangulation_params params { };
    params.filter_epipolar        = gates.filter_epipolar && fundamental != nullptr ? 1 : 0;
    params.epipolar_threshold     = gates.epipolar_threshold;
    params.reprojection_threshold = gates.reprojection_threshold;
    params.min_depth              = gates.min_depth;
    params.max_depth              = gates.max_depth;

    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float) && sizeof(cv::Point3d) == 3 * sizeof(double));

    std::scoped_lock lock { detail::context_mutex() };

    detail::check
    (
        zs_triangulate_keypoints_host
        (
            detail::context(),
            projection_0.val,                                   // cv::Matx stores row-major, which is what the C entry takes
            projection_1.val,
            fundamental ? fundamental->val : nullptr,
            translation.val,
            reinterpret_cast<const float*>(points_0.data()),
            reinterpret_cast<const float*>(points_1.data()),
            static_cast<int>(count),
            &params,
            reinterpret_cast<double*>(points3d.data()),
            keep.data(),
            nullptr
        ),
        "zs_triangulate_keypoints_host"
    );
}
