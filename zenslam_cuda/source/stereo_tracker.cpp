#include "zenslam_cuda/stereo_tracker.h"

#include <stdexcept>
#include <vector>

#include "context.h"

zenslam::cuda::stereo_tracker::stereo_tracker(const detection_options& detection, const tracking_options& tracking, const cv::Size image_size, const int landmark_capacity)
{
    if (detection.algorithm == detection_algorithm::SIMPLE || detection.feature_detector != feature_type::FAST || detection.descriptor != descriptor_type::ORB)
        throw std::invalid_argument("stereo_tracker: algorithm GRID or PARALLEL_GRID with feature FAST and descriptor ORB runs on the GPU");

    if (detail::context() == nullptr)
        throw std::runtime_error("stereo_tracker: no sm_100 device (there is no CPU fallback in this backend)");

    zs_tracker_options tracker_options { };
    tracker_options.width          = image_size.width;
    tracker_options.height         = image_size.height;
    tracker_options.cell_w         = detection.cell_size.width;
    tracker_options.cell_h         = detection.cell_size.height;
    tracker_options.fast_threshold = detection.fast_threshold;
    tracker_options.klt_win_w      = tracking.klt_window_size.width;
    tracker_options.klt_win_h      = tracking.klt_window_size.height;
    tracker_options.klt_max_level  = tracking.klt_max_level;
    tracker_options.klt_threshold  = tracking.klt_threshold;
    tracker_options.capacity       = 0;
    tracker_options.first_index    = static_cast<int>(keypoint::index_next);
    tracker_options.sequences      = 1;
    tracker_options.parallel_grid  = detection.algorithm == detection_algorithm::PARALLEL_GRID ? 1 : 0;
    tracker_options.landmark_capacity       = landmark_capacity;
    tracker_options.landmark_match_radius   = tracking.landmark_match_radius;
    tracker_options.landmark_match_distance = tracking.landmark_match_distance;

    std::scoped_lock lock { detail::context_mutex() };

    detail::check(zs_tracker_create(detail::context(), &tracker_options, &_tracker), "zs_tracker_create");

    _capacity = zs_tracker_capacity(_tracker);
}

zenslam::cuda::stereo_tracker::~stereo_tracker()
{
    std::scoped_lock lock { detail::context_mutex() };

    zs_tracker_destroy(_tracker);
}

void zenslam::cuda::stereo_tracker::set_predictions(const int camera, const std::map<size_t, cv::Point2f>& predictions)
{
    std::vector<int>   index { };
    std::vector<float> xy { };

    index.reserve(predictions.size());
    xy.reserve(2 * predictions.size());

    for (const auto& [key, point] : predictions)       // std::map iterates in ascending key order, as the C entry requires
    {
        index.push_back(static_cast<int>(key));
        xy.push_back(point.x);
        xy.push_back(point.y);
    }

    std::scoped_lock lock { detail::context_mutex() };

    detail::check(zs_tracker_set_predictions(_tracker, 0, camera, index.data(), xy.data(), static_cast<int>(index.size())), "zs_tracker_set_predictions");
}

auto zenslam::cuda::stereo_tracker::add_landmarks(const std::map<size_t, std::pair<cv::Point3d, cv::Mat>>& landmarks) -> int
{
    std::vector<int>    index { };
    std::vector<double> xyz { };
    std::vector<uchar>  descriptors { };

    for (const auto& [key, landmark] : landmarks)      // ascending key order, as map::operator+=(const map&) iterates
    {
        const auto& [point, descriptor] = landmark;

        CV_Assert(descriptor.type() == CV_8UC1 && descriptor.rows == 1 && descriptor.cols == 32);

        index.push_back(static_cast<int>(key));
        xyz.insert(xyz.end(), { point.x, point.y, point.z });
        descriptors.insert(descriptors.end(), descriptor.ptr<uchar>(0), descriptor.ptr<uchar>(0) + 32);
    }

    int added = 0;

    std::scoped_lock lock { detail::context_mutex() };

    detail::check(zs_tracker_landmarks_add_host(_tracker, 0, index.data(), xyz.data(), descriptors.data(), static_cast<int>(index.size()), &added), "zs_tracker_landmarks_add_host");

    return added;
}

void zenslam::cuda::stereo_tracker::set_camera_center(const cv::Point3d& center)
{
    const double xyz[3] = { center.x, center.y, center.z };

    std::scoped_lock lock { detail::context_mutex() };

    detail::check(zs_tracker_set_camera_center(_tracker, 0, xyz), "zs_tracker_set_camera_center");
}

namespace
{
    // host side of zs_tracker_results for one sequence, and its conversion to the reference's maps
    struct host_maps
    {
        explicit host_maps(const int capacity) :
            capacity { capacity }
        {
            for (auto camera = 0; camera < 2; ++camera)
            {
                index[camera].resize(capacity);
                xy[camera].resize(2 * static_cast<size_t>(capacity));
                response[camera].resize(capacity);
                descriptors[camera] = cv::Mat(capacity, 32, CV_8UC1);

                results.index[camera]    = index[camera].data();
                results.xy[camera]       = xy[camera].data();
                results.response[camera] = response[camera].data();
                results.desc[camera]     = descriptors[camera].data;
            }

            results.cap        = capacity;
            results.n          = count;
            results.next_index = &index_next;
        }

        [[nodiscard]] auto to_maps() const -> std::array<zenslam::map<zenslam::keypoint>, 2>
        {
            std::array<zenslam::map<zenslam::keypoint>, 2> keypoints { };

            for (auto camera = 0; camera < 2; ++camera)
            {
                for (auto i = 0; i < count[camera]; ++i)
                {
                    // FAST keypoints: size 7, angle -1, octave 0, class_id -1 (tracked copies keep everything but pt)
                    const cv::KeyPoint keypoint_cv { xy[camera][2 * i], xy[camera][2 * i + 1], 7.0f, -1.0f, response[camera][i], 0, -1 };

                    keypoints[camera].add(zenslam::keypoint { keypoint_cv, static_cast<size_t>(index[camera][i]), descriptors[camera].row(i).clone() });
                }
            }

            return keypoints;
        }

        int                capacity   = 0;
        int                count[2]   = { 0, 0 };
        int                index_next = 0;
        std::vector<int>   index[2]   = { };
        std::vector<float> xy[2]      = { };
        std::vector<float> response[2] = { };
        cv::Mat            descriptors[2] = { };
        zs_tracker_results results { };
    };
}

auto zenslam::cuda::stereo_tracker::track(const cv::Mat& undistorted_0, const cv::Mat& undistorted_1) -> std::array<map<keypoint>, 2>
{
    CV_Assert(undistorted_0.type() == CV_8UC1 && undistorted_1.type() == CV_8UC1 && undistorted_0.size() == undistorted_1.size());

    cv::Mat image_0 = undistorted_0;
    cv::Mat image_1 = undistorted_1;

    if (image_0.step != image_1.step)
    {
        image_0 = image_0.clone();
        image_1 = image_1.clone();
    }

    host_maps host { _capacity };

    {
        std::scoped_lock lock { detail::context_mutex() };

        detail::check(zs_tracker_track_host(_tracker, image_0.data, image_1.data, image_0.step, 0, &host.results), "zs_tracker_track_host");
    }

    // new keypoints took sequential indices on the device, exactly as keypoint::index_next++ would have handed them out
    keypoint::index_next = static_cast<size_t>(host.index_next);

    return host.to_maps();
}

auto zenslam::cuda::stereo_tracker::download() -> std::array<map<keypoint>, 2>
{
    host_maps host { _capacity };

    {
        std::scoped_lock lock { detail::context_mutex() };

        detail::check(zs_tracker_download(_tracker, &host.results), "zs_tracker_download");
    }

    return host.to_maps();
}

auto zenslam::cuda::stereo_tracker::filter_epipolar(const cv::Matx33d& fundamental, const double threshold) -> std::array<map<keypoint>, 2>
{
    {
        std::scoped_lock lock { detail::context_mutex() };

        // cv::Matx stores row-major, which is what the C entry takes
        detail::check(zs_tracker_filter_epipolar(_tracker, 0, fundamental.val, threshold), "zs_tracker_filter_epipolar");
    }

    return download();
}
