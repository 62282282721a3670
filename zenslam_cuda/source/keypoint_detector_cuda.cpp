#include "zenslam_cuda/keypoint_detector_cuda.h"

#include <algorithm>
#include <ranges>
#include <stdexcept>

#include <opencv2/imgproc.hpp>

#include "context.h"

zenslam::cuda::keypoint_detector_cuda::keypoint_detector_cuda(const detection_options& options) :
    _options { options }
{
    // keypoint_detector_grid.cpp:12-36 (same switch in _parallel / _simple): FAST or ORB as the detector, ORB as the
    // describer; SIFT / FREAK stay on the CPU classes
    if ((options.feature_detector != feature_type::FAST && options.feature_detector != feature_type::ORB) || options.descriptor != descriptor_type::ORB)
        throw std::invalid_argument("keypoint_detector_cuda: feature FAST or ORB with descriptor ORB runs on the GPU");

    // cv::ORB::detect on a <= 62 px cell ROI returns nothing (31 px edge threshold): `feature: ORB` only makes sense with SIMPLE
    if (options.feature_detector == feature_type::ORB && options.algorithm != detection_algorithm::SIMPLE)
        throw std::invalid_argument("keypoint_detector_cuda: feature ORB needs algorithm SIMPLE");

    if (detail::context() == nullptr)
        throw std::runtime_error("keypoint_detector_cuda: no sm_100 device (there is no CPU fallback in this backend)");
}

std::vector<zenslam::keypoint> zenslam::cuda::keypoint_detector_cuda::detect_keypoints(const cv::Mat& image, const map<keypoint>& keypoints_existing) const
{
    CV_Assert(image.type() == CV_8UC1);

    if (_options.algorithm == detection_algorithm::SIMPLE)
        return detect_simple(image, keypoints_existing);

    const auto cell   = _options.cell_size;
    const auto grid_w = image.cols / cell.width;
    const auto grid_h = image.rows / cell.height;
    const auto cells  = static_cast<size_t>(grid_w) * static_cast<size_t>(grid_h);

    if (cells == 0)
        return { };

    // occupancy, cell row-major (keypoint_detector_grid.cpp:47-64: truncating cast, then integer division)
    std::vector<uchar> occupied(cells, 0);

    for (const auto& existing : keypoints_existing | std::views::values)
    {
        const auto grid_x = static_cast<int>(existing.pt.x) / cell.width;
        const auto grid_y = static_cast<int>(existing.pt.y) / cell.height;

        if (grid_x >= 0 && grid_x < grid_w && grid_y >= 0 && grid_y < grid_h)
            occupied[static_cast<size_t>(grid_y) * grid_w + grid_x] = 1;
    }

    std::vector<float> x(cells), y(cells), response(cells);
    cv::Mat            descriptors(static_cast<int>(cells), 32, CV_8UC1);
    int                count = 0;

    {
        std::scoped_lock lock { detail::context_mutex() };

        // GRID: keypoint_detector_grid.cpp:39-150; PARALLEL_GRID adds cv::cornerSubPix (keypoint_detector_parallel.cpp:160-170)
        const auto entry = _options.algorithm == detection_algorithm::PARALLEL_GRID ? zs_detect_keypoints_parallel_host : zs_detect_keypoints_grid_host;

        detail::check
        (
            entry
            (
                detail::context(),
                image.data,
                image.cols,
                image.rows,
                image.step,
                cell.width,
                cell.height,
                _options.fast_threshold,
                occupied.data(),
                x.data(),
                y.data(),
                response.data(),
                descriptors.data,
                &count
            ),
            "zs_detect_keypoints_grid_host"
        );
    }

    descriptors = descriptors.rowRange(0, count);

    std::vector<keypoint> keypoints { };
    keypoints.reserve(count);

    for (auto i = 0; i < count; ++i)
    {
        // what cv::FAST emits: size 7, angle -1, response = score, octave 0, class_id -1; descriptor is a row
        // view into the shared matrix, as in keypoint_detector_grid.cpp:144
        const cv::KeyPoint keypoint_cv { x[i], y[i], 7.0f, -1.0f, response[i], 0, -1 };

        keypoints.emplace_back(keypoint_cv, keypoint::index_next, descriptors.row(i));

        keypoint::index_next++;
    }

    return keypoints;
}

std::vector<zenslam::keypoint> zenslam::cuda::keypoint_detector_cuda::detect_simple(const cv::Mat& image, const map<keypoint>& keypoints_existing) const
{
    // keypoint_detector_simple.cpp:41-48: 255 everywhere, filled discs of radius min(cell) / 2 at the existing keypoints
    cv::Mat mask { image.size(), CV_8UC1, cv::Scalar(255) };

    for (const auto& existing : keypoints_existing | std::views::values)
    {
        cv::circle(mask, existing.pt, std::min(_options.cell_size.width, _options.cell_size.height) / 2, cv::Scalar(0), -1);
    }

    const auto is_orb = _options.feature_detector == feature_type::ORB;

    // ORB: nfeatures 500 plus the ties retainBest keeps; FAST: the reference puts no cap on the count, one corner per
    // 2 x 2 pixels is the bound non-maximum suppression leaves
    const auto capacity = is_orb ? 500 + 32 * 8 + 64 : ((image.cols + 1) / 2) * ((image.rows + 1) / 2);

    std::vector<float> x(capacity), y(capacity), response(capacity), size(capacity, 7.0f), angle(capacity, -1.0f);
    std::vector<int>   octave(capacity, 0);
    cv::Mat            descriptors(capacity, 32, CV_8UC1);
    int                count = 0;

    {
        std::scoped_lock lock { detail::context_mutex() };

        if (is_orb)
        {
            // cv::ORB::create(500, 1.2f, 8, 31, 0, 2, cv::ORB::HARRIS_SCORE, 31, fast_threshold) (keypoint_detector_simple.cpp:17)
            detail::check
            (
                zs_detect_keypoints_orb_host
                (
                    detail::context(), image.data, image.cols, image.rows, image.step, mask.data, mask.step,
                    500, 1.2f, 8, 31, 31, _options.fast_threshold,
                    x.data(), y.data(), size.data(), angle.data(), response.data(), octave.data(), descriptors.data, capacity, &count
                ),
                "zs_detect_keypoints_orb_host"
            );
        }
        else
        {
            detail::check
            (
                zs_detect_keypoints_simple_host
                (
                    detail::context(), image.data, image.cols, image.rows, image.step, mask.data, mask.step, _options.fast_threshold,
                    x.data(), y.data(), response.data(), descriptors.data, capacity, &count
                ),
                "zs_detect_keypoints_simple_host"
            );
        }
    }

    std::vector<keypoint> keypoints { };
    keypoints.reserve(count);

    for (auto i = 0; i < count; ++i)
    {
        const cv::KeyPoint keypoint_cv { x[i], y[i], size[i], angle[i], response[i], octave[i], -1 };

        // keypoint_detector_simple.cpp:59 clones the descriptor row
        keypoints.emplace_back(keypoint_cv, keypoint::index_next, descriptors.row(i).clone());

        keypoint::index_next++;
    }

    return keypoints;
}
