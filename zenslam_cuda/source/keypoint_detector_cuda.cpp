#include "zenslam_cuda/keypoint_detector_cuda.h"

#include <ranges>
#include <stdexcept>

#include "context.h"

zenslam::cuda::keypoint_detector_cuda::keypoint_detector_cuda(const detection_options& options) :
    _options { options }
{
    if (options.feature_detector != feature_type::FAST || options.descriptor != descriptor_type::ORB)
        throw std::invalid_argument("keypoint_detector_cuda: only feature FAST with descriptor ORB runs on the GPU");

    if (detail::context() == nullptr)
        throw std::runtime_error("keypoint_detector_cuda: no sm_100 device (there is no CPU fallback in this backend)");
}

std::vector<zenslam::keypoint> zenslam::cuda::keypoint_detector_cuda::detect_keypoints(const cv::Mat& image, const map<keypoint>& keypoints_existing) const
{
    CV_Assert(image.type() == CV_8UC1);

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

        detail::check
        (
            zs_detect_keypoints_grid_host
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
