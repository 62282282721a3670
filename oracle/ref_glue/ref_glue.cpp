// TEST INFRASTRUCTURE (oracle/ref_glue): a C entry over the REFERENCE's detector classes, whose sources are compiled from where
// they lie under /root/reference (oracle/build_ref.py) -- zenslam::keypoint_detector_grid / _parallel / _simple with their own
// headers, on the OpenCV stand-in backed by the C oracle (opencv_over_oracle.cpp).
#include <cstdint>
#include <cstring>
#include <memory>

#include "zenslam/detection/keypoint_detector_grid.h"
#include "zenslam/detection/keypoint_detector_parallel.h"
#include "zenslam/detection/keypoint_detector_simple.h"

size_t zenslam::keypoint::index_next = 0;        // types/keypoint.cpp of the reference

// algorithm: 0 SIMPLE, 1 GRID, 2 PARALLEL_GRID (detection_algorithm.h); feature FAST, descriptor ORB.
// existing: n_existing rows of (index, x, y).  Outputs sized cap: xy [cap][2], response, size, angle [cap], octave, index [cap],
// desc [cap][32].  Returns the number of keypoints (or -1 - needed when cap is too small); *index_next is keypoint::index_next
// before (in) and after (out) the call.
extern "C" __attribute__((visibility("default")))
int zref_detect_keypoints(int algorithm, const uint8_t* img, int w, int h, int pitch, int cell_w, int cell_h, int fast_threshold,
                          const float* existing, int n_existing, long long* index_next, float* xy, float* response, float* size,
                          float* angle, int* octave, long long* index, uint8_t* desc, int cap)
{
    zenslam::detection_options options { };
    options.cell_size        = cv::Size(cell_w, cell_h);
    options.fast_threshold   = fast_threshold;
    options.feature_detector = zenslam::feature_type::FAST;
    options.descriptor       = zenslam::descriptor_type::ORB;
    options.algorithm        = algorithm == 0 ? zenslam::detection_algorithm::SIMPLE
                             : algorithm == 1 ? zenslam::detection_algorithm::GRID : zenslam::detection_algorithm::PARALLEL_GRID;

    // the switch of keypoint_tracker.cpp:27-38
    std::unique_ptr<zenslam::keypoint_detector> detector;
    if (algorithm == 0) detector = std::make_unique<zenslam::keypoint_detector_simple>(options);
    else if (algorithm == 1) detector = std::make_unique<zenslam::keypoint_detector_grid>(options);
    else detector = std::make_unique<zenslam::keypoint_detector_parallel>(options);

    zenslam::map<zenslam::keypoint> map_existing;
    for (int i = 0; i < n_existing; ++i)
    {
        zenslam::keypoint k { };
        k.index = static_cast<size_t>(existing[3 * i]);
        k.pt    = cv::Point2f(existing[3 * i + 1], existing[3 * i + 2]);
        map_existing.add(k);
    }

    const cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img), static_cast<size_t>(pitch));

    zenslam::keypoint::index_next = static_cast<size_t>(*index_next);
    const auto keypoints          = detector->detect_keypoints(image, map_existing);
    *index_next                   = static_cast<long long>(zenslam::keypoint::index_next);

    const int n = static_cast<int>(keypoints.size());
    if (n > cap) return -1 - n;
    for (int i = 0; i < n; ++i)
    {
        const auto& k = keypoints[i];
        xy[2 * i] = k.pt.x; xy[2 * i + 1] = k.pt.y;
        response[i] = k.response; size[i] = k.size; angle[i] = k.angle; octave[i] = k.octave;
        index[i] = static_cast<long long>(k.index);
        std::memcpy(desc + 32 * static_cast<size_t>(i), k.descriptor.ptr<uint8_t>(0), 32);
    }
    return n;
}
