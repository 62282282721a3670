// TEST INFRASTRUCTURE (oracle/ref_glue): stands in for the reference's zenslam/utils/utils_opencv.h (which includes the frame,
// calibration and viz types).  The detector sources use exactly one thing from it: element-wise division of two cv::Size
// (utils_opencv.h:92-95), restated here.
#pragma once
// what the detector sources get through the real header's include chain
#include <algorithm>
#include <future>
#include <optional>
#include <ranges>
#include <vector>

#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>
inline cv::Size operator/(const cv::Size& a, const cv::Size& b) { return cv::Size(a.width / b.width, a.height / b.height); }
