#pragma once

// Process-wide zs_context shared by the adapter classes.  The reference calls all three seams from the single
// slam_thread worker, strictly sequentially (slam_thread.cpp:139 -> tracker.cpp:40,51), so one context / one
// stream is the faithful mapping; the mutex only protects against a second tracker instance.

#include <mutex>

#include "zenslam_cuda.h"

namespace zenslam::cuda::detail
{
    /** nullptr when no sm_100 device is usable */
    auto context() -> zs_context*;
    auto context_mutex() -> std::mutex&;

    /** throws cv::Exception(StsError) carrying zs_last_error_string() when status != ZS_OK */
    void check(zs_status status, const char* what);
}
