#pragma once

// zs_context objects shared by the adapter classes.  The reference calls the three tracking seams (pyr_lk, keypoint_detector,
// matcher) from the single slam_thread worker, strictly sequentially (slam_thread.cpp:139 -> tracker.cpp:40,51), so ONE
// context / one stream is the faithful mapping for them -- and it keeps the LK pyramid cache of that context effective; its
// mutex only protects against a second tracker instance.  processor::process, on the other hand, converts the two camera
// images on two threads at once (processor.cpp:25-55): zenslam::cuda::process_image therefore leases one of two further
// contexts (own stream, own scratch), so that the two images do not serialise behind one mutex.

#include <mutex>

#include "zenslam_cuda.h"

namespace zenslam::cuda::detail
{
    /** the tracking context; nullptr when no sm_100 device is usable */
    auto context() -> zs_context*;
    auto context_mutex() -> std::mutex&;

    /** a pre-processing context, held for the lifetime of the lease (blocks while both are busy) */
    class preprocessing_lease
    {
    public:
        preprocessing_lease();
        ~preprocessing_lease();

        preprocessing_lease(const preprocessing_lease&)            = delete;
        preprocessing_lease& operator=(const preprocessing_lease&) = delete;

        [[nodiscard]] auto get() const -> zs_context* { return _context; }

    private:
        zs_context* _context = nullptr;
        int         _slot    = -1;
    };

    /** throws cv::Exception(StsError) carrying zs_last_error_string() when status != ZS_OK */
    void check(zs_status status, const char* what);
}
