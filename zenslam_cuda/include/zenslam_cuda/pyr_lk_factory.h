#pragma once

#include <memory>

#include "zenslam/tracking/pyr_lk.h"

namespace zenslam::cuda
{
    /** Returns an empty pointer when no B200 is usable, so the caller falls back exactly as it does for
     *  Metal (zenslam_app/source/application.cpp:14-15, zenslam_core/source/slam_thread.cpp:32-35). */
    auto create_cuda_pyr_lk() -> std::shared_ptr<zenslam::pyr_lk>;
}
