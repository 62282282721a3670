#include "context.h"

#include <cstdlib>
#include <string>

#include <opencv2/core.hpp>

namespace
{
    struct holder
    {
        zs_context* ctx = nullptr;

        holder()
        {
            if (!zs_is_available())
                return;

            // ZENSLAM_CUDA_DEVICE selects the GPU (default 0); one process per GPU is the scaling model
            const char* env    = std::getenv("ZENSLAM_CUDA_DEVICE");
            const int   device = env ? std::atoi(env) : 0;

            if (zs_context_create(device, nullptr, &ctx) != ZS_OK)
                ctx = nullptr;
        }

        ~holder()
        {
            if (ctx)
                zs_context_destroy(ctx);
        }
    };
}

auto zenslam::cuda::detail::context() -> zs_context*
{
    static holder instance { };
    return instance.ctx;
}

auto zenslam::cuda::detail::context_mutex() -> std::mutex&
{
    static std::mutex mutex { };
    return mutex;
}

void zenslam::cuda::detail::check(const zs_status status, const char* what)
{
    if (status == ZS_OK)
        return;

    const std::string message = std::string(what) + ": " + zs_status_string(status) + ": " + zs_last_error_string();

    CV_Error(cv::Error::StsError, message);
}
