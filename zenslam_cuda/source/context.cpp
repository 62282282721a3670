#include "context.h"

#include <condition_variable>
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

namespace
{
    // two pre-processing contexts (one per camera image thread of processor::process), created on first use
    struct preprocessing_pool
    {
        static constexpr int    size = 2;
        holder*                 contexts[size] = { nullptr, nullptr };
        bool                    busy[size]     = { false, false };
        std::mutex              mutex { };
        std::condition_variable released { };

        ~preprocessing_pool()
        {
            for (auto* context : contexts)
                delete context;
        }
    };

    auto pool() -> preprocessing_pool&
    {
        static preprocessing_pool instance { };
        return instance;
    }
}

zenslam::cuda::detail::preprocessing_lease::preprocessing_lease()
{
    auto& p = pool();

    std::unique_lock lock { p.mutex };

    p.released.wait(lock, [&p] { return !p.busy[0] || !p.busy[1]; });

    _slot         = p.busy[0] ? 1 : 0;
    p.busy[_slot] = true;

    if (p.contexts[_slot] == nullptr)
        p.contexts[_slot] = new holder { };

    _context = p.contexts[_slot]->ctx;
}

zenslam::cuda::detail::preprocessing_lease::~preprocessing_lease()
{
    auto& p = pool();

    {
        std::scoped_lock lock { p.mutex };
        p.busy[_slot] = false;
    }

    p.released.notify_one();
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
