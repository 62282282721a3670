// TEST-ONLY stand-in, see opencv2/core.hpp
#pragma once
namespace spdlog
{
    template <typename... A> void warn(A&&...) { }
    template <typename... A> void error(A&&...) { }
    template <typename... A> void info(A&&...) { }
    template <typename... A> void debug(A&&...) { }
}
