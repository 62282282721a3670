// TEST-ONLY: g++ 13 has no std::ranges::to (C++23), which the reference's map.h uses; a minimal stand-in so the
// syntax check can parse that header.  Force-included by tests/test_adapter_syntax.py only.
#pragma once
#include <ranges>
#include <vector>
#if !defined(__cpp_lib_ranges_to_container)
namespace std::ranges
{
    template <template <class...> class C> struct zs_to_closure { };
    template <template <class...> class C> constexpr auto to() { return zs_to_closure<C> { }; }
    template <class R, template <class...> class C> auto operator|(R&& r, zs_to_closure<C>)
    {
        C<range_value_t<R>> out;
        for (auto&& x : r) out.push_back(x);
        return out;
    }
}
#endif
