// catch2/catch_all.hpp -- a minimal stand-in for Catch2 v3 (TEST INFRASTRUCTURE).
//
// Catch2 is not installed in this image and cannot be fetched.  This shim provides just what the
// reference's test/testFFT.cpp and test/testIIR.cpp use, so that they compile UNMODIFIED against the
// drop-in headers: TEST_CASE, SECTION (the test body is re-entered once per leaf section, as Catch2
// does), REQUIRE, BENCHMARK and SUCCEED.
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

namespace catch_shim
{
struct test_case {
    const char *name;
    void (*fn)();
};
std::vector<test_case> &registry();
struct registrar {
    registrar(const char *name, void (*fn)())
    {
        registry().push_back({ name, fn });
    }
};

// flat SECTION re-entry: in pass p of a test case only the p-th SECTION met in that pass runs
struct run_state {
    int pass = 0;       // which section this pass executes
    int seen = 0;       // sections met so far in this pass
    int total = 0;      // sections met in the previous complete pass
    int assertions = 0;
    int failures = 0;
    std::string current_section;
};
run_state &state();

struct section_guard {
    bool active;
    explicit section_guard(const char *name)
    {
        run_state &s = state();
        active = (s.seen == s.pass);
        if (active)
            s.current_section = name;
        s.seen++;
    }
    explicit operator bool() const
    {
        return active;
    }
};

struct assertion_failed : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void require(bool ok, const char *expr, const char *file, int line)
{
    run_state &s = state();
    s.assertions++;
    if (!ok) {
        s.failures++;
        char buf[1024];
        std::snprintf(buf, sizeof(buf), "%s:%d: REQUIRE( %s ) failed [section: %s]", file, line, expr, s.current_section.c_str());
        throw assertion_failed(buf);
    }
}

struct bench_runner {
    std::string name;
    template <typename F>
    bench_runner &operator=(F &&f)
    {
        using clock = std::chrono::steady_clock;
        f(); // warm-up (first GPU call creates plans / banks)
        int iters = 0;
        const auto t0 = clock::now();
        double elapsed = 0;
        do {
            auto keep = f();
            (void)keep;
            iters++;
            elapsed = std::chrono::duration<double>(clock::now() - t0).count();
        } while (elapsed < 0.2 && iters < 1000);
        std::printf("    benchmark %-50s %10.2f us/iter (%d iters)\n", name.c_str(), elapsed / iters * 1e6, iters);
        return *this;
    }
};
} // namespace catch_shim

#define CATCH_SHIM_CAT2(a, b) a##b
#define CATCH_SHIM_CAT(a, b) CATCH_SHIM_CAT2(a, b)

#define TEST_CASE(...) CATCH_SHIM_TEST_CASE(CATCH_SHIM_CAT(catch_shim_test_, __COUNTER__), __VA_ARGS__)
#define CATCH_SHIM_FIRST(a, ...) a
#define CATCH_SHIM_TEST_CASE(fn, ...)                                                        \
    static void fn();                                                                        \
    static ::catch_shim::registrar CATCH_SHIM_CAT(fn, _reg)(CATCH_SHIM_FIRST(__VA_ARGS__, 0), &fn); \
    static void fn()

#define SECTION(name) if (::catch_shim::section_guard CATCH_SHIM_CAT(catch_shim_sec_, __LINE__){ name })
#define REQUIRE(...) ::catch_shim::require(static_cast<bool>(__VA_ARGS__), #__VA_ARGS__, __FILE__, __LINE__)
#define SUCCEED(...) (::catch_shim::state().assertions++)
#define BENCHMARK(name) ::catch_shim::bench_runner{ name } = [&]()
