// shim_main.cpp -- test runner for the Catch2 stand-in (TEST INFRASTRUCTURE).
// Accepts and ignores Catch2's command line (the reference CI passes --order rand --warn NoAssertions).
#include "catch2/catch_all.hpp"

namespace catch_shim
{
std::vector<test_case> &registry()
{
    static std::vector<test_case> r;
    return r;
}
run_state &state()
{
    static run_state s;
    return s;
}
} // namespace catch_shim

int main(int, char **)
{
    using namespace catch_shim;
    int failed_cases = 0, total_assertions = 0;
    for (const test_case &tc : registry()) {
        run_state &s = state();
        s = run_state{};
        bool failed = false;
        int passes = 0;
        do {
            s.seen = 0;
            s.current_section = "<none>";
            try {
                tc.fn();
            } catch (const assertion_failed &e) {
                std::printf("  FAILED  %s\n", e.what());
                failed = true;
            } catch (const std::exception &e) {
                std::printf("  EXCEPTION in \"%s\" [section: %s]: %s\n", tc.name, s.current_section.c_str(), e.what());
                failed = true;
            }
            s.total = s.seen;
            s.pass++;
            passes++;
        } while (s.pass < s.total);
        std::printf("%s  \"%s\"  (%d section pass%s, %d assertions)\n", failed ? "[FAIL]" : "[ ok ]", tc.name, passes,
                    passes == 1 ? "" : "es", s.assertions);
        total_assertions += s.assertions;
        failed_cases += failed ? 1 : 0;
    }
    std::printf("%zu test cases, %d failed, %d assertions\n", registry().size(), failed_cases, total_assertions);
    return failed_cases ? 1 : 0;
}
