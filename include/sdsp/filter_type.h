// sdsp/filter_type.h -- drop-in for the reference's include/sdsp/filter_type.h (enum at :6).
// The integer values are part of the contract: they are the first field of the golden CSV fixtures
// (reference test/testIIR.cpp:18-19) and equal SDSP_B200_LOW_PASS.. in sdsp_b200.h.
#pragma once

namespace sdsp
{
enum class filter_type { none = 0, low_pass = 1, high_pass = 2, band_pass = 3 };
}
