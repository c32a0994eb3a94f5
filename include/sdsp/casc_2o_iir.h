// sdsp/casc_2o_iir.h -- drop-in replacement for the reference's header of the same name, backed by
// libsdsp_b200.so (CUDA, sm_100a).  Host-only C++17.
//
//     sdsp::casc_2o_iir<4> lp;            // 4 biquads = 8th-order Butterworth
//     lp.set_lp_coeff(f0, fs);            // designers run on the host (scalar, once per filter)
//     lp.process(buf.begin(), buf.end()); // the recurrence runs on the GPU; history stays in the object
//
// Class names, member functions, default arguments and static_asserts follow reference
// include/sdsp/casc_2o_iir.h (cited per item).  The object is copyable and carries its history
// between process() calls exactly like the reference object does (m_mem); the 3-slot ring + m_pos
// of the reference (casc_2o_iir.h:11,15) is replaced by "two most recent values per row", which is
// the same information.  For many channels use sdsp::iir_bank below (an addition): coefficients and
// history then stay resident on the device.
#pragma once
#include <array>
#include <cstddef>
#include <iterator>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../sdsp_b200.h"
#include "filter_type.h"

namespace sdsp
{
namespace detail
{
    inline void iir_check(int status, const char *what)
    {
        if (status != SDSP_B200_OK)
            throw std::runtime_error(std::string(what) + ": " + sdsp_b200_last_error());
    }

    template <typename V>
    constexpr int iir_precision()
    {
        static_assert(std::is_same_v<V, float> || std::is_same_v<V, double>, "samples must be float or double");
        return std::is_same_v<V, float> ? SDSP_B200_F32 : SDSP_B200_F64;
    }

    // state + coefficients shared by the four classes
    template <size_t m_t>
    struct iir_core {
        double m_gain{ 1.0 };
        std::array<std::array<double, 2>, m_t + 1> m_mem{};   // row r: { x[n-1], x[n-2] }
        std::array<std::array<double, 3>, m_t> m_b_coeff{};
        std::array<std::array<double, 3>, m_t> m_a_coeff{};

        template <typename iter_t>
        void run(int numerator, iter_t begin, iter_t end)
        {
            using V = typename std::iterator_traits<iter_t>::value_type;
            const auto count = std::distance(begin, end);
            if (count <= 0)
                return;
            const size_t n = static_cast<size_t>(count);
            V *first = &*begin;
            // vector / array / pointer iterators are contiguous; anything else is staged
            const bool contiguous = (&*(begin + (count - 1)) == first + (count - 1));
            std::vector<V> staged;
            if (!contiguous) {
                staged.assign(begin, end);
                first = staged.data();
            }
            iir_check(sdsp_b200_iir_process_once(static_cast<int>(m_t), numerator, iir_precision<V>(), m_gain, &m_b_coeff[0][0],
                                                 &m_a_coeff[0][0], &m_mem[0][0], first, n, 0),
                      "sdsp_b200_iir_process_once");
            if (!contiguous)
                std::copy(staged.begin(), staged.end(), begin);
        }
    };
} // namespace detail

// ---- casc_2o_iir<m_t>: reference casc_2o_iir.h:8-215 -----------------------------------------
template <size_t m_t>
class casc_2o_iir {
private:
    detail::iir_core<m_t> m_core;
    filter_type m_f_type{ filter_type::none };

public:
    casc_2o_iir()
    {
        static_assert(m_t % 2 == 0, "M must be even!"); // casc_2o_iir.h:25
        static_assert(m_t <= SDSP_B200_IIR_MAX_SECTIONS_ONCE, "libsdsp_b200 chains at most 64 sections per filter object");
    }

    // coefficients and type, never the history: casc_2o_iir.h:28-34
    void copy_coeff_from(const casc_2o_iir<m_t> &other_filter)
    {
        m_core.m_gain = other_filter.m_core.m_gain;
        m_core.m_b_coeff = other_filter.m_core.m_b_coeff;
        m_core.m_a_coeff = other_filter.m_core.m_a_coeff;
        m_f_type = other_filter.m_f_type;
    }

    // casc_2o_iir.h:36-80
    template <typename iter_t>
    void process(iter_t begin, iter_t end)
    {
        m_core.run(SDSP_B200_NUM_GENERIC, begin, end);
    }

    // casc_2o_iir.h:82-138
    void set_bp_coeff(double f0, double fs, double q, double gain_in = 1.0)
    {
        m_f_type = filter_type::band_pass;
        detail::iir_check(sdsp_b200_iir_design_bp(static_cast<int>(m_t), f0, fs, q, gain_in, &m_core.m_gain, &m_core.m_b_coeff[0][0],
                                                  &m_core.m_a_coeff[0][0]),
                          "sdsp_b200_iir_design_bp");
    }
    // casc_2o_iir.h:140-166
    void set_hp_coeff(double f0, double fs, double gain_in = 1.0)
    {
        m_f_type = filter_type::high_pass;
        detail::iir_check(sdsp_b200_iir_design_hp(static_cast<int>(m_t), f0, fs, gain_in, &m_core.m_gain, &m_core.m_b_coeff[0][0],
                                                  &m_core.m_a_coeff[0][0]),
                          "sdsp_b200_iir_design_hp");
    }
    // casc_2o_iir.h:168-194
    void set_lp_coeff(double f0, double fs, double gain_in = 1.0)
    {
        m_f_type = filter_type::low_pass;
        detail::iir_check(sdsp_b200_iir_design_lp(static_cast<int>(m_t), f0, fs, gain_in, &m_core.m_gain, &m_core.m_b_coeff[0][0],
                                                  &m_core.m_a_coeff[0][0]),
                          "sdsp_b200_iir_design_lp");
    }
    // casc_2o_iir.h:196-214
    void preload_filter(double value)
    {
        detail::iir_check(sdsp_b200_iir_preload_state(static_cast<int>(m_t), static_cast<int>(m_f_type), m_core.m_gain,
                                                      &m_core.m_b_coeff[0][0], &m_core.m_a_coeff[0][0], value, &m_core.m_mem[0][0]),
                          "sdsp_b200_iir_preload_state");
    }

    // additions: what a channel bank needs to adopt this object's design
    double gain() const { return m_core.m_gain; }
    const double *b_coeff() const { return &m_core.m_b_coeff[0][0]; }
    const double *a_coeff() const { return &m_core.m_a_coeff[0][0]; }
    filter_type type() const { return m_f_type; }
};

// ---- fixed-numerator classes: reference casc_2o_iir.h:217-468 --------------------------------
// numerators {1,2,1} / {1,-2,1} / {1,0,-1} are hard-wired in the kernel; no b coefficients, no preload.
// copy_coeff_from works here (the reference's version, :274-278 / :332-336 / :390-394, names members
// that do not exist and cannot be instantiated).
#define SDSP_B200_FIXED_IIR(CLASS, NUMERATOR, SETTER_DECL, DESIGN_CALL)                                         \
    template <size_t m_t>                                                                                      \
    class CLASS {                                                                                              \
    private:                                                                                                   \
        detail::iir_core<m_t> m_core;                                                                          \
                                                                                                               \
    public:                                                                                                    \
        CLASS()                                                                                                \
        {                                                                                                      \
            static_assert(m_t % 2 == 0, "M must be even!");                                                    \
            static_assert(m_t <= SDSP_B200_IIR_MAX_SECTIONS_ONCE, "libsdsp_b200 chains at most 64 sections per filter object");                             \
        }                                                                                                      \
        void copy_coeff_from(const CLASS<m_t> &other_filter)                                                   \
        {                                                                                                      \
            m_core.m_gain = other_filter.m_core.m_gain;                                                        \
            m_core.m_a_coeff = other_filter.m_core.m_a_coeff;                                                  \
        }                                                                                                      \
        template <typename iter_t>                                                                             \
        void process(iter_t begin, iter_t end)                                                                 \
        {                                                                                                      \
            m_core.run(NUMERATOR, begin, end);                                                                 \
        }                                                                                                      \
        void SETTER_DECL                                                                                       \
        {                                                                                                      \
            detail::iir_check(DESIGN_CALL, #CLASS " design");                                                  \
        }                                                                                                      \
    };

SDSP_B200_FIXED_IIR(casc_2o_iir_lp, SDSP_B200_NUM_LP, set_lp_coeff(double f0, double fs, double gain_in = 1.0),
                    sdsp_b200_iir_design_lp(static_cast<int>(m_t), f0, fs, gain_in, &m_core.m_gain, &m_core.m_b_coeff[0][0],
                                            &m_core.m_a_coeff[0][0]))
SDSP_B200_FIXED_IIR(casc_2o_iir_hp, SDSP_B200_NUM_HP, set_hp_coeff(double f0, double fs, double gain_in = 1.0),
                    sdsp_b200_iir_design_hp(static_cast<int>(m_t), f0, fs, gain_in, &m_core.m_gain, &m_core.m_b_coeff[0][0],
                                            &m_core.m_a_coeff[0][0]))
SDSP_B200_FIXED_IIR(casc_2o_iir_bp, SDSP_B200_NUM_BP, set_bp_coeff(double f0, double fs, double q, double gain_in = 1.0),
                    sdsp_b200_iir_design_bp(static_cast<int>(m_t), f0, fs, q, gain_in, &m_core.m_gain, &m_core.m_b_coeff[0][0],
                                            &m_core.m_a_coeff[0][0]))
#undef SDSP_B200_FIXED_IIR

// ---- addition: a bank of channels resident on the device -------------------------------------
// The batched counterpart of "one casc_2o_iir<m_t> object per channel".  data is planar
// [channel][sample]; V is float or double.
template <size_t m_t, typename V = float>
class iir_bank {
    static_assert(m_t >= 1 && m_t <= 8, "a device-resident bank holds 1..8 sections per channel");
    sdsp_b200_iir_bank m_bank{ nullptr };
    size_t m_channels{ 0 };

public:
    explicit iir_bank(size_t n_channels, int numerator = SDSP_B200_NUM_GENERIC, int device = 0) : m_channels(n_channels)
    {
        detail::iir_check(sdsp_b200_iir_bank_create(&m_bank, static_cast<int>(m_t), n_channels, detail::iir_precision<V>(), numerator,
                                                    device),
                          "sdsp_b200_iir_bank_create");
    }
    ~iir_bank()
    {
        sdsp_b200_iir_bank_destroy(m_bank);
    }
    iir_bank(const iir_bank &) = delete;
    iir_bank &operator=(const iir_bank &) = delete;

    size_t channels() const { return m_channels; }
    // channel `c` adopts the design of a host filter object (history is not copied)
    void copy_coeff_from(size_t c, const casc_2o_iir<m_t> &f)
    {
        const double g = f.gain();
        detail::iir_check(sdsp_b200_iir_bank_set_coeffs(m_bank, c, 1, &g, f.b_coeff(), f.a_coeff()), "sdsp_b200_iir_bank_set_coeffs");
    }
    void set_coeffs(size_t first, size_t count, const double *gain, const double *b, const double *a)
    {
        detail::iir_check(sdsp_b200_iir_bank_set_coeffs(m_bank, first, count, gain, b, a), "sdsp_b200_iir_bank_set_coeffs");
    }
    void reset()
    {
        detail::iir_check(sdsp_b200_iir_bank_reset_state(m_bank), "sdsp_b200_iir_bank_reset_state");
    }
    // host data, synchronous
    void process(V *data, size_t n_samples, size_t channel_stride, int path = SDSP_B200_IIR_AUTO)
    {
        detail::iir_check(sdsp_b200_iir_bank_process(m_bank, data, n_samples, channel_stride, SDSP_B200_PTR_HOST, path, nullptr),
                          "sdsp_b200_iir_bank_process");
    }
    // device data, asynchronous on `stream`
    void process_device(V *data, size_t n_samples, size_t channel_stride, void *stream = nullptr, int path = SDSP_B200_IIR_AUTO)
    {
        detail::iir_check(sdsp_b200_iir_bank_process(m_bank, data, n_samples, channel_stride, SDSP_B200_PTR_DEVICE, path, stream),
                          "sdsp_b200_iir_bank_process");
    }
    sdsp_b200_iir_bank handle() const { return m_bank; }
};
} // namespace sdsp
