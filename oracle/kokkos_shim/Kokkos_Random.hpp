// TEST INFRASTRUCTURE ONLY -- stand-in for Kokkos_Random.hpp (see Kokkos_Core.hpp here).
// The reference's sampler (reference MeasuresFunctors.hpp:68-120, MeasuresKokkos.hpp:551)
// draws uniform doubles from Kokkos::Random_XorShift64_Pool. The real pool's stream is
// not reproduced (Kokkos is absent); only "uniform in [a,b)" is, so sampling parity is
// distributional, exactly as in the reference's own test
// (reference src/tests/Test_StateVectorKokkos_Param.cpp:1325-1394).
#pragma once
#include "Kokkos_Core.hpp"
#include <atomic>

namespace Kokkos {

struct XorShift64State {
    std::uint64_t s;
    std::uint64_t next() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return s;
    }
    double drand(double a, double b) {
        return a + (b - a) * (static_cast<double>(next() >> 11) * 0x1.0p-53);
    }
    float frand(float a, float b) { return static_cast<float>(drand(a, b)); }
};

template <class ExecSpace = DefaultExecutionSpace> class Random_XorShift64_Pool {
    std::uint64_t seed_;
    std::shared_ptr<std::atomic<std::uint64_t>> counter_;

  public:
    using generator_type = XorShift64State;
    explicit Random_XorShift64_Pool(std::uint64_t seed = 1)
        : seed_(seed), counter_(std::make_shared<std::atomic<std::uint64_t>>(0)) {}
    XorShift64State get_state() const {
        // splitmix64 of (seed, draw counter): one independent stream per acquisition
        std::uint64_t z = seed_ + 0x9E3779B97F4A7C15ull * (1 + counter_->fetch_add(1));
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        XorShift64State g{z ? z : 0x2545F4914F6CDD1Dull};
        g.next();
        return g;
    }
    void free_state(const XorShift64State &) const {}
};

} // namespace Kokkos
