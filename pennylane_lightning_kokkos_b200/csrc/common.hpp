// b2sv: shared host-side helpers (errors, CUDA checks, complex alias).
#pragma once
#include <complex>
#include <cstdint>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace b2sv {

using cplx = std::complex<double>;

// Mirrors the reference's LightningException text (reference util/Error.hpp:115-142):
// "[file][Line:n][Method:f]: Error in PennyLane Lightning: msg"
struct Error : public std::runtime_error {
    explicit Error(const std::string &m) : std::runtime_error(m) {}
};

[[noreturn]] inline void abort_with(const std::string &msg, const char *file, int line,
                                    const char *func) {
    std::ostringstream os;
    os << "[" << file << "][Line:" << line << "][Method:" << func
       << "]: Error in PennyLane Lightning: " << msg;
    throw Error(os.str());
}

#define B2_ABORT(msg) ::b2sv::abort_with((msg), __FILE__, __LINE__, __func__)
#define B2_ABORT_IF(cond, msg)                                                                  \
    do {                                                                                        \
        if (cond)                                                                               \
            B2_ABORT(msg);                                                                      \
    } while (0)
#define B2_ASSERT(cond) B2_ABORT_IF(!(cond), "Assertion failed: " #cond)

#define CUDA_CHECK(call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            B2_ABORT(std::string("CUDA error: ") + cudaGetErrorString(e__) + " in " #call);     \
    } while (0)

inline uint64_t bit(int p) { return uint64_t(1) << p; }

// Per-device one-time setup (function attributes are per device, a process may hold states on
// several): true the first time it is called for the current device with this `mask`.
inline bool first_use_on_device(uint64_t &mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t b = uint64_t(1) << (dev & 63);
    if (mask & b)
        return false;
    mask |= b;
    return true;
}
inline int sm_count_current_device() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int &n = cached[dev & 63];
    if (!n)
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

} // namespace b2sv
