// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// A small OpenMP stand-in for the subset of the Kokkos API that the reference
// (PennyLane-Lightning-Kokkos, mounted read-only at /root/reference) uses. Kokkos itself
// is an un-vendored third-party dependency of the reference (fetched at build time,
// reference CMakeLists.txt:104-121, tag 3.7.00) and is absent from this image. All
// per-amplitude arithmetic lives in the reference's own headers; Kokkos supplies only
// the loop / reduce / scan skeleton, the View container, the complex type and an RNG
// pool. With this header on the include path the reference headers compile UNMODIFIED
// (g++ -std=c++17 -fopenmp) and serve as the executable oracle and the timed CPU
// baseline (oracle/Makefile -> oracle/_ref/libref_oracle.so).
//
// Semantics reproduced: RangePolicy parallel_for == one static-scheduled OpenMP loop
// (Kokkos-OpenMP behaviour); parallel_reduce == per-thread partial sums combined in
// thread order; parallel_scan == exclusive/inclusive scan through the functor's
// (k, update, final) protocol; TeamPolicy == league loop, team of one thread.
// NOT reproduced: Kokkos' exact reduction order and its XorShift64 pool stream
// (sampling parity is unpinned).
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <numeric>
#include <sstream>
#include <unordered_map>
#include <unordered_set>
#include <variant>
#include <memory>
#include <mutex>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]

namespace Kokkos {

// ----------------------------------------------------------------------------------
// complex
// ----------------------------------------------------------------------------------
template <class T> class complex {
    T re_{};
    T im_{};

  public:
    using value_type = T;
    constexpr complex() = default;
    constexpr complex(const T &re) : re_(re), im_(T(0)) {}
    constexpr complex(const T &re, const T &im) : re_(re), im_(im) {}
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>>
    constexpr complex(const complex<U> &o)
        : re_(static_cast<T>(o.real())), im_(static_cast<T>(o.imag())) {}
    constexpr complex(const std::complex<T> &o) : re_(o.real()), im_(o.imag()) {}
    operator std::complex<T>() const { return {re_, im_}; }

    template <class U> complex &operator=(const complex<U> &o) {
        re_ = static_cast<T>(o.real());
        im_ = static_cast<T>(o.imag());
        return *this;
    }
    complex &operator=(const T &re) {
        re_ = re;
        im_ = T(0);
        return *this;
    }

    constexpr T real() const { return re_; }
    constexpr T imag() const { return im_; }
    T &real() { return re_; }
    T &imag() { return im_; }
    void real(T v) { re_ = v; }
    void imag(T v) { im_ = v; }

    template <class U> complex &operator+=(const complex<U> &o) {
        re_ += o.real();
        im_ += o.imag();
        return *this;
    }
    complex &operator+=(const T &s) {
        re_ += s;
        return *this;
    }
    template <class U> complex &operator-=(const complex<U> &o) {
        re_ -= o.real();
        im_ -= o.imag();
        return *this;
    }
    complex &operator-=(const T &s) {
        re_ -= s;
        return *this;
    }
    template <class U> complex &operator*=(const complex<U> &o) {
        const T r = re_ * o.real() - im_ * o.imag();
        const T i = re_ * o.imag() + im_ * o.real();
        re_ = r;
        im_ = i;
        return *this;
    }
    complex &operator*=(const T &s) {
        re_ *= s;
        im_ *= s;
        return *this;
    }
    complex &operator/=(const T &s) {
        re_ /= s;
        im_ /= s;
        return *this;
    }
};
template <class T> complex(T, T) -> complex<T>;

template <class A, class B> using ct_t = std::common_type_t<A, B>;
template <class B> using arith_t = std::enable_if_t<std::is_arithmetic_v<B>, int>;

template <class A, class B>
inline complex<ct_t<A, B>> operator+(const complex<A> &x, const complex<B> &y) {
    using R = ct_t<A, B>;
    return {R(x.real()) + R(y.real()), R(x.imag()) + R(y.imag())};
}
template <class A, class B, arith_t<B> = 0>
inline complex<ct_t<A, B>> operator+(const complex<A> &x, const B &y) {
    using R = ct_t<A, B>;
    return {R(x.real()) + R(y), R(x.imag())};
}
template <class A, class B, arith_t<A> = 0>
inline complex<ct_t<A, B>> operator+(const A &x, const complex<B> &y) {
    using R = ct_t<A, B>;
    return {R(x) + R(y.real()), R(y.imag())};
}
template <class A, class B>
inline complex<ct_t<A, B>> operator-(const complex<A> &x, const complex<B> &y) {
    using R = ct_t<A, B>;
    return {R(x.real()) - R(y.real()), R(x.imag()) - R(y.imag())};
}
template <class A, class B, arith_t<B> = 0>
inline complex<ct_t<A, B>> operator-(const complex<A> &x, const B &y) {
    using R = ct_t<A, B>;
    return {R(x.real()) - R(y), R(x.imag())};
}
template <class A, class B, arith_t<A> = 0>
inline complex<ct_t<A, B>> operator-(const A &x, const complex<B> &y) {
    using R = ct_t<A, B>;
    return {R(x) - R(y.real()), -R(y.imag())};
}
template <class A> inline complex<A> operator-(const complex<A> &x) {
    return {-x.real(), -x.imag()};
}
template <class A> inline complex<A> operator+(const complex<A> &x) { return x; }
template <class A, class B>
inline complex<ct_t<A, B>> operator*(const complex<A> &x, const complex<B> &y) {
    using R = ct_t<A, B>;
    return {R(x.real()) * R(y.real()) - R(x.imag()) * R(y.imag()),
            R(x.real()) * R(y.imag()) + R(x.imag()) * R(y.real())};
}
template <class A, class B, arith_t<B> = 0>
inline complex<ct_t<A, B>> operator*(const complex<A> &x, const B &y) {
    using R = ct_t<A, B>;
    return {R(x.real()) * R(y), R(x.imag()) * R(y)};
}
template <class A, class B, arith_t<A> = 0>
inline complex<ct_t<A, B>> operator*(const A &x, const complex<B> &y) {
    using R = ct_t<A, B>;
    return {R(x) * R(y.real()), R(x) * R(y.imag())};
}
template <class A, class B, arith_t<B> = 0>
inline complex<ct_t<A, B>> operator/(const complex<A> &x, const B &y) {
    using R = ct_t<A, B>;
    return {R(x.real()) / R(y), R(x.imag()) / R(y)};
}
template <class A, class B>
inline complex<ct_t<A, B>> operator/(const complex<A> &x, const complex<B> &y) {
    using R = ct_t<A, B>;
    const R d = R(y.real()) * R(y.real()) + R(y.imag()) * R(y.imag());
    return {(R(x.real()) * R(y.real()) + R(x.imag()) * R(y.imag())) / d,
            (R(x.imag()) * R(y.real()) - R(x.real()) * R(y.imag())) / d};
}
template <class A, class B>
inline bool operator==(const complex<A> &x, const complex<B> &y) {
    return x.real() == y.real() && x.imag() == y.imag();
}
template <class A, class B>
inline bool operator!=(const complex<A> &x, const complex<B> &y) {
    return !(x == y);
}
template <class T> inline T real(const complex<T> &x) { return x.real(); }
template <class T> inline T imag(const complex<T> &x) { return x.imag(); }
template <class T> inline complex<T> conj(const complex<T> &x) {
    return {x.real(), -x.imag()};
}
template <class T> inline T abs(const complex<T> &x) {
    return std::hypot(x.real(), x.imag());
}
template <class T> inline complex<T> exp(const complex<T> &x) {
    const T e = std::exp(x.real());
    return {e * std::cos(x.imag()), e * std::sin(x.imag())};
}
template <class T> inline std::ostream &operator<<(std::ostream &os, const complex<T> &x) {
    return os << std::complex<T>(x);
}
template <class T, class = arith_t<T>> inline T cos(T x) { return std::cos(x); }
template <class T, class = arith_t<T>> inline T sin(T x) { return std::sin(x); }
template <class T, class = arith_t<T>> inline T exp(T x) { return std::exp(x); }
template <class T, class = arith_t<T>> inline T sqrt(T x) { return std::sqrt(x); }

namespace Experimental {
template <class T> inline void swap(T &a, T &b) {
    T t = a;
    a = b;
    b = t;
}
} // namespace Experimental

namespace Impl {
inline int bit_count(unsigned long long x) { return __builtin_popcountll(x); }
} // namespace Impl

template <class T, class V> inline void atomic_add(T *dst, const V &v) {
    const T val = static_cast<T>(v);
#pragma omp atomic
    *dst += val;
}

// ----------------------------------------------------------------------------------
// spaces, traits, View
// ----------------------------------------------------------------------------------
struct HostSpace {};
enum MemoryTraitsFlags : unsigned { Unmanaged = 0x01 };
template <unsigned F> struct MemoryTraits {};

// Bump allocator handed out by team_scratch(0).
struct ScratchMemorySpace {
    char *base = nullptr;
    mutable std::size_t offset = 0;
    std::size_t capacity = 0;
    void *get(std::size_t bytes) const {
        const std::size_t aligned = (offset + 15) & ~std::size_t(15);
        assert(aligned + bytes <= capacity);
        offset = aligned + bytes;
        return base + aligned;
    }
};

struct OpenMP {
    using scratch_memory_space = ScratchMemorySpace;
    using execution_space = OpenMP;
    using memory_space = HostSpace;
};
using DefaultExecutionSpace = OpenMP;
using DefaultHostExecutionSpace = OpenMP;

template <class DataType, class... Props> class View;

template <class T, class... Props> class View<T *, Props...> {
    using NC = std::remove_const_t<T>;
    std::shared_ptr<NC[]> owner_;
    T *ptr_ = nullptr;
    std::size_t n_ = 0;

  public:
    using value_type = T;
    View() = default;
    View(const std::string & /*label*/, std::size_t n) : n_(n) {
        if (n) {
            // zero-initialised, like Kokkos
            owner_ = std::shared_ptr<NC[]>(
                static_cast<NC *>(std::calloc(n, sizeof(NC))),
                [](NC *p) { std::free(p); });
            ptr_ = owner_.get();
        }
    }
    View(T *ptr, std::size_t n) : ptr_(ptr), n_(n) {}
    View(const ScratchMemorySpace &s, std::size_t n)
        : ptr_(static_cast<T *>(s.get(n * sizeof(NC)))), n_(n) {}
    static std::size_t shmem_size(std::size_t n) { return n * sizeof(NC) + 16; }

    T &operator()(std::size_t i) const { return ptr_[i]; }
    T &operator[](std::size_t i) const { return ptr_[i]; }
    std::size_t size() const { return n_; }
    std::size_t extent(int) const { return n_; }
    T *data() const { return ptr_; }
};

template <class DV, class SV> inline void deep_copy(const DV &dst, const SV &src) {
    using D = std::remove_const_t<typename DV::value_type>;
    using S = std::remove_const_t<typename SV::value_type>;
    static_assert(std::is_same_v<D, S>, "deep_copy between different element types");
    const std::size_t n = std::min(dst.size(), src.size());
    if (n && static_cast<const void *>(dst.data()) != static_cast<const void *>(src.data()))
        std::memcpy(const_cast<D *>(dst.data()), src.data(), n * sizeof(D));
}

template <class V> inline V create_mirror_view_and_copy(HostSpace, const V &v) { return v; }

// ----------------------------------------------------------------------------------
// policies
// ----------------------------------------------------------------------------------
template <class... P> struct RangePolicy {
    std::size_t begin_, end_;
    RangePolicy(std::size_t b, std::size_t e) : begin_(b), end_(e) {}
    std::size_t begin() const { return begin_; }
    std::size_t end() const { return end_; }
};

struct AUTO_t {};
static constexpr AUTO_t AUTO{};
struct PerTeamValue {
    std::size_t bytes;
};
inline PerTeamValue PerTeam(std::size_t b) { return {b}; }

struct TeamMember {
    std::size_t league_rank_ = 0;
    ScratchMemorySpace scratch_;
    std::size_t league_rank() const { return league_rank_; }
    int team_rank() const { return 0; }
    int team_size() const { return 1; }
    void team_barrier() const {}
    const ScratchMemorySpace &team_scratch(int) const { return scratch_; }
};

template <class... P> struct TeamPolicy {
    using member_type = TeamMember;
    std::size_t league_ = 0;
    std::size_t scratch_bytes_ = 0;
    TeamPolicy(std::size_t league, AUTO_t, std::size_t /*vec*/ = 1) : league_(league) {}
    TeamPolicy(std::size_t league, int, std::size_t = 1) : league_(league) {}
    TeamPolicy &set_scratch_size(int, PerTeamValue v) {
        scratch_bytes_ = v.bytes;
        return *this;
    }
};

struct NestedRange {
    std::size_t n;
};
inline NestedRange ThreadVectorRange(const TeamMember &, std::size_t n) { return {n}; }
inline NestedRange TeamThreadRange(const TeamMember &, std::size_t n) { return {n}; }
inline NestedRange TeamVectorRange(const TeamMember &, std::size_t n) { return {n}; }

enum class Iterate { Default, Left, Right };
template <unsigned N, Iterate A = Iterate::Default, Iterate B = Iterate::Default> struct Rank {};
template <class R> struct MDRangePolicy {
    std::array<long, 2> lo, hi;
    MDRangePolicy(std::array<long, 2> l, std::array<long, 2> h) : lo(l), hi(h) {}
};

// ----------------------------------------------------------------------------------
// parallel dispatch
// ----------------------------------------------------------------------------------
template <class F> inline void parallel_for(const NestedRange &r, const F &f) {
    for (std::size_t i = 0; i < r.n; i++)
        f(i);
}

template <class... P, class F>
inline void parallel_for(const RangePolicy<P...> &p, const F &f) {
    const long long b = static_cast<long long>(p.begin()), e = static_cast<long long>(p.end());
#pragma omp parallel for schedule(static)
    for (long long i = b; i < e; i++)
        f(static_cast<std::size_t>(i));
}
template <class I, class F, std::enable_if_t<std::is_integral_v<I>, int> = 0>
inline void parallel_for(const I n, const F &f) {
    parallel_for(RangePolicy<>(0, static_cast<std::size_t>(n)), f);
}
template <class... P, class F>
inline void parallel_for(const std::string &, const RangePolicy<P...> &p, const F &f) {
    parallel_for(p, f);
}
template <class... P, class F>
inline void parallel_for(const TeamPolicy<P...> &p, const F &f) {
    const long long n = static_cast<long long>(p.league_);
#pragma omp parallel
    {
        std::vector<char> buf(p.scratch_bytes_ + 64);
#pragma omp for schedule(static)
        for (long long i = 0; i < n; i++) {
            TeamMember m;
            m.league_rank_ = static_cast<std::size_t>(i);
            m.scratch_.base = buf.data();
            m.scratch_.offset = 0;
            m.scratch_.capacity = buf.size();
            f(m);
        }
    }
}
template <class... P, class F>
inline void parallel_for(const std::string &, const TeamPolicy<P...> &p, const F &f) {
    parallel_for(p, f);
}
template <class R, class F> inline void parallel_for(const MDRangePolicy<R> &p, const F &f) {
    // first index parallel, second serial: keeps the reference's atomic_add targets
    // (indexed by the first index) race-free and the summation order deterministic.
#pragma omp parallel for schedule(static)
    for (long i = p.lo[0]; i < p.hi[0]; i++)
        for (long j = p.lo[1]; j < p.hi[1]; j++)
            f(static_cast<std::size_t>(i), static_cast<std::size_t>(j));
}
template <class R, class F>
inline void parallel_for(const std::string &, const MDRangePolicy<R> &p, const F &f) {
    parallel_for(p, f);
}

template <class F, class T>
inline void parallel_reduce(const std::size_t n, const F &f, T &result) {
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    std::vector<T> partial(static_cast<std::size_t>(nt) * 8, T(0)); // padded
#pragma omp parallel num_threads(nt)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        T acc = T(0);
#pragma omp for schedule(static) nowait
        for (long long i = 0; i < static_cast<long long>(n); i++)
            f(static_cast<std::size_t>(i), acc);
        partial[static_cast<std::size_t>(t) * 8] = acc;
    }
    T total = T(0);
    for (int t = 0; t < nt; t++)
        total += partial[static_cast<std::size_t>(t) * 8];
    result = total;
}
template <class... P, class F, class T>
inline void parallel_reduce(const RangePolicy<P...> &p, const F &f, T &result) {
    assert(p.begin() == 0);
    parallel_reduce(p.end(), f, result);
}

template <class... P, class F> inline void parallel_scan(const RangePolicy<P...> &p, const F &f) {
    // value type is deduced from the functor's second argument
    using first_arg = std::size_t;
    (void)sizeof(first_arg);
    scan_impl(p, f, &F::operator());
}
template <class... P, class F, class C, class K, class U>
inline void scan_impl(const RangePolicy<P...> &p, const F &f,
                      void (C::*)(K, U &, const bool) const) {
    U update = U(0);
    for (std::size_t k = p.begin(); k < p.end(); k++)
        f(k, update, true);
}

inline void fence() {}

// ----------------------------------------------------------------------------------
// runtime lifecycle
// ----------------------------------------------------------------------------------
#define KSHIM_SETTING(name, type)                                                           \
  private:                                                                                   \
    type name##_{};                                                                          \
    bool has_##name##_ = false;                                                              \
                                                                                             \
  public:                                                                                    \
    InitializationSettings &set_##name(const type &v) {                                      \
        name##_ = v;                                                                         \
        has_##name##_ = true;                                                                \
        return *this;                                                                        \
    }                                                                                        \
    bool has_##name() const { return has_##name##_; }                                        \
    const type &get_##name() const { return name##_; }

class InitializationSettings {
    KSHIM_SETTING(num_threads, int)
    KSHIM_SETTING(device_id, int)
    KSHIM_SETTING(map_device_id_by, std::string)
    KSHIM_SETTING(disable_warnings, bool)
    KSHIM_SETTING(print_configuration, bool)
    KSHIM_SETTING(tune_internals, bool)
    KSHIM_SETTING(tools_libs, std::string)
    KSHIM_SETTING(tools_help, bool)
    KSHIM_SETTING(tools_args, std::string)
};
#undef KSHIM_SETTING

namespace detail {
inline int &state() {
    static int s = 0; // 0 = fresh, 1 = initialised, 2 = finalised
    return s;
}
} // namespace detail
inline void initialize(const InitializationSettings &s = {}) {
    detail::state() = 1;
#ifdef _OPENMP
    if (s.has_num_threads() && s.get_num_threads() > 0)
        omp_set_num_threads(s.get_num_threads());
#else
    (void)s;
#endif
}
inline void finalize() { detail::state() = 2; }
inline bool is_initialized() { return detail::state() == 1; }
inline bool is_finalized() { return detail::state() == 2; }
inline void print_configuration(std::ostream &os, bool = false) {
    os << "Kokkos stand-in (oracle/kokkos_shim): OpenMP host backend\n";
}

} // namespace Kokkos
