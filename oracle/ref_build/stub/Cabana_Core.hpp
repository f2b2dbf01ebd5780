// TEST INFRASTRUCTURE (oracle/_ref build only).
//
// Stand-in for the slice of Cabana / Kokkos that the reference's Library/Storage/*.hpp touches, so
// that those headers (and through them FEM/DATA_TYPE.h, FEM/IPC.h, Grid/SPATIAL_HASH.h,
// FEM/FRICTION.h) compile from where they lie: an AoSoA with Cabana's memory layout (bins of
// VectorLength elements, each member stored as T[D0][VectorLength] / T[VectorLength]), slices with
// (i) / (i, j) access, deep_copy, and Kokkos::parallel_for over a RangePolicy mapped to an OpenMP
// loop (Kokkos::OpenMP is what the reference builds with, CMakeLists.txt STORAGE_ENABLED_OPENMP).
// Storage plumbing only -- no algorithm of the contact path lives here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define KOKKOS_LAMBDA [=]
#define KOKKOS_INLINE_FUNCTION inline

namespace Kokkos {
struct HostSpace {};
struct OpenMP {};
struct Serial {};
template <class E, class M> struct Device {};
template <class ExSpace> struct RangePolicy {
    long b, e;
    RangePolicy(long b_, long e_) : b(b_), e(e_) {}
};
template <class ExSpace, class F> void parallel_for(const RangePolicy<ExSpace>& p, const F& f, const std::string& = "")
{
    if (std::is_same<ExSpace, OpenMP>::value) {
#pragma omp parallel for schedule(static)
        for (long i = p.b; i < p.e; ++i) f((int)i);
    }
    else
        for (long i = p.b; i < p.e; ++i) f((int)i);
}
inline void fence() {}
} // namespace Kokkos

namespace Cabana {

template <class... Ts> struct MemberTypes {
    static constexpr std::size_t size = sizeof...(Ts);
};
template <class MT> struct Tuple {};

namespace detail {
template <class M> struct MemberInfo { // scalar member
    using value_type = M;
    static constexpr std::size_t extent = 1;
};
template <class M, std::size_t D0> struct MemberInfo<M[D0]> {
    using value_type = M;
    static constexpr std::size_t extent = D0;
};
template <std::size_t I, class... Ts> struct Nth;
template <std::size_t I, class T, class... Ts> struct Nth<I, T, Ts...> : Nth<I - 1, Ts...> {};
template <class T, class... Ts> struct Nth<0, T, Ts...> { using type = T; };
template <std::size_t VLEN, class... Ts> constexpr std::size_t bin_bytes()
{
    return (0 + ... + (sizeof(typename MemberInfo<Ts>::value_type) * MemberInfo<Ts>::extent * VLEN));
}
template <std::size_t VLEN, std::size_t I, class... Ts> struct OffsetOf;
template <std::size_t VLEN, std::size_t I, class T, class... Ts> struct OffsetOf<VLEN, I, T, Ts...> {
    static constexpr std::size_t value = sizeof(typename MemberInfo<T>::value_type) * MemberInfo<T>::extent * VLEN + OffsetOf<VLEN, I - 1, Ts...>::value;
};
template <std::size_t VLEN, class T, class... Ts> struct OffsetOf<VLEN, 0, T, Ts...> { static constexpr std::size_t value = 0; };
template <std::size_t VLEN> struct OffsetOf<VLEN, 0> { static constexpr std::size_t value = 0; };
} // namespace detail

template <class MT, class Device, std::size_t VLEN> class AoSoA;

template <class... Ts, class Device, std::size_t VLEN>
class AoSoA<MemberTypes<Ts...>, Device, VLEN> {
public:
    static constexpr std::size_t vector_length = VLEN;
    static constexpr std::size_t bin_stride = detail::bin_bytes<VLEN, Ts...>();
    template <std::size_t I> using member_type = typename detail::Nth<I, Ts...>::type;
    template <std::size_t I> static constexpr std::size_t member_offset() { return detail::OffsetOf<VLEN, I, Ts...>::value; }

    AoSoA() {}
    AoSoA(const std::string&, std::size_t n = 0) { resize(n); }
    ~AoSoA() { std::free(buf_); }
    void resize(std::size_t n)
    {
        const std::size_t bins = (n + VLEN - 1) / VLEN, bytes = bins * bin_stride;
        if (bytes > cap_bytes_) { // 64-byte aligned backing store (VECTOR<double, 3> is aligned(32)); contents are kept
            char* nb = static_cast<char*>(std::aligned_alloc(64, (bytes + 63) / 64 * 64));
            std::memset(nb, 0, bytes);
            if (buf_) std::memcpy(nb, buf_, cap_bytes_);
            std::free(buf_);
            buf_ = nb;
            cap_bytes_ = bytes;
        }
        n_ = n;
    }
    void reserve(std::size_t n) { if (n > n_) resize(n); }
    std::size_t size() const { return n_; }
    std::size_t capacity() const { return (n_ + VLEN - 1) / VLEN * VLEN; }
    char* base() const { return buf_; }

    AoSoA(const AoSoA& o) { *this = o; }
    AoSoA& operator=(const AoSoA& o)
    {
        if (this == &o) return *this;
        resize(o.n_);
        if (o.n_) std::memcpy(buf_, o.buf_, (o.n_ + VLEN - 1) / VLEN * bin_stride);
        return *this;
    }

private:
    std::size_t n_ = 0, cap_bytes_ = 0;
    char* buf_ = nullptr;
};

template <class T, std::size_t EXT, std::size_t VLEN> struct Slice {
    char* base; // address of this member inside bin 0
    std::size_t bin_stride;
    std::size_t extent(int d) const { return d == 0 ? 0 : EXT; }
    T& operator()(std::size_t i) const { return *reinterpret_cast<T*>(base + (i / VLEN) * bin_stride + (i % VLEN) * sizeof(T)); }
    T& operator()(std::size_t i, std::size_t j) const
    {
        return *reinterpret_cast<T*>(base + (i / VLEN) * bin_stride + (j * VLEN + i % VLEN) * sizeof(T));
    }
};

template <std::size_t I, class AosoaT>
auto slice(const AosoaT& a)
{
    using M = typename AosoaT::template member_type<I>;
    using Info = detail::MemberInfo<M>;
    return Slice<typename Info::value_type, Info::extent, AosoaT::vector_length>{a.base() + AosoaT::template member_offset<I>(), AosoaT::bin_stride};
}

template <class A, class B> void deep_copy(A& dst, const B& src) { dst = src; }

} // namespace Cabana
