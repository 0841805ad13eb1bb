// Reference: include/madrona/span.hpp -- non-owning (ptr,len) view that also
// binds to braced lists ("addToGraph<...>({dep_a, dep_b})").
#pragma once
#include <madrona/types.hpp>
#include <initializer_list>
namespace madrona {
namespace mwGPU {
template <typename T> struct RemoveConst { using type = T; };
template <typename T> struct RemoveConst<const T> { using type = T; };
}

template <typename T>
class Span {
public:
    MB2_HD constexpr Span(T *ptr, CountT n) : ptr_(ptr), n_(n) {}
    MB2_HD constexpr Span(
        std::initializer_list<typename mwGPU::RemoveConst<T>::type> l)
        : ptr_(l.begin()), n_((CountT)l.size()) {}
    template <typename U, CountT N>
    MB2_HD constexpr Span(U (&arr)[N]) : ptr_(arr), n_(N) {}
    MB2_HD constexpr T *data() const { return ptr_; }
    MB2_HD constexpr CountT size() const { return n_; }
    MB2_HD constexpr T &operator[](CountT i) const { return ptr_[i]; }
    MB2_HD constexpr T *begin() const { return ptr_; }
    MB2_HD constexpr T *end() const { return ptr_ + n_; }
private:
    T *ptr_;
    CountT n_;
};
}
