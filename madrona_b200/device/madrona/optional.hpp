// Minimal madrona::Optional (reference: include/madrona/optional.hpp).
#pragma once
#include <madrona/types.hpp>
#include <new>
namespace madrona {
template <typename T>
class Optional {
public:
    MB2_HD static Optional none() { return Optional(); }
    MB2_HD static Optional make(const T &v) { return Optional(v); }
    MB2_HD Optional() : has_(false) {}
    MB2_HD Optional(const T &v) : has_(true) { new (&storage_) T(v); }
    MB2_HD bool has_value() const { return has_; }
    MB2_HD T &operator*() { return *reinterpret_cast<T *>(&storage_); }
    MB2_HD const T &operator*() const { return *reinterpret_cast<const T *>(&storage_); }
    MB2_HD T *operator->() { return reinterpret_cast<T *>(&storage_); }
    MB2_HD const T *operator->() const { return reinterpret_cast<const T *>(&storage_); }
private:
    alignas(T) char storage_[sizeof(T)];
    bool has_;
};
}
