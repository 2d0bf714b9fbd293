// cereal/types/vector.hpp stand-in (see ../cereal.hpp): size, then the elements (arithmetic element types as one block)
#pragma once
#include "../cereal.hpp"
#include <vector>
namespace cereal {
template <class Archive, class T, class A>
std::enable_if_t<std::is_arithmetic<T>::value && !std::is_same<T, bool>::value> save(Archive& ar, const std::vector<T, A>& v) {
    ar(make_size_tag(static_cast<size_type>(v.size())));
    ar(binary_data(v.data(), v.size() * sizeof(T)));
}
template <class Archive, class T, class A>
std::enable_if_t<std::is_arithmetic<T>::value && !std::is_same<T, bool>::value> load(Archive& ar, std::vector<T, A>& v) {
    size_type n = 0;
    ar(make_size_tag(n));
    v.resize(static_cast<std::size_t>(n));
    ar(binary_data(v.data(), static_cast<std::size_t>(n) * sizeof(T)));
}
template <class Archive, class T, class A>
std::enable_if_t<!std::is_arithmetic<T>::value> save(Archive& ar, const std::vector<T, A>& v) {
    ar(make_size_tag(static_cast<size_type>(v.size())));
    for (const auto& e : v) ar(e);
}
template <class Archive, class T, class A>
std::enable_if_t<!std::is_arithmetic<T>::value> load(Archive& ar, std::vector<T, A>& v) {
    size_type n = 0;
    ar(make_size_tag(n));
    v.clear();
    v.resize(static_cast<std::size_t>(n));
    for (auto& e : v) ar(e);
}
template <class Archive, class A>
void save(Archive& ar, const std::vector<bool, A>& v) {
    ar(make_size_tag(static_cast<size_type>(v.size())));
    for (const bool b : v) { const std::uint8_t x = b ? 1 : 0; ar(x); }
}
template <class Archive, class A>
void load(Archive& ar, std::vector<bool, A>& v) {
    size_type n = 0;
    ar(make_size_tag(n));
    v.resize(static_cast<std::size_t>(n));
    for (std::size_t i = 0; i < v.size(); i++) { std::uint8_t x = 0; ar(x); v[i] = x != 0; }
}
}
