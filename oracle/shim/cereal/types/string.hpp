// cereal/types/string.hpp stand-in (see ../cereal.hpp): size, then the characters
#pragma once
#include "../cereal.hpp"
#include <string>
namespace cereal {
template <class Archive, class C, class Tr, class Al>
void save(Archive& ar, const std::basic_string<C, Tr, Al>& s) {
    ar(make_size_tag(static_cast<size_type>(s.size())));
    ar(binary_data(s.data(), s.size() * sizeof(C)));
}
template <class Archive, class C, class Tr, class Al>
void load(Archive& ar, std::basic_string<C, Tr, Al>& s) {
    size_type n = 0;
    ar(make_size_tag(n));
    s.resize(static_cast<std::size_t>(n));
    ar(binary_data(const_cast<C*>(s.data()), static_cast<std::size_t>(n) * sizeof(C)));
}
}
