// cereal/types/unordered_map.hpp stand-in (see ../cereal.hpp): size, then (key, value) pairs
#pragma once
#include "../cereal.hpp"
#include <unordered_map>
namespace cereal {
template <class Archive, class K, class V, class H, class E, class A>
void save(Archive& ar, const std::unordered_map<K, V, H, E, A>& m) {
    ar(make_size_tag(static_cast<size_type>(m.size())));
    for (const auto& kv : m) ar(kv.first, kv.second);
}
template <class Archive, class K, class V, class H, class E, class A>
void load(Archive& ar, std::unordered_map<K, V, H, E, A>& m) {
    size_type n = 0;
    ar(make_size_tag(n));
    m.clear();
    for (size_type i = 0; i < n; i++) {
        K k{};
        V v{};
        ar(k, v);
        m.emplace(std::move(k), std::move(v));
    }
}
}
