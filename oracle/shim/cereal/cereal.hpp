// Stand-in for the cereal serialization library (absent from this image; the reference pins it via conda,
// /root/reference/environment.yml).  TEST INFRASTRUCTURE ONLY: it lets the reference's own translation units compile
// into oracle/_ref AND makes their archives work, so that lm::save_state_to_file / load_state_from_file
// (src/user.cpp:104-112) round-trip an asset tree — including one that holds the plugin's accel::cuda_bvh (SURVEY.md
// row N4).  Only the part of cereal's interface the reference uses (include/lm/serial.h, include/lm/serialtype.h and the
// serialize() members of its types) is provided:
//
//   * an archive's operator() takes any number of values and serialises each with the first of
//         t.serialize(ar)  |  serialize(ar, t)  |  t.save(ar) / t.load(ar)  |  save(ar, t) / load(ar, t)
//     (non-member functions are found in namespace cereal and by argument-dependent lookup, as with cereal); enums go
//     through their underlying type;
//   * NameValuePair / SizeTag / BinaryData wrappers and the CEREAL_* macros the reference spells out;
//   * cereal/types/{string,vector,unordered_map}.hpp.
//
// The byte stream is this shim's own (sizes as 64-bit, arithmetic values raw): files written here are read back here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <iosfwd>
#include <type_traits>
#include <utility>

#define CEREAL_SAVE_FUNCTION_NAME save
#define CEREAL_LOAD_FUNCTION_NAME load
#define CEREAL_SERIALIZE_FUNCTION_NAME serialize
#define CEREAL_NVP_(name, value) ::cereal::make_nvp_(name, value)
#define CEREAL_NVP(value) ::cereal::make_nvp_(#value, value)
#define CEREAL_REGISTER_ARCHIVE(Archive)
#define CEREAL_SETUP_ARCHIVE_TRAITS(In, Out)
#define CEREAL_ARCHIVE_RESTRICT(In, Out) \
    typename std::enable_if<std::is_same<Archive, In>::value || std::is_same<Archive, Out>::value, void>::type

namespace cereal {

enum Flags { AllowEmptyClassElision = 1 };
using size_type = std::uint64_t;

template <class T> struct NameValuePair { T&& value; };
template <class T> struct SizeTag { T&& size; };
template <class T> struct BinaryData { T data; std::uint64_t size; };

template <class T> NameValuePair<T> make_nvp_(const char*, T&& v) { return { std::forward<T>(v) }; }
template <class T> NameValuePair<T> make_nvp(const char*, T&& v) { return { std::forward<T>(v) }; }
template <class T> SizeTag<T> make_size_tag(T&& v) { return { std::forward<T>(v) }; }
template <class T> BinaryData<T> binary_data(T&& d, std::size_t n) { return { std::forward<T>(d), n }; }

namespace detail {
template <int N> struct prio : prio<N - 1> {};
template <> struct prio<0> {};

// -- output: the value is seen as a non-const lvalue for serialize() (one function for both directions, as in cereal)
template <class A, class T> auto out(A& ar, T& t, prio<4>) -> decltype(t.serialize(ar), void()) { t.serialize(ar); }
template <class A, class T> auto out(A& ar, T& t, prio<3>) -> decltype(serialize(ar, t), void()) { serialize(ar, t); }
template <class A, class T> auto out(A& ar, T& t, prio<2>) -> decltype(std::as_const(t).save(ar), void()) { std::as_const(t).save(ar); }
template <class A, class T> auto out(A& ar, T& t, prio<1>) -> decltype(save(ar, std::as_const(t)), void()) { save(ar, std::as_const(t)); }
template <class A, class T> auto out(A& ar, T& t, prio<0>) -> std::enable_if_t<std::is_enum<T>::value> {
    auto v = static_cast<std::underlying_type_t<T>>(t);
    ar(v);
}
// -- input
template <class A, class T> auto in(A& ar, T& t, prio<4>) -> decltype(t.serialize(ar), void()) { t.serialize(ar); }
template <class A, class T> auto in(A& ar, T& t, prio<3>) -> decltype(serialize(ar, t), void()) { serialize(ar, t); }
template <class A, class T> auto in(A& ar, T& t, prio<2>) -> decltype(t.load(ar), void()) { t.load(ar); }
template <class A, class T> auto in(A& ar, T& t, prio<1>) -> decltype(load(ar, t), void()) { load(ar, t); }
template <class A, class T> auto in(A& ar, T& t, prio<0>) -> std::enable_if_t<std::is_enum<T>::value> {
    std::underlying_type_t<T> v{};
    ar(v);
    t = static_cast<T>(v);
}
}  // namespace detail

template <class Derived, int F = 0>
class OutputArchive {
public:
    explicit OutputArchive(Derived* d) : self_(d) {}
    template <class... Ts> Derived& operator()(Ts&&... args) {
        (process(args), ...);
        return *self_;
    }
private:
    template <class T> void process(T& t) {
        detail::out(*self_, const_cast<std::remove_const_t<T>&>(t), detail::prio<4>{});
    }
    Derived* self_;
};

template <class Derived, int F = 0>
class InputArchive {
public:
    explicit InputArchive(Derived* d) : self_(d) {}
    template <class... Ts> Derived& operator()(Ts&&... args) {
        (process(args), ...);
        return *self_;
    }
private:
    template <class T> void process(T& t) {
        detail::in(*self_, const_cast<std::remove_const_t<T>&>(t), detail::prio<4>{});
    }
    Derived* self_;
};

}
