// Stand-in for the cereal serialization library (absent from this image; the
// reference pins it via conda, /root/reference/environment.yml).  TEST
// INFRASTRUCTURE ONLY: it lets the reference's own translation units compile
// into oracle/_ref.  Serialization (lm::save_state_to_file)
// is out of scope, so archives here accept every expression the reference
// writes (include/lm/serial.h, include/lm/serialtype.h) and perform no I/O.
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

template <class T> struct NameValuePair { T&& value; };
template <class T> struct SizeTag { T&& size; };
template <class T> struct BinaryData { T data; std::uint64_t size; };

template <class T> NameValuePair<T> make_nvp_(const char*, T&& v) { return { std::forward<T>(v) }; }
template <class T> NameValuePair<T> make_nvp(const char*, T&& v) { return { std::forward<T>(v) }; }
template <class T> SizeTag<T> make_size_tag(T&& v) { return { std::forward<T>(v) }; }
template <class T> BinaryData<T> binary_data(T&& d, std::size_t n) { return { std::forward<T>(d), n }; }

// Archives: operator() swallows its arguments (no-op serialization).
template <class Derived, int F = 0>
class OutputArchive {
public:
    explicit OutputArchive(Derived* d) : self_(d) {}
    template <class... Ts> Derived& operator()(Ts&&...) { return *self_; }
private:
    Derived* self_;
};

template <class Derived, int F = 0>
class InputArchive {
public:
    explicit InputArchive(Derived* d) : self_(d) {}
    // Value-initialise plain outputs so loads never read garbage.
    template <class... Ts> Derived& operator()(Ts&&...) { return *self_; }
private:
    Derived* self_;
};

}
