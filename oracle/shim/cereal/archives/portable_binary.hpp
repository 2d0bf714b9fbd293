// Stand-in for cereal's portable binary archives (see ../cereal.hpp).
#pragma once
#include "../cereal.hpp"
#include <istream>
#include <ostream>
namespace cereal {
class PortableBinaryOutputArchive : public OutputArchive<PortableBinaryOutputArchive, AllowEmptyClassElision> {
public:
    explicit PortableBinaryOutputArchive(std::ostream& s)
        : OutputArchive<PortableBinaryOutputArchive, AllowEmptyClassElision>(this), s_(s) {}
    template <std::size_t DataSize> void saveBinary(const void* d, std::size_t n) {
        s_.write(static_cast<const char*>(d), std::streamsize(n));
    }
private:
    std::ostream& s_;
};
class PortableBinaryInputArchive : public InputArchive<PortableBinaryInputArchive, AllowEmptyClassElision> {
public:
    explicit PortableBinaryInputArchive(std::istream& s)
        : InputArchive<PortableBinaryInputArchive, AllowEmptyClassElision>(this), s_(s) {}
    template <std::size_t DataSize> void loadBinary(void* const d, std::size_t n) {
        s_.read(static_cast<char*>(d), std::streamsize(n));
    }
private:
    std::istream& s_;
};
}
