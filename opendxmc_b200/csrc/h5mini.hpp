// h5mini.hpp — a minimal reader / writer for the subset of the HDF5 file format that OpenDXMC's save files use
// (R:src/libopendxmc/hdf5wrapper.cpp:73-360 shows every HDF5 call the application makes):
//   * groups (nested, "old style": symbol table = B-tree v1 + local heap + symbol nodes),
//   * attributes on groups: scalar or 1-D, 64-bit float / unsigned integer,
//   * datasets of f64 / u64 / u8 / variable-length strings, rank 1-3, stored contiguously or as ONE chunk with the
//     deflate (zlib, level 6) filter - what H5::DSetCreatPropList::setChunk(whole extent) + setDeflate(6) produces
//     (R:...hdf5wrapper.cpp:145-151); the reader walks chunk B-trees of any size and also accepts compact layout,
//     fixed-length strings and the shuffle filter.
// File structures follow the public "HDF5 File Format Specification Version 2.0": superblock version 0 (read: 0 and 1,
// with a user block), version-1 object headers, version-1 B-trees, version-1 attribute / dataspace / datatype
// messages, version-3 data layout, version-1 filter pipeline, global heap collections for variable-length strings.
// No HDF5 library exists in this image; the reader is pinned against a genuine libhdf5-written file that ships with
// scipy (tests/test_h5mini.py), the writer against the reader.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace h5mini {

enum class Type : uint8_t { F64, U64, U8, I64, I32, U32, F32, U16, I16, I8, String, Unknown };

size_t typeSize(Type t);

struct Attribute {
    Type type = Type::Unknown;
    std::vector<uint64_t> dims; // empty: scalar
    std::vector<uint8_t> data;  // raw little-endian elements (numeric types)
    std::vector<std::string> strings; // Type::String
    uint64_t count() const;
};

struct Dataset {
    Type type = Type::Unknown;
    std::vector<uint64_t> dims; // HDF5 order (slowest first)
    std::vector<uint8_t> data;  // raw little-endian elements (numeric types)
    std::vector<std::string> strings; // Type::String
    bool deflate = false;       // writer: one chunk + deflate level 6; reader: the dataset was filtered
    uint64_t count() const;
};

struct Group {
    std::map<std::string, std::unique_ptr<Group>> groups;
    std::map<std::string, Dataset> datasets;
    std::map<std::string, Attribute> attributes;
    std::vector<std::string> attributeOrder; // creation order (the writer keeps it)
};

// An in-memory image of a file: load() parses a file into it, save() serialises it.
class File {
public:
    Group root;
    static std::unique_ptr<File> load(const std::string& path, std::string* error = nullptr);
    bool save(const std::string& path, std::string* error = nullptr) const;

    // path helpers ("/a/b/c" or "a/b/c")
    static std::vector<std::string> split(const std::string& path);
    Group* group(const std::string& path, bool create = false);
    const Group* group(const std::string& path) const;
    Dataset* dataset(const std::string& path);
    const Dataset* dataset(const std::string& path) const;
    bool exists(const std::string& path) const; // group or dataset
    Dataset& createDataset(const std::string& path); // creates the parent groups
};

} // namespace h5mini
