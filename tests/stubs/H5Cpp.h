// In-memory stand-in for the HDF5 C++ API (<H5Cpp.h>), tests only (see QObject in this directory): files, groups,
// data sets and attributes live in process memory, keyed by the file path, so that OpenDXMC's hdf5wrapper.cpp can save
// a scene and load it back in the same process.  Not a file format.
#pragma once
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

using hsize_t = unsigned long long;
using hid_t = long long;
enum H5T_class_t { H5T_NO_CLASS = -1, H5T_INTEGER = 0, H5T_FLOAT = 1, H5T_STRING = 3 };
constexpr std::size_t H5T_VARIABLE = static_cast<std::size_t>(-1);
constexpr unsigned H5F_ACC_RDONLY = 0u, H5F_ACC_TRUNC = 2u;

namespace H5 {

class Exception : public std::runtime_error {
public:
    Exception(const std::string& f = "", const std::string& m = "") : std::runtime_error(m), func(f), msg(m) { }
    static void dontPrint() { }
    void printErrorStack() const { }
    std::string getFuncName() const { return func; }
    std::string getDetailMsg() const { return msg; }
    const char* getCFuncName() const { return func.c_str(); }
    const char* getCDetailMsg() const { return msg.c_str(); }

private:
    std::string func, msg;
};
class FileIException : public Exception { using Exception::Exception; };
class DataSetIException : public Exception { using Exception::Exception; };
class DataSpaceIException : public Exception { using Exception::Exception; };
class GroupIException : public Exception { using Exception::Exception; };
class AttributeIException : public Exception { using Exception::Exception; };

class DataType {
public:
    DataType() = default;
    DataType(H5T_class_t c, std::size_t s) : cls(c), size(s) { }
    std::size_t getSize() const { return size; }
    H5T_class_t getClass() const { return cls; }
    H5T_class_t cls = H5T_NO_CLASS;
    std::size_t size = 0;
};
class PredType : public DataType {
public:
    using DataType::DataType;
    static inline const DataType& mk(H5T_class_t c, std::size_t s)
    {
        static std::map<std::pair<int, std::size_t>, PredType> all;
        auto& t = all[{ c, s }];
        t.cls = c;
        t.size = s;
        return t;
    }
    static const PredType NATIVE_DOUBLE, NATIVE_UINT64, NATIVE_UINT8, NATIVE_UINT, NATIVE_INT, C_S1;
};
inline const PredType PredType::NATIVE_DOUBLE { H5T_FLOAT, 8 };
inline const PredType PredType::NATIVE_UINT64 { H5T_INTEGER, 8 };
inline const PredType PredType::NATIVE_UINT8 { H5T_INTEGER, 1 };
inline const PredType PredType::NATIVE_UINT { H5T_INTEGER, 4 };
inline const PredType PredType::NATIVE_INT { H5T_INTEGER, 4 };
inline const PredType PredType::C_S1 { H5T_STRING, 1 };
class StrType : public DataType {
public:
    StrType() = default;
    StrType(const PredType&, std::size_t s) : DataType(H5T_STRING, s) { }
};
class FloatType : public DataType { using DataType::DataType; };
class IntType : public DataType { using DataType::DataType; };

class DataSpace {
public:
    DataSpace() = default; // scalar
    DataSpace(int rank, const hsize_t* d) : dims(d, d + rank) { }
    int getSimpleExtentNdims() const { return static_cast<int>(dims.size()); }
    int getSimpleExtentDims(hsize_t* out) const
    {
        for (std::size_t i = 0; i < dims.size(); ++i)
            out[i] = dims[i];
        return static_cast<int>(dims.size());
    }
    long long getSimpleExtentNpoints() const
    {
        long long n = 1;
        for (auto d : dims)
            n *= static_cast<long long>(d);
        return n;
    }
    std::vector<hsize_t> dims;
};

class DSetCreatPropList {
public:
    void setChunk(int, const hsize_t*) { }
    void setDeflate(int) { }
};

struct Node { // data set or attribute payload
    DataType type;
    DataSpace space;
    std::vector<unsigned char> bytes;
    std::vector<std::string> strings;
};
struct Store {
    std::map<std::string, std::shared_ptr<Node>> datasets;
    std::map<std::string, std::map<std::string, std::shared_ptr<Node>>> groups; // group path -> attributes
};

class AbstractDs {
public:
    explicit AbstractDs(std::shared_ptr<Node> n = nullptr) : node(std::move(n)) { }
    DataSpace getSpace() const { return node->space; }
    H5T_class_t getTypeClass() const { return node->type.cls; }
    FloatType getFloatType() const { return FloatType(node->type.cls, node->type.size); }
    IntType getIntType() const { return IntType(node->type.cls, node->type.size); }
    DataType getDataType() const { return node->type; }

protected:
    void put(const void* buf, const DataType& t)
    {
        node->type = t;
        const auto n = static_cast<std::size_t>(node->space.getSimpleExtentNpoints());
        if (t.cls == H5T_STRING) {
            const char* const* s = static_cast<const char* const*>(buf);
            node->strings.assign(s, s + n);
        } else {
            node->bytes.assign(static_cast<const unsigned char*>(buf), static_cast<const unsigned char*>(buf) + n * t.size);
        }
    }
    void get(void* buf, const DataType& t) const
    {
        if (t.cls == H5T_STRING) {
            char** out = static_cast<char**>(buf);
            for (std::size_t i = 0; i < node->strings.size(); ++i) {
                out[i] = static_cast<char*>(std::malloc(node->strings[i].size() + 1));
                std::memcpy(out[i], node->strings[i].c_str(), node->strings[i].size() + 1);
            }
        } else {
            std::memcpy(buf, node->bytes.data(), node->bytes.size());
        }
    }
    std::shared_ptr<Node> node;
};
class DataSet : public AbstractDs {
public:
    using AbstractDs::AbstractDs;
    void write(const void* buf, const DataType& t) { put(buf, t); }
    void read(void* buf, const DataType& t) const { get(buf, t); }
};
class Attribute : public AbstractDs {
public:
    using AbstractDs::AbstractDs;
    void write(const DataType& t, const void* buf) { put(buf, t); }
    void read(const DataType& t, void* buf) const { get(buf, t); }
};

class H5Location {
public:
    H5Location() = default;
    H5Location(std::shared_ptr<Store> s, std::string p) : store(std::move(s)), path(std::move(p)) { }
    bool attrExists(const char* name) const { return store->groups[path].count(name) != 0; }
    bool attrExists(const std::string& name) const { return attrExists(name.c_str()); }
    Attribute createAttribute(const char* name, const DataType& t, const DataSpace& sp)
    {
        auto n = std::make_shared<Node>();
        n->type = t;
        n->space = sp;
        store->groups[path][name] = n;
        return Attribute(n);
    }
    Attribute openAttribute(const char* name) const
    {
        auto it = store->groups[path].find(name);
        if (it == store->groups[path].end())
            throw AttributeIException("openAttribute", name);
        return Attribute(it->second);
    }

protected:
    std::shared_ptr<Store> store;
    std::string path;
};
class Group : public H5Location {
public:
    using H5Location::H5Location;
};
class H5File : public H5Location {
public:
    H5File(const char* name, unsigned flags)
    {
        static std::map<std::string, std::shared_ptr<Store>> files;
        if (flags == H5F_ACC_TRUNC || !files.count(name)) {
            if (flags == H5F_ACC_RDONLY)
                throw FileIException("H5File", std::string("no such file: ") + name);
            files[name] = std::make_shared<Store>();
        }
        store = files[name];
        path = "/";
        store->groups["/"];
    }
    H5File(const std::string& name, unsigned flags) : H5File(name.c_str(), flags) { }
    static std::string norm(const std::string& p) { return !p.empty() && p[0] == '/' ? p : "/" + p; }
    bool nameExists(const char* p) const { return store->groups.count(norm(p)) || store->datasets.count(norm(p)); }
    bool nameExists(const std::string& p) const { return nameExists(p.c_str()); }
    Group createGroup(const char* p)
    {
        store->groups[norm(p)];
        return Group(store, norm(p));
    }
    Group openGroup(const char* p) const
    {
        if (!store->groups.count(norm(p)))
            throw GroupIException("openGroup", p);
        return Group(store, norm(p));
    }
    DataSet createDataSet(const char* p, const DataType& t, const DataSpace& sp, const DSetCreatPropList& = DSetCreatPropList())
    {
        auto n = std::make_shared<Node>();
        n->type = t;
        n->space = sp;
        store->datasets[norm(p)] = n;
        return DataSet(n);
    }
    DataSet openDataSet(const char* p) const
    {
        auto it = store->datasets.find(norm(p));
        if (it == store->datasets.end())
            throw DataSetIException("openDataSet", p);
        return DataSet(it->second);
    }
    void close() { }
};

} // namespace H5
