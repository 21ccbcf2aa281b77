// <H5Cpp.h> for a box without the HDF5 library: the part of the HDF5 C++ API that OpenDXMC's hdf5wrapper.cpp calls
// (R:src/libopendxmc/hdf5wrapper.cpp:73-360), implemented over libdxmc_b200's own HDF5 reader / writer (include/dxb.h:
// dxb_h5_*, opendxmc_b200/csrc/h5mini.*).  Files are REAL files in the HDF5 format: H5File(path, H5F_ACC_TRUNC) builds an
// image that close() / the destructor writes, H5File(path, H5F_ACC_RDONLY) parses one.  Test infrastructure (like the Qt
// stand-ins in this directory): it lets the reference's own save / load code, compiled unmodified, round-trip scenes and
// beams through the file format (oracle/ref_driver.cpp: h5roundtrip).
#pragma once
#include "dxb.h"

#include <cstdlib>
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
    DataType(H5T_class_t c, std::size_t s, bool sign = false) : cls(c), size(s), isSigned(sign) { }
    std::size_t getSize() const { return size; }
    H5T_class_t getClass() const { return cls; }
    // dxb_h5 type code of this in-memory type
    int code() const
    {
        if (cls == H5T_FLOAT)
            return size == 8 ? DXB_H5_F64 : (size == 4 ? DXB_H5_F32 : DXB_H5_UNKNOWN);
        if (cls == H5T_STRING)
            return DXB_H5_STRING;
        if (cls == H5T_INTEGER)
            switch (size) {
            case 8: return isSigned ? DXB_H5_I64 : DXB_H5_U64;
            case 4: return isSigned ? DXB_H5_I32 : DXB_H5_U32;
            case 2: return isSigned ? DXB_H5_I16 : DXB_H5_U16;
            case 1: return isSigned ? DXB_H5_I8 : DXB_H5_U8;
            default: break;
            }
        return DXB_H5_UNKNOWN;
    }
    static DataType fromCode(int c)
    {
        switch (c) {
        case DXB_H5_F64: return { H5T_FLOAT, 8 };
        case DXB_H5_F32: return { H5T_FLOAT, 4 };
        case DXB_H5_U64: return { H5T_INTEGER, 8 };
        case DXB_H5_I64: return { H5T_INTEGER, 8, true };
        case DXB_H5_U32: return { H5T_INTEGER, 4 };
        case DXB_H5_I32: return { H5T_INTEGER, 4, true };
        case DXB_H5_U16: return { H5T_INTEGER, 2 };
        case DXB_H5_I16: return { H5T_INTEGER, 2, true };
        case DXB_H5_U8: return { H5T_INTEGER, 1 };
        case DXB_H5_I8: return { H5T_INTEGER, 1, true };
        case DXB_H5_STRING: return { H5T_STRING, H5T_VARIABLE };
        default: return {};
        }
    }
    H5T_class_t cls = H5T_NO_CLASS;
    std::size_t size = 0;
    bool isSigned = false;
};
class PredType : public DataType {
public:
    using DataType::DataType;
    static const PredType NATIVE_DOUBLE, NATIVE_UINT64, NATIVE_UINT8, NATIVE_UINT, NATIVE_INT, C_S1;
};
inline const PredType PredType::NATIVE_DOUBLE { H5T_FLOAT, 8 };
inline const PredType PredType::NATIVE_UINT64 { H5T_INTEGER, 8 };
inline const PredType PredType::NATIVE_UINT8 { H5T_INTEGER, 1 };
inline const PredType PredType::NATIVE_UINT { H5T_INTEGER, 4 };
inline const PredType PredType::NATIVE_INT { H5T_INTEGER, 4, true };
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
    void setChunk(int, const hsize_t*) { chunked = true; }
    void setDeflate(int level) { deflate = level > 0; }
    bool chunked = false, deflate = false;
};

// the file image behind an H5File and everything opened from it
struct Image {
    dxb_h5* h = nullptr;
    std::string path;
    bool writing = false, closed = false;
    ~Image() { flush(); }
    void flush()
    {
        if (h && writing && !closed)
            dxb_h5_save(h, path.c_str());
        closed = true;
        if (h)
            dxb_h5_close(h);
        h = nullptr;
    }
};

class DataSet {
public:
    DataSet() = default;
    DataSet(std::shared_ptr<Image> i, std::string p, DataType t, DataSpace s, bool z) : img(std::move(i)), path(std::move(p)), type(t), space(std::move(s)), deflate(z) { }
    DataSpace getSpace() const { return space; }
    H5T_class_t getTypeClass() const { return type.cls; }
    DataType getDataType() const { return type; }
    void write(const void* buf, const DataType& memType)
    {
        const auto n = static_cast<std::size_t>(space.getSimpleExtentNpoints());
        int rc;
        if (memType.cls == H5T_STRING) {
            rc = dxb_h5_put_strings(img->h, path.c_str(), n, static_cast<const char* const*>(buf));
        } else {
            std::vector<uint64_t> d(space.dims.begin(), space.dims.end());
            rc = dxb_h5_put_dataset(img->h, path.c_str(), memType.code(), static_cast<int>(d.size()), d.data(), buf, deflate ? 1 : 0);
        }
        if (rc != DXB_OK)
            throw DataSetIException("DataSet::write", path);
    }
    void read(void* buf, const DataType& memType) const
    {
        const auto n = static_cast<std::size_t>(space.getSimpleExtentNpoints());
        if (memType.cls == H5T_STRING) {
            // like H5Dread of variable-length strings: one malloc'ed C string per element
            char** out = static_cast<char**>(buf);
            for (std::size_t i = 0; i < n; ++i) {
                const char* s = dxb_h5_dataset_string(img->h, path.c_str(), i);
                if (!s)
                    throw DataSetIException("DataSet::read", path);
                out[i] = static_cast<char*>(std::malloc(std::strlen(s) + 1));
                std::strcpy(out[i], s);
            }
        } else {
            if (memType.code() != type.code() || dxb_h5_dataset_read(img->h, path.c_str(), buf, n * type.size) != DXB_OK)
                throw DataSetIException("DataSet::read", path + ": stored type differs from the requested one (no conversion in this shim)");
        }
    }

private:
    std::shared_ptr<Image> img;
    std::string path;
    DataType type;
    DataSpace space;
    bool deflate = false;
};

class Attribute {
public:
    Attribute() = default;
    Attribute(std::shared_ptr<Image> i, std::string g, std::string n, DataType t, DataSpace s) : img(std::move(i)), group(std::move(g)), name(std::move(n)), type(t), space(std::move(s)) { }
    DataSpace getSpace() const { return space; }
    H5T_class_t getTypeClass() const { return type.cls; }
    FloatType getFloatType() const { return FloatType(type.cls, type.size); }
    IntType getIntType() const { return IntType(type.cls, type.size); }
    void write(const DataType& memType, const void* buf)
    {
        const long long n = space.dims.empty() ? -1 : static_cast<long long>(space.dims[0]);
        if (dxb_h5_put_attribute(img->h, group.c_str(), name.c_str(), memType.code(), n, buf) != DXB_OK)
            throw AttributeIException("Attribute::write", name);
    }
    void read(const DataType& memType, void* buf) const
    {
        const auto n = static_cast<std::size_t>(space.getSimpleExtentNpoints());
        if (memType.code() != type.code() || dxb_h5_attribute_read(img->h, group.c_str(), name.c_str(), buf, n * type.size) != DXB_OK)
            throw AttributeIException("Attribute::read", name);
    }

private:
    std::shared_ptr<Image> img;
    std::string group, name;
    DataType type;
    DataSpace space;
};

class H5Location {
public:
    H5Location() = default;
    H5Location(std::shared_ptr<Image> i, std::string p) : img(std::move(i)), path(std::move(p)) { }
    bool attrExists(const char* name) const { return dxb_h5_attribute_info(img->h, path.c_str(), name, nullptr, nullptr) == DXB_OK; }
    bool attrExists(const std::string& name) const { return attrExists(name.c_str()); }
    Attribute createAttribute(const char* name, const DataType& t, const DataSpace& sp)
    {
        if (sp.dims.size() > 1)
            throw AttributeIException("createAttribute", "rank > 1");
        return Attribute(img, path, name, t, sp);
    }
    Attribute openAttribute(const char* name) const
    {
        int code = 0;
        int64_t n = 0;
        if (dxb_h5_attribute_info(img->h, path.c_str(), name, &code, &n) != DXB_OK)
            throw AttributeIException("openAttribute", name);
        DataSpace sp;
        if (n >= 0) {
            const hsize_t d = static_cast<hsize_t>(n);
            sp = DataSpace(1, &d);
        }
        return Attribute(img, path, name, DataType::fromCode(code), sp);
    }

protected:
    std::shared_ptr<Image> img;
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
        img = std::make_shared<Image>();
        img->path = name;
        path = "/";
        if (flags == H5F_ACC_TRUNC) {
            img->h = dxb_h5_create();
            img->writing = true;
        } else if (dxb_h5_open(&img->h, name) != DXB_OK) {
            throw FileIException("H5File", std::string("cannot open ") + name);
        }
    }
    H5File(const std::string& name, unsigned flags) : H5File(name.c_str(), flags) { }
    bool nameExists(const char* p) const { return dxb_h5_exists(img->h, p) != 0; }
    bool nameExists(const std::string& p) const { return nameExists(p.c_str()); }
    Group createGroup(const char* p)
    {
        if (dxb_h5_make_group(img->h, p) != DXB_OK)
            throw GroupIException("createGroup", p);
        return Group(img, p);
    }
    Group openGroup(const char* p) const
    {
        if (dxb_h5_exists(img->h, p) != 1)
            throw GroupIException("openGroup", p);
        return Group(img, p);
    }
    DataSet createDataSet(const char* p, const DataType& t, const DataSpace& sp, const DSetCreatPropList& plist = DSetCreatPropList())
    {
        return DataSet(img, p, t, sp, plist.chunked && plist.deflate);
    }
    DataSet openDataSet(const char* p) const
    {
        int code = 0, rank = 0, z = 0;
        uint64_t dims[8];
        if (dxb_h5_dataset_info(img->h, p, &code, &rank, dims, &z) != DXB_OK)
            throw DataSetIException("openDataSet", p);
        std::vector<hsize_t> d(dims, dims + rank);
        return DataSet(img, p, DataType::fromCode(code), DataSpace(rank, d.data()), z != 0);
    }
    void close() { img->flush(); }
};

} // namespace H5
