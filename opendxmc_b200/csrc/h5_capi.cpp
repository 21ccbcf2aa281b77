// h5_capi.cpp — C ABI over h5mini (the save-file format of OpenDXMC): a generic object-level API
// (dxb_h5_*: what the H5Cpp shim under tests/stubs/ and the Python tests drive).  The scene-level entry points
// dxb_save_scene / dxb_load_scene, which move the arrays of R:src/libopendxmc/hdf5wrapper.cpp:384-459 between a file
// and a context, live in context.cu next to the device buffers they fill.
#include "../../include/dxb.h"
#include "h5mini.hpp"

#include <cstring>
#include <string>

struct dxb_h5 {
    h5mini::File file;
    std::string error;
    std::string scratch;
};

namespace {

h5mini::Type typeOf(int t)
{
    switch (t) {
    case DXB_H5_F64: return h5mini::Type::F64;
    case DXB_H5_U64: return h5mini::Type::U64;
    case DXB_H5_U8: return h5mini::Type::U8;
    case DXB_H5_STRING: return h5mini::Type::String;
    case DXB_H5_I64: return h5mini::Type::I64;
    case DXB_H5_I32: return h5mini::Type::I32;
    case DXB_H5_U32: return h5mini::Type::U32;
    case DXB_H5_F32: return h5mini::Type::F32;
    case DXB_H5_U16: return h5mini::Type::U16;
    case DXB_H5_I16: return h5mini::Type::I16;
    case DXB_H5_I8: return h5mini::Type::I8;
    default: return h5mini::Type::Unknown;
    }
}
int codeOf(h5mini::Type t)
{
    switch (t) {
    case h5mini::Type::F64: return DXB_H5_F64;
    case h5mini::Type::U64: return DXB_H5_U64;
    case h5mini::Type::U8: return DXB_H5_U8;
    case h5mini::Type::String: return DXB_H5_STRING;
    case h5mini::Type::I64: return DXB_H5_I64;
    case h5mini::Type::I32: return DXB_H5_I32;
    case h5mini::Type::U32: return DXB_H5_U32;
    case h5mini::Type::F32: return DXB_H5_F32;
    case h5mini::Type::U16: return DXB_H5_U16;
    case h5mini::Type::I16: return DXB_H5_I16;
    case h5mini::Type::I8: return DXB_H5_I8;
    default: return DXB_H5_UNKNOWN;
    }
}

} // namespace

extern "C" {

dxb_h5* dxb_h5_create(void) { return new dxb_h5(); }

int dxb_h5_open(dxb_h5** out, const char* path)
{
    if (!out || !path)
        return DXB_EINVAL;
    *out = nullptr;
    std::string err;
    auto f = h5mini::File::load(path, &err);
    if (!f)
        return DXB_EINVAL;
    auto* h = new dxb_h5();
    h->file = std::move(*f);
    *out = h;
    return DXB_OK;
}

void dxb_h5_close(dxb_h5* h) { delete h; }

const char* dxb_h5_error(const dxb_h5* h) { return h ? h->error.c_str() : "null handle"; }

int dxb_h5_save(dxb_h5* h, const char* path)
{
    if (!h || !path)
        return DXB_EINVAL;
    return h->file.save(path, &h->error) ? DXB_OK : DXB_EINVAL;
}

int dxb_h5_exists(const dxb_h5* h, const char* path)
{
    if (!h || !path)
        return 0;
    if (h->file.dataset(path))
        return 2;
    return h->file.group(path) ? 1 : 0;
}

int dxb_h5_make_group(dxb_h5* h, const char* path)
{
    if (!h || !path)
        return DXB_EINVAL;
    return h->file.group(path, true) ? DXB_OK : DXB_EINVAL;
}

const char* dxb_h5_list(dxb_h5* h, const char* group_path)
{
    if (!h || !group_path)
        return "";
    const h5mini::Group* g = h->file.group(group_path);
    h->scratch.clear();
    if (!g)
        return "";
    for (const auto& kv : g->groups)
        h->scratch += "g " + kv.first + "\n";
    for (const auto& kv : g->datasets)
        h->scratch += "d " + kv.first + "\n";
    for (const std::string& a : g->attributeOrder)
        h->scratch += "a " + a + "\n";
    return h->scratch.c_str();
}

int dxb_h5_put_dataset(dxb_h5* h, const char* path, int type, int rank, const uint64_t* dims, const void* data, int deflate)
{
    if (!h || !path || rank < 0 || rank > 8 || (rank > 0 && !dims))
        return DXB_EINVAL;
    const h5mini::Type t = typeOf(type);
    if (t == h5mini::Type::Unknown || t == h5mini::Type::String)
        return DXB_EINVAL;
    try {
        h5mini::Dataset& d = h->file.createDataset(path);
        d = h5mini::Dataset();
        d.type = t;
        d.dims.assign(dims, dims + rank);
        d.deflate = deflate != 0;
        const size_t bytes = d.count() * h5mini::typeSize(t);
        if (bytes && !data)
            return DXB_EINVAL;
        d.data.assign(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + bytes);
    } catch (const std::exception& e) {
        h->error = e.what();
        return DXB_EINVAL;
    }
    return DXB_OK;
}

int dxb_h5_put_strings(dxb_h5* h, const char* path, uint64_t n, const char* const* strings)
{
    if (!h || !path || (n && !strings))
        return DXB_EINVAL;
    try {
        h5mini::Dataset& d = h->file.createDataset(path);
        d = h5mini::Dataset();
        d.type = h5mini::Type::String;
        d.dims = { n };
        for (uint64_t i = 0; i < n; ++i)
            d.strings.emplace_back(strings[i] ? strings[i] : "");
    } catch (const std::exception& e) {
        h->error = e.what();
        return DXB_EINVAL;
    }
    return DXB_OK;
}

int dxb_h5_put_attribute(dxb_h5* h, const char* group_path, const char* name, int type, int64_t n, const void* data)
{
    if (!h || !group_path || !name || !data)
        return DXB_EINVAL;
    const h5mini::Type t = typeOf(type);
    if (t == h5mini::Type::Unknown || t == h5mini::Type::String)
        return DXB_EINVAL;
    h5mini::Group* g = h->file.group(group_path, true);
    if (!g)
        return DXB_EINVAL;
    h5mini::Attribute a;
    a.type = t;
    if (n >= 0)
        a.dims = { static_cast<uint64_t>(n) }; // n < 0: scalar dataspace
    const size_t bytes = a.count() * h5mini::typeSize(t);
    a.data.assign(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + bytes);
    if (!g->attributes.count(name))
        g->attributeOrder.push_back(name);
    g->attributes[name] = std::move(a);
    return DXB_OK;
}

int dxb_h5_dataset_info(const dxb_h5* h, const char* path, int* type, int* rank, uint64_t dims[8], int* deflate)
{
    if (!h || !path)
        return DXB_EINVAL;
    const h5mini::Dataset* d = h->file.dataset(path);
    if (!d)
        return DXB_EINVAL;
    if (type)
        *type = codeOf(d->type);
    if (rank)
        *rank = static_cast<int>(d->dims.size());
    if (dims)
        for (size_t i = 0; i < d->dims.size() && i < 8; ++i)
            dims[i] = d->dims[i];
    if (deflate)
        *deflate = d->deflate ? 1 : 0;
    return DXB_OK;
}

int dxb_h5_dataset_read(const dxb_h5* h, const char* path, void* out, uint64_t out_bytes)
{
    if (!h || !path || !out)
        return DXB_EINVAL;
    const h5mini::Dataset* d = h->file.dataset(path);
    if (!d || d->type == h5mini::Type::String || d->data.size() != out_bytes)
        return DXB_EINVAL;
    std::memcpy(out, d->data.data(), out_bytes);
    return DXB_OK;
}

const char* dxb_h5_dataset_string(const dxb_h5* h, const char* path, uint64_t index)
{
    if (!h || !path)
        return nullptr;
    const h5mini::Dataset* d = h->file.dataset(path);
    if (!d || d->type != h5mini::Type::String || index >= d->strings.size())
        return nullptr;
    return d->strings[index].c_str();
}

int dxb_h5_attribute_info(const dxb_h5* h, const char* group_path, const char* name, int* type, int64_t* n)
{
    if (!h || !group_path || !name)
        return DXB_EINVAL;
    const h5mini::Group* g = h->file.group(group_path);
    if (!g)
        return DXB_EINVAL;
    auto it = g->attributes.find(name);
    if (it == g->attributes.end())
        return DXB_EINVAL;
    if (type)
        *type = codeOf(it->second.type);
    if (n)
        *n = it->second.dims.empty() ? -1 : static_cast<int64_t>(it->second.count());
    return DXB_OK;
}

int dxb_h5_attribute_read(const dxb_h5* h, const char* group_path, const char* name, void* out, uint64_t out_bytes)
{
    if (!h || !group_path || !name || !out)
        return DXB_EINVAL;
    const h5mini::Group* g = h->file.group(group_path);
    if (!g)
        return DXB_EINVAL;
    auto it = g->attributes.find(name);
    if (it == g->attributes.end() || it->second.type == h5mini::Type::String || it->second.data.size() != out_bytes)
        return DXB_EINVAL;
    std::memcpy(out, it->second.data.data(), out_bytes);
    return DXB_OK;
}

const char* dxb_h5_attribute_string(const dxb_h5* h, const char* group_path, const char* name, uint64_t index)
{
    if (!h || !group_path || !name)
        return nullptr;
    const h5mini::Group* g = h->file.group(group_path);
    if (!g)
        return nullptr;
    auto it = g->attributes.find(name);
    if (it == g->attributes.end() || it->second.type != h5mini::Type::String || index >= it->second.strings.size())
        return nullptr;
    return it->second.strings[index].c_str();
}

} // extern "C"
