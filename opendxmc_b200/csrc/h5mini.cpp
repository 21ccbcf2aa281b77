// h5mini.cpp — see h5mini.hpp.  Structures per the "HDF5 File Format Specification Version 2.0" (section numbers in the
// comments): II.A superblock, III.A B-trees v1, III.C symbol nodes, III.D local heaps, III.E global heaps, IV.A object
// headers and their messages.
#include "h5mini.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <zlib.h>

namespace h5mini {

namespace {

constexpr uint64_t kUndef = 0xFFFFFFFFFFFFFFFFull;
const unsigned char kSignature[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };
constexpr int kGroupLeafK = 32;     // symbol node: up to 2K entries
constexpr int kGroupInternalK = 16; // group B-tree node: up to 2K children
constexpr int kChunkK = 32;         // chunk B-tree node (the library's default for version-0 superblocks)

struct Err : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------ little-endian byte buffer (writer)
struct Out {
    std::vector<uint8_t> b;
    uint64_t size() const { return b.size(); }
    void align(size_t a = 8)
    {
        while (b.size() % a)
            b.push_back(0);
    }
    uint64_t reserve(size_t n)
    {
        align();
        const uint64_t at = b.size();
        b.resize(b.size() + n, 0);
        return at;
    }
    void put(uint64_t at, const void* p, size_t n) { std::memcpy(b.data() + at, p, n); }
};

struct Buf { // a message / structure under construction
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void u16(uint16_t v)
    {
        for (int i = 0; i < 2; ++i)
            b.push_back(static_cast<uint8_t>(v >> (8 * i)));
    }
    void u32(uint32_t v)
    {
        for (int i = 0; i < 4; ++i)
            b.push_back(static_cast<uint8_t>(v >> (8 * i)));
    }
    void u64(uint64_t v)
    {
        for (int i = 0; i < 8; ++i)
            b.push_back(static_cast<uint8_t>(v >> (8 * i)));
    }
    void bytes(const void* p, size_t n)
    {
        const auto* c = static_cast<const uint8_t*>(p);
        b.insert(b.end(), c, c + n);
    }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad8()
    {
        while (b.size() % 8)
            b.push_back(0);
    }
};

// ------------------------------------------------------------------ message encoders
Buf datatypeMessage(Type t)
{
    Buf m;
    auto fixed = [&](uint32_t size, bool sign) {
        m.u8(0x10 | 0); // version 1, class 0
        m.u8(sign ? 0x08 : 0x00);
        m.u8(0);
        m.u8(0);
        m.u32(size);
        m.u16(0);
        m.u16(static_cast<uint16_t>(8 * size));
    };
    switch (t) {
    case Type::F64:
        m.u8(0x10 | 1);
        m.u8(0x20); // little endian, mantissa normalisation: implied leading 1
        m.u8(63);   // sign bit position
        m.u8(0);
        m.u32(8);
        m.u16(0);   // bit offset
        m.u16(64);  // precision
        m.u8(52);   // exponent location
        m.u8(11);   // exponent size
        m.u8(0);    // mantissa location
        m.u8(52);   // mantissa size
        m.u32(1023);
        break;
    case Type::F32:
        m.u8(0x10 | 1);
        m.u8(0x20);
        m.u8(31);
        m.u8(0);
        m.u32(4);
        m.u16(0);
        m.u16(32);
        m.u8(23);
        m.u8(8);
        m.u8(0);
        m.u8(23);
        m.u32(127);
        break;
    case Type::U64: fixed(8, false); break;
    case Type::I64: fixed(8, true); break;
    case Type::U32: fixed(4, false); break;
    case Type::I32: fixed(4, true); break;
    case Type::U16: fixed(2, false); break;
    case Type::I16: fixed(2, true); break;
    case Type::U8: fixed(1, false); break;
    case Type::I8: fixed(1, true); break;
    case Type::String: // variable-length string of C_S1 characters (H5::StrType(C_S1, H5T_VARIABLE))
        m.u8(0x10 | 9);
        m.u8(0x01); // type: string; padding: null terminated
        m.u8(0x00); // character set: ASCII
        m.u8(0);
        m.u32(16);  // {length u32, collection address u64, object index u32}
        m.u8(0x10 | 3); // base type: 1-byte string
        m.u8(0);
        m.u8(0);
        m.u8(0);
        m.u32(1);
        break;
    default: throw Err("h5mini: cannot write this datatype");
    }
    return m;
}

Buf dataspaceMessage(const std::vector<uint64_t>& dims)
{
    Buf m;
    m.u8(1); // version
    m.u8(static_cast<uint8_t>(dims.size()));
    m.u8(0); // no maximum dimensions
    m.u8(0);
    m.u32(0);
    for (uint64_t d : dims)
        m.u64(d);
    return m;
}

struct Msg {
    uint16_t type;
    Buf body;
};

// version-1 object header with the given messages; returns its address
uint64_t writeObjectHeader(Out& out, std::vector<Msg>& msgs)
{
    Buf all;
    for (Msg& m : msgs) {
        m.body.pad8();
        if (m.body.b.size() > 0xFFF8)
            throw Err("h5mini: header message too large");
        all.u16(m.type);
        all.u16(static_cast<uint16_t>(m.body.b.size()));
        all.u8(0);
        all.zeros(3);
        all.bytes(m.body.b.data(), m.body.b.size());
    }
    Buf h;
    h.u8(1);
    h.u8(0);
    h.u16(static_cast<uint16_t>(msgs.size()));
    h.u32(1); // reference count
    h.u32(static_cast<uint32_t>(all.b.size()));
    h.u32(0); // alignment of the first message to 8 bytes
    h.bytes(all.b.data(), all.b.size());
    const uint64_t at = out.reserve(h.b.size());
    out.put(at, h.b.data(), h.b.size());
    return at;
}

Msg attributeMessage(const std::string& name, const Attribute& a, Out& out);

// global heap collection holding `strings`; returns per string {collection address, index}
std::vector<std::pair<uint64_t, uint32_t>> writeGlobalHeap(Out& out, const std::vector<std::string>& strings)
{
    std::vector<std::pair<uint64_t, uint32_t>> refs;
    size_t i = 0;
    while (i < strings.size()) {
        // one collection: 16-byte header + objects; at least 4096 bytes; object indices are 16 bit
        Buf c;
        size_t used = 16;
        size_t first = i;
        uint16_t idx = 1;
        while (i < strings.size() && idx < 0xFFFF) {
            const size_t need = 16 + ((strings[i].size() + 7) / 8) * 8;
            if (i > first && used + need + 16 > (1u << 24))
                break;
            used += need;
            ++i;
            ++idx;
        }
        const size_t total = std::max<size_t>(4096, ((used + 16 + 7) / 8) * 8);
        c.bytes("GCOL", 4);
        c.u8(1);
        c.zeros(3);
        c.u64(total);
        idx = 1;
        const uint64_t at = out.reserve(total);
        for (size_t k = first; k < i; ++k, ++idx) {
            c.u16(idx);
            c.u16(0);
            c.u32(0);
            c.u64(strings[k].size());
            c.bytes(strings[k].data(), strings[k].size());
            c.pad8();
            refs.emplace_back(at, idx);
        }
        // free space: object 0 spans the rest of the collection (its size field includes its own header)
        const size_t rest = total - c.b.size();
        if (rest >= 16) {
            c.u16(0);
            c.u16(0);
            c.u32(0);
            c.u64(rest);
        }
        out.put(at, c.b.data(), c.b.size());
    }
    return refs;
}

std::vector<uint8_t> vlenElements(Out& out, const std::vector<std::string>& strings)
{
    const auto refs = writeGlobalHeap(out, strings);
    Buf d;
    for (size_t k = 0; k < strings.size(); ++k) {
        d.u32(static_cast<uint32_t>(strings[k].size()));
        d.u64(refs[k].first);
        d.u32(refs[k].second);
    }
    return d.b;
}

Msg attributeMessage(const std::string& name, const Attribute& a, Out& out)
{
    Buf dt = datatypeMessage(a.type), ds = dataspaceMessage(a.dims);
    Msg m;
    m.type = 0x000C;
    Buf& b = m.body;
    b.u8(1);
    b.u8(0);
    b.u16(static_cast<uint16_t>(name.size() + 1));
    b.u16(static_cast<uint16_t>(dt.b.size()));
    b.u16(static_cast<uint16_t>(ds.b.size()));
    b.bytes(name.c_str(), name.size() + 1);
    b.pad8();
    b.bytes(dt.b.data(), dt.b.size());
    b.pad8();
    b.bytes(ds.b.data(), ds.b.size());
    b.pad8();
    if (a.type == Type::String) {
        const auto raw = vlenElements(out, a.strings);
        b.bytes(raw.data(), raw.size());
    } else {
        if (a.data.size() != a.count() * typeSize(a.type))
            throw Err("h5mini: attribute '" + name + "' data size does not match its extent");
        b.bytes(a.data.data(), a.data.size());
    }
    return m;
}

uint64_t writeDataset(Out& out, const std::string& name, const Dataset& d)
{
    std::vector<uint8_t> raw;
    const std::vector<uint8_t>* data = &d.data;
    if (d.type == Type::String) {
        if (d.strings.size() != d.count())
            throw Err("h5mini: string dataset '" + name + "' extent does not match");
        raw = vlenElements(out, d.strings);
        data = &raw;
    } else if (d.data.size() != d.count() * typeSize(d.type)) {
        throw Err("h5mini: dataset '" + name + "' data size does not match its extent");
    }
    const size_t esize = d.type == Type::String ? 16 : typeSize(d.type);
    std::vector<Msg> msgs;
    msgs.push_back({ 0x0001, dataspaceMessage(d.dims) });
    msgs.push_back({ 0x0003, datatypeMessage(d.type) });
    {
        Msg f;
        f.type = 0x0005; // fill value, version 2: early allocation, written if set, default value
        f.body.u8(2);
        f.body.u8(d.deflate ? 3 : 1);
        f.body.u8(2);
        f.body.u8(1);
        f.body.u32(0);
        msgs.push_back(std::move(f));
    }
    Msg layout;
    layout.type = 0x0008;
    layout.body.u8(3);
    const bool chunked = d.deflate && d.type != Type::String && !d.dims.empty() && !data->empty();
    if (!chunked) {
        uint64_t at = kUndef;
        if (!data->empty()) {
            at = out.reserve(data->size());
            out.put(at, data->data(), data->size());
        }
        layout.body.u8(1);
        layout.body.u64(at);
        layout.body.u64(data->size());
    } else {
        // ONE chunk spanning the extent, deflate level 6 (setChunk(rank, dims) + setDeflate(6), R:...hdf5wrapper.cpp:147-150)
        uLongf bound = compressBound(static_cast<uLong>(data->size()));
        std::vector<uint8_t> z(bound);
        if (compress2(z.data(), &bound, data->data(), static_cast<uLong>(data->size()), 6) != Z_OK)
            throw Err("h5mini: deflate failed");
        if (bound > 0xFFFFFFFFull)
            throw Err("h5mini: chunk larger than 4 GiB");
        const uint64_t chunkAt = out.reserve(bound);
        out.put(chunkAt, z.data(), bound);
        const size_t nd = d.dims.size() + 1;
        const size_t keySize = 8 + 8 * nd;
        Buf t;
        t.bytes("TREE", 4);
        t.u8(1); // node type: raw data chunks
        t.u8(0); // leaf
        t.u16(1);
        t.u64(kUndef);
        t.u64(kUndef);
        t.u32(static_cast<uint32_t>(bound));
        t.u32(0);
        for (size_t k = 0; k < nd; ++k)
            t.u64(0);
        t.u64(chunkAt);
        t.u32(0); // the key after the last child: the next chunk position along the slowest dimension
        t.u32(0);
        for (size_t k = 0; k < nd; ++k)
            t.u64(k == 0 ? d.dims[0] : 0);
        const size_t nodeSize = 24 + (2 * kChunkK + 1) * keySize + 2 * kChunkK * 8;
        const uint64_t treeAt = out.reserve(nodeSize);
        out.put(treeAt, t.b.data(), t.b.size());
        layout.body.u8(2);
        layout.body.u8(static_cast<uint8_t>(nd));
        layout.body.u64(treeAt);
        for (uint64_t dim : d.dims) {
            if (dim > 0xFFFFFFFFull)
                throw Err("h5mini: chunk dimension too large");
            layout.body.u32(static_cast<uint32_t>(dim));
        }
        layout.body.u32(static_cast<uint32_t>(esize));
        Msg filt;
        filt.type = 0x000B;
        filt.body.u8(1);
        filt.body.u8(1);
        filt.body.u16(0);
        filt.body.u32(0);
        filt.body.u16(1); // deflate
        filt.body.u16(0); // no name
        filt.body.u16(1); // optional
        filt.body.u16(1); // one client value
        filt.body.u32(6);
        filt.body.u32(0); // padding: odd number of values
        msgs.push_back(std::move(filt));
    }
    msgs.push_back(std::move(layout));
    return writeObjectHeader(out, msgs);
}

struct GroupAddr {
    uint64_t header, btree, heap;
};

GroupAddr writeGroup(Out& out, const Group& g)
{
    // children first: (name, object header address, cached B-tree / heap for groups)
    struct Child {
        std::string name;
        uint64_t header;
        bool isGroup;
        uint64_t btree, heap;
    };
    std::vector<Child> children;
    for (const auto& [name, sub] : g.groups) {
        const GroupAddr a = writeGroup(out, *sub);
        children.push_back({ name, a.header, true, a.btree, a.heap });
    }
    for (const auto& [name, ds] : g.datasets) {
        if (g.groups.count(name))
            throw Err("h5mini: '" + name + "' is both a group and a dataset");
        children.push_back({ name, writeDataset(out, name, ds), false, 0, 0 });
    }
    std::sort(children.begin(), children.end(), [](const Child& a, const Child& b) { return a.name < b.name; });
    if (children.size() > static_cast<size_t>(2 * kGroupLeafK) * 2 * kGroupInternalK)
        throw Err("h5mini: too many links in one group");
    // local heap: offset 0 = "" (the name of the first B-tree key), then the link names
    Buf heap;
    heap.zeros(8);
    std::vector<uint64_t> nameOff(children.size());
    for (size_t i = 0; i < children.size(); ++i) {
        nameOff[i] = heap.b.size();
        heap.bytes(children[i].name.c_str(), children[i].name.size() + 1);
        heap.pad8();
    }
    // a free block fills the data segment to its allocated size: {next free offset (1 = none), size of the block}
    const size_t freeAt = heap.b.size();
    const size_t segSize = std::max<size_t>(freeAt + 16, 88);
    heap.u64(1);
    heap.u64(segSize - freeAt);
    heap.zeros(segSize - heap.b.size());
    const uint64_t segAt = out.reserve(segSize);
    out.put(segAt, heap.b.data(), heap.b.size());
    Buf hh;
    hh.bytes("HEAP", 4);
    hh.u8(0);
    hh.zeros(3);
    hh.u64(segSize);
    hh.u64(freeAt);
    hh.u64(segAt);
    const uint64_t heapAt = out.reserve(hh.b.size());
    out.put(heapAt, hh.b.data(), hh.b.size());
    // symbol nodes of up to 2K entries each
    const size_t perNode = 2 * kGroupLeafK;
    const size_t nNodes = std::max<size_t>(1, (children.size() + perNode - 1) / perNode);
    std::vector<uint64_t> nodeAt(nNodes), lastName(nNodes, 0);
    for (size_t n = 0; n < nNodes; ++n) {
        const size_t b = n * perNode, e = std::min(children.size(), b + perNode);
        Buf s;
        s.bytes("SNOD", 4);
        s.u8(1);
        s.u8(0);
        s.u16(static_cast<uint16_t>(e - b));
        for (size_t i = b; i < e; ++i) {
            s.u64(nameOff[i]);
            s.u64(children[i].header);
            s.u32(children[i].isGroup ? 1 : 0); // cache type 1: the group's B-tree and heap addresses follow
            s.u32(0);
            s.u64(children[i].isGroup ? children[i].btree : 0);
            s.u64(children[i].isGroup ? children[i].heap : 0);
            lastName[n] = nameOff[i];
        }
        nodeAt[n] = out.reserve(8 + perNode * 40);
        out.put(nodeAt[n], s.b.data(), s.b.size());
    }
    Buf t;
    t.bytes("TREE", 4);
    t.u8(0); // group node
    t.u8(0); // level 0: children are symbol nodes
    t.u16(static_cast<uint16_t>(children.empty() ? 0 : nNodes));
    t.u64(kUndef);
    t.u64(kUndef);
    t.u64(0); // key 0: the empty string
    if (!children.empty())
        for (size_t n = 0; n < nNodes; ++n) {
            t.u64(nodeAt[n]);
            t.u64(lastName[n]); // key n + 1: the largest name in child n
        }
    const uint64_t treeAt = out.reserve(24 + (2 * kGroupInternalK + 1) * 8 + 2 * kGroupInternalK * 8);
    out.put(treeAt, t.b.data(), t.b.size());
    std::vector<Msg> msgs;
    {
        Msg st;
        st.type = 0x0011;
        st.body.u64(treeAt);
        st.body.u64(heapAt);
        msgs.push_back(std::move(st));
    }
    std::vector<std::string> order = g.attributeOrder;
    for (const auto& kv : g.attributes)
        if (std::find(order.begin(), order.end(), kv.first) == order.end())
            order.push_back(kv.first);
    for (const std::string& name : order) {
        auto it = g.attributes.find(name);
        if (it != g.attributes.end())
            msgs.push_back(attributeMessage(name, it->second, out));
    }
    return { writeObjectHeader(out, msgs), treeAt, heapAt };
}

// ------------------------------------------------------------------ reader
struct In {
    const std::vector<uint8_t>& b;
    uint64_t base = 0; // superblock base address: every file address is relative to it
    int so = 8, sl = 8;
    explicit In(const std::vector<uint8_t>& bytes)
        : b(bytes)
    {
    }
    void need(uint64_t at, uint64_t n) const
    {
        if (at > b.size() || n > b.size() - at)
            throw Err("h5mini: structure runs past the end of the file");
    }
    uint64_t le(uint64_t at, int n) const
    {
        need(at, n);
        uint64_t v = 0;
        for (int i = 0; i < n; ++i)
            v |= static_cast<uint64_t>(b[at + i]) << (8 * i);
        return v;
    }
    uint64_t off(uint64_t at) const
    {
        const uint64_t v = le(at, so);
        return (so == 8 ? v == kUndef : v == ((1ull << (8 * so)) - 1)) ? kUndef : v + base;
    }
    uint64_t len(uint64_t at) const { return le(at, sl); }
    bool tag(uint64_t at, const char* t) const
    {
        need(at, 4);
        return std::memcmp(b.data() + at, t, 4) == 0;
    }
};

struct TypeInfo {
    Type type = Type::Unknown;
    size_t size = 0;     // element size in the file
    bool vlenString = false, fixedString = false;
};

TypeInfo decodeDatatype(const In& in, uint64_t at)
{
    TypeInfo t;
    const int cls = static_cast<int>(in.le(at, 1)) & 0x0F;
    const uint32_t bits = static_cast<uint32_t>(in.le(at + 1, 3));
    t.size = static_cast<size_t>(in.le(at + 4, 4));
    if (cls == 0) {
        const bool sign = (bits & 0x08) != 0;
        if (bits & 0x01)
            return t; // big endian: not produced on the platforms OpenDXMC runs on
        switch (t.size) {
        case 1: t.type = sign ? Type::I8 : Type::U8; break;
        case 2: t.type = sign ? Type::I16 : Type::U16; break;
        case 4: t.type = sign ? Type::I32 : Type::U32; break;
        case 8: t.type = sign ? Type::I64 : Type::U64; break;
        default: break;
        }
    } else if (cls == 1) {
        if (bits & 0x01)
            return t;
        if (t.size == 8)
            t.type = Type::F64;
        else if (t.size == 4)
            t.type = Type::F32;
    } else if (cls == 3) {
        t.type = Type::String;
        t.fixedString = true;
    } else if (cls == 9) {
        if ((bits & 0x0F) == 1) {
            t.type = Type::String;
            t.vlenString = true;
        }
    }
    return t;
}

std::vector<uint64_t> decodeDataspace(const In& in, uint64_t at)
{
    const int version = static_cast<int>(in.le(at, 1));
    const int rank = static_cast<int>(in.le(at + 1, 1));
    uint64_t p;
    if (version == 1)
        p = at + 8;
    else if (version == 2)
        p = at + 4;
    else
        throw Err("h5mini: unsupported dataspace message version");
    std::vector<uint64_t> dims(rank);
    for (int i = 0; i < rank; ++i)
        dims[i] = in.len(p + static_cast<uint64_t>(i) * in.sl);
    return dims;
}

std::string globalHeapObject(const In& in, uint64_t collection, uint32_t index)
{
    if (collection == kUndef || !in.tag(collection, "GCOL"))
        throw Err("h5mini: bad global heap reference");
    const uint64_t total = in.len(collection + 8);
    uint64_t p = collection + 8 + in.sl;
    const uint64_t end = collection + total;
    while (p + 8 + in.sl <= end) {
        const uint32_t idx = static_cast<uint32_t>(in.le(p, 2));
        const uint64_t size = in.len(p + 8);
        if (idx == 0)
            break;
        if (idx == index) {
            in.need(p + 8 + in.sl, size);
            return std::string(reinterpret_cast<const char*>(in.b.data() + p + 8 + in.sl), size);
        }
        p += 8 + in.sl + ((size + 7) / 8) * 8;
    }
    throw Err("h5mini: global heap object not found");
}

void decodeElements(const In& in, const TypeInfo& t, const uint8_t* raw, uint64_t count, std::vector<uint8_t>& data, std::vector<std::string>& strings)
{
    if (t.type == Type::String) {
        strings.resize(count);
        for (uint64_t i = 0; i < count; ++i) {
            const uint8_t* e = raw + i * t.size;
            if (t.fixedString) {
                size_t n = 0;
                while (n < t.size && e[n])
                    ++n;
                strings[i].assign(reinterpret_cast<const char*>(e), n);
            } else {
                uint64_t lenv = 0, addr = 0, idx = 0;
                for (int k = 0; k < 4; ++k)
                    lenv |= static_cast<uint64_t>(e[k]) << (8 * k);
                for (int k = 0; k < in.so; ++k)
                    addr |= static_cast<uint64_t>(e[4 + k]) << (8 * k);
                for (int k = 0; k < 4; ++k)
                    idx |= static_cast<uint64_t>(e[4 + in.so + k]) << (8 * k);
                if (lenv == 0 && addr == 0) {
                    strings[i].clear();
                    continue;
                }
                std::string s = globalHeapObject(in, addr + in.base, static_cast<uint32_t>(idx));
                if (s.size() > lenv)
                    s.resize(lenv);
                strings[i] = s;
            }
        }
    } else {
        data.assign(raw, raw + count * t.size);
    }
}

struct Filter {
    int id;
    std::vector<uint32_t> values;
};

std::vector<uint8_t> unfilter(std::vector<uint8_t> chunk, const std::vector<Filter>& pipeline, uint32_t mask, size_t expected, size_t esize)
{
    for (int f = static_cast<int>(pipeline.size()) - 1; f >= 0; --f) {
        if (mask & (1u << f))
            continue;
        const Filter& ft = pipeline[f];
        if (ft.id == 1) { // deflate
            std::vector<uint8_t> o(std::max<size_t>(expected, 64));
            for (;;) {
                uLongf n = static_cast<uLongf>(o.size());
                const int rc = uncompress(o.data(), &n, chunk.data(), static_cast<uLong>(chunk.size()));
                if (rc == Z_OK) {
                    o.resize(n);
                    break;
                }
                if (rc != Z_BUF_ERROR)
                    throw Err("h5mini: inflate failed");
                o.resize(o.size() * 2);
            }
            chunk.swap(o);
        } else if (ft.id == 2) { // shuffle: bytes of equal significance were stored together
            const size_t es = ft.values.empty() ? esize : ft.values[0];
            if (es > 1 && chunk.size() >= es) {
                const size_t n = chunk.size() / es;
                std::vector<uint8_t> o(chunk.size());
                for (size_t k = 0; k < es; ++k)
                    for (size_t i = 0; i < n; ++i)
                        o[i * es + k] = chunk[k * n + i];
                for (size_t i = n * es; i < chunk.size(); ++i)
                    o[i] = chunk[i];
                chunk.swap(o);
            }
        } else if (ft.id == 3) { // fletcher32 checksum appended
            if (chunk.size() >= 4)
                chunk.resize(chunk.size() - 4);
        } else {
            throw Err("h5mini: unsupported filter " + std::to_string(ft.id));
        }
    }
    return chunk;
}

struct Message {
    int type;
    uint64_t at, size;
};

std::vector<Message> readMessages(const In& in, uint64_t header)
{
    std::vector<Message> msgs;
    if (in.le(header, 1) != 1)
        throw Err("h5mini: only version-1 object headers are supported");
    const int total = static_cast<int>(in.le(header + 2, 2));
    std::vector<std::pair<uint64_t, uint64_t>> blocks { { header + 16, in.le(header + 8, 4) } };
    for (size_t b = 0; b < blocks.size() && static_cast<int>(msgs.size()) < total; ++b) {
        uint64_t p = blocks[b].first;
        const uint64_t end = p + blocks[b].second;
        in.need(blocks[b].first, blocks[b].second);
        while (p + 8 <= end && static_cast<int>(msgs.size()) < total) {
            const int type = static_cast<int>(in.le(p, 2));
            const uint64_t size = in.le(p + 2, 2);
            if (p + 8 + size > end)
                throw Err("h5mini: header message runs past its block");
            msgs.push_back({ type, p + 8, size });
            if (type == 0x0010)
                blocks.emplace_back(in.off(p + 8), in.len(p + 8 + in.so));
            p += 8 + size;
        }
    }
    return msgs;
}

Attribute readAttribute(const In& in, const Message& m, std::string& name)
{
    const int version = static_cast<int>(in.le(m.at, 1));
    const uint64_t nameSize = in.le(m.at + 2, 2), dtSize = in.le(m.at + 4, 2), dsSize = in.le(m.at + 6, 2);
    uint64_t p = m.at + 8;
    if (version == 3)
        ++p; // name character set
    auto padded = [&](uint64_t n) { return version == 1 ? ((n + 7) / 8) * 8 : n; };
    in.need(p, nameSize);
    name.assign(reinterpret_cast<const char*>(in.b.data() + p), nameSize ? nameSize - 1 : 0);
    name = std::string(name.c_str());
    p += padded(nameSize);
    const TypeInfo t = decodeDatatype(in, p);
    p += padded(dtSize);
    Attribute a;
    a.dims = decodeDataspace(in, p);
    p += padded(dsSize);
    a.type = t.type;
    const uint64_t count = a.count();
    if (t.type != Type::Unknown) {
        in.need(p, count * t.size);
        decodeElements(in, t, in.b.data() + p, count, a.data, a.strings);
    }
    return a;
}

void readChunkTree(const In& in, uint64_t node, size_t nd, std::vector<std::pair<std::vector<uint64_t>, std::pair<uint64_t, std::pair<uint32_t, uint32_t>>>>& chunks)
{
    if (node == kUndef)
        return;
    if (!in.tag(node, "TREE") || in.le(node + 4, 1) != 1)
        throw Err("h5mini: bad chunk B-tree node");
    const int level = static_cast<int>(in.le(node + 5, 1));
    const int used = static_cast<int>(in.le(node + 6, 2));
    const uint64_t keySize = 8 + 8 * nd;
    uint64_t p = node + 8 + 2 * in.so;
    for (int i = 0; i < used; ++i) {
        const uint32_t size = static_cast<uint32_t>(in.le(p, 4)), mask = static_cast<uint32_t>(in.le(p + 4, 4));
        std::vector<uint64_t> offs(nd);
        for (size_t k = 0; k < nd; ++k)
            offs[k] = in.le(p + 8 + 8 * k, 8);
        const uint64_t child = in.off(p + keySize);
        if (level > 0)
            readChunkTree(in, child, nd, chunks);
        else
            chunks.push_back({ offs, { child, { size, mask } } });
        p += keySize + in.so;
    }
}

Dataset readDataset(const In& in, const std::vector<Message>& msgs)
{
    Dataset d;
    TypeInfo t;
    std::vector<Filter> pipeline;
    const Message* layout = nullptr;
    for (const Message& m : msgs) {
        if (m.type == 0x0001)
            d.dims = decodeDataspace(in, m.at);
        else if (m.type == 0x0003)
            t = decodeDatatype(in, m.at);
        else if (m.type == 0x0008)
            layout = &m;
        else if (m.type == 0x000B) {
            const int version = static_cast<int>(in.le(m.at, 1));
            const int n = static_cast<int>(in.le(m.at + 1, 1));
            uint64_t p = m.at + (version == 1 ? 8 : 2);
            for (int i = 0; i < n; ++i) {
                Filter f;
                f.id = static_cast<int>(in.le(p, 2));
                uint64_t nameLen = 0;
                if (version == 1 || f.id >= 256) {
                    nameLen = in.le(p + 2, 2);
                    p += 2;
                }
                const int nv = static_cast<int>(in.le(p + 4, 2));
                p += 6;
                p += version == 1 ? ((nameLen + 7) / 8) * 8 : nameLen;
                for (int k = 0; k < nv; ++k)
                    f.values.push_back(static_cast<uint32_t>(in.le(p + 4 * k, 4)));
                p += 4 * nv;
                if (version == 1 && (nv & 1))
                    p += 4;
                pipeline.push_back(f);
            }
        }
    }
    d.type = t.type;
    d.deflate = !pipeline.empty();
    if (!layout || t.type == Type::Unknown)
        return d;
    const uint64_t count = d.count();
    const uint64_t bytes = count * t.size;
    std::vector<uint8_t> raw(bytes, 0);
    const int version = static_cast<int>(in.le(layout->at, 1));
    if (version < 1 || version > 3)
        throw Err("h5mini: unsupported data layout message version " + std::to_string(version));
    // versions 1 and 2 (HDF5 1.6 and earlier): dimensionality, class, 5 reserved bytes, [address], 4-byte dimension sizes,
    // [compact: size + data]; version 3: class, then class-specific fields
    int cls;
    uint64_t compactAt = 0, compactSize = 0, address = kUndef, chunkDimsAt = 0;
    size_t nd = 0;
    if (version == 3) {
        cls = static_cast<int>(in.le(layout->at + 1, 1));
        if (cls == 0) {
            compactSize = in.le(layout->at + 2, 2);
            compactAt = layout->at + 4;
        } else if (cls == 1) {
            address = in.off(layout->at + 2);
        } else if (cls == 2) {
            nd = static_cast<size_t>(in.le(layout->at + 2, 1));
            address = in.off(layout->at + 3);
            chunkDimsAt = layout->at + 3 + in.so;
        }
    } else {
        nd = static_cast<size_t>(in.le(layout->at + 1, 1));
        cls = static_cast<int>(in.le(layout->at + 2, 1));
        uint64_t p = layout->at + 8;
        if (cls != 0) {
            address = in.off(p);
            p += in.so;
        }
        chunkDimsAt = p;
        p += 4 * nd;
        if (cls == 0) {
            compactSize = in.le(p, 4);
            compactAt = p + 4;
        }
    }
    if (cls == 0) { // compact
        in.need(compactAt, compactSize);
        std::memcpy(raw.data(), in.b.data() + compactAt, std::min<uint64_t>(compactSize, bytes));
    } else if (cls == 1) { // contiguous
        const uint64_t at = address;
        if (at != kUndef && bytes) {
            in.need(at, bytes);
            std::memcpy(raw.data(), in.b.data() + at, bytes);
        }
    } else if (cls == 2) { // chunked
        const uint64_t tree = address;
        if (nd != d.dims.size() + 1)
            throw Err("h5mini: chunk dimensionality does not match the dataspace");
        std::vector<uint64_t> cdim(nd);
        for (size_t k = 0; k < nd; ++k)
            cdim[k] = in.le(chunkDimsAt + 4 * k, 4);
        uint64_t chunkBytes = 1;
        for (uint64_t c : cdim)
            chunkBytes *= c;
        std::vector<std::pair<std::vector<uint64_t>, std::pair<uint64_t, std::pair<uint32_t, uint32_t>>>> chunks;
        readChunkTree(in, tree, nd, chunks);
        const size_t rank = d.dims.size();
        for (const auto& ch : chunks) {
            in.need(ch.second.first, ch.second.second.first);
            std::vector<uint8_t> c(in.b.begin() + ch.second.first, in.b.begin() + ch.second.first + ch.second.second.first);
            if (!pipeline.empty())
                c = unfilter(std::move(c), pipeline, ch.second.second.second, chunkBytes, t.size);
            if (c.size() < chunkBytes)
                c.resize(chunkBytes, 0);
            // copy the part of the chunk that lies inside the extent, row by row along the fastest dimension
            std::vector<uint64_t> idx(rank, 0);
            const uint64_t rowElems = rank ? std::min<uint64_t>(cdim[rank - 1], d.dims[rank - 1] > ch.first[rank - 1] ? d.dims[rank - 1] - ch.first[rank - 1] : 0) : 1;
            if (rowElems == 0)
                continue;
            for (;;) {
                bool inside = true;
                uint64_t dst = 0, src = 0;
                for (size_t k = 0; k < rank; ++k) {
                    const uint64_t g = ch.first[k] + idx[k];
                    if (g >= d.dims[k])
                        inside = false;
                    dst = dst * d.dims[k] + g;
                    src = src * cdim[k] + idx[k];
                }
                if (inside)
                    std::memcpy(raw.data() + dst * t.size, c.data() + src * t.size, rowElems * t.size);
                // next row: advance all but the fastest dimension
                int k = static_cast<int>(rank) - 2;
                for (; k >= 0; --k) {
                    if (++idx[k] < cdim[k])
                        break;
                    idx[k] = 0;
                }
                if (k < 0)
                    break;
            }
        }
    } else {
        throw Err("h5mini: unsupported layout class");
    }
    decodeElements(in, t, raw.data(), count, d.data, d.strings);
    return d;
}

void readGroupInto(const In& in, uint64_t header, Group& g, int depth);

void readSymbolTree(const In& in, uint64_t node, uint64_t heapData, Group& g, int depth)
{
    if (node == kUndef)
        return;
    if (!in.tag(node, "TREE") || in.le(node + 4, 1) != 0)
        throw Err("h5mini: bad group B-tree node");
    const int level = static_cast<int>(in.le(node + 5, 1));
    const int used = static_cast<int>(in.le(node + 6, 2));
    uint64_t p = node + 8 + 2 * in.so + in.sl; // past key 0
    for (int i = 0; i < used; ++i) {
        const uint64_t child = in.off(p);
        p += in.so + in.sl;
        if (level > 0) {
            readSymbolTree(in, child, heapData, g, depth);
            continue;
        }
        if (!in.tag(child, "SNOD"))
            throw Err("h5mini: bad symbol node");
        const int n = static_cast<int>(in.le(child + 6, 2));
        const uint64_t entry = 2 * in.so + 8 + 16;
        for (int k = 0; k < n; ++k) {
            const uint64_t e = child + 8 + k * entry;
            const uint64_t nameAt = heapData + in.le(e, in.so);
            in.need(nameAt, 1);
            const std::string name(reinterpret_cast<const char*>(in.b.data() + nameAt));
            const uint64_t objHeader = in.off(e + in.so);
            const uint32_t cache = static_cast<uint32_t>(in.le(e + 2 * in.so, 4));
            if (cache == 2 || objHeader == kUndef)
                continue; // symbolic link
            const std::vector<Message> msgs = readMessages(in, objHeader);
            const bool isGroup = std::any_of(msgs.begin(), msgs.end(), [](const Message& m) { return m.type == 0x0011; });
            if (isGroup) {
                auto sub = std::make_unique<Group>();
                readGroupInto(in, objHeader, *sub, depth + 1);
                g.groups[name] = std::move(sub);
            } else {
                g.datasets[name] = readDataset(in, msgs);
            }
        }
    }
}

void readGroupInto(const In& in, uint64_t header, Group& g, int depth)
{
    if (depth > 64)
        throw Err("h5mini: groups nested too deeply");
    const std::vector<Message> msgs = readMessages(in, header);
    for (const Message& m : msgs) {
        if (m.type == 0x000C) {
            std::string name;
            Attribute a = readAttribute(in, m, name);
            g.attributeOrder.push_back(name);
            g.attributes[name] = std::move(a);
        }
    }
    for (const Message& m : msgs) {
        if (m.type != 0x0011)
            continue;
        const uint64_t tree = in.off(m.at), heap = in.off(m.at + in.so);
        if (heap == kUndef || !in.tag(heap, "HEAP"))
            throw Err("h5mini: bad local heap");
        const uint64_t heapData = in.off(heap + 8 + 2 * in.sl);
        readSymbolTree(in, tree, heapData, g, depth);
    }
}

} // namespace

size_t typeSize(Type t)
{
    switch (t) {
    case Type::F64:
    case Type::U64:
    case Type::I64: return 8;
    case Type::F32:
    case Type::U32:
    case Type::I32: return 4;
    case Type::U16:
    case Type::I16: return 2;
    case Type::U8:
    case Type::I8: return 1;
    default: return 0;
    }
}

uint64_t Attribute::count() const
{
    uint64_t n = 1;
    for (uint64_t d : dims)
        n *= d;
    return n;
}
uint64_t Dataset::count() const
{
    uint64_t n = 1;
    for (uint64_t d : dims)
        n *= d;
    return n;
}

std::vector<std::string> File::split(const std::string& path)
{
    std::vector<std::string> parts;
    std::string cur;
    for (char ch : path) {
        if (ch == '/') {
            if (!cur.empty())
                parts.push_back(cur);
            cur.clear();
        } else {
            cur.push_back(ch);
        }
    }
    if (!cur.empty())
        parts.push_back(cur);
    return parts;
}

Group* File::group(const std::string& path, bool create)
{
    Group* g = &root;
    for (const std::string& name : split(path)) {
        auto it = g->groups.find(name);
        if (it == g->groups.end()) {
            if (!create || g->datasets.count(name))
                return nullptr;
            it = g->groups.emplace(name, std::make_unique<Group>()).first;
        }
        g = it->second.get();
    }
    return g;
}
const Group* File::group(const std::string& path) const { return const_cast<File*>(this)->group(path, false); }

Dataset* File::dataset(const std::string& path)
{
    std::vector<std::string> parts = split(path);
    if (parts.empty())
        return nullptr;
    const std::string leaf = parts.back();
    parts.pop_back();
    Group* g = &root;
    for (const std::string& name : parts) {
        auto it = g->groups.find(name);
        if (it == g->groups.end())
            return nullptr;
        g = it->second.get();
    }
    auto it = g->datasets.find(leaf);
    return it == g->datasets.end() ? nullptr : &it->second;
}
const Dataset* File::dataset(const std::string& path) const { return const_cast<File*>(this)->dataset(path); }

bool File::exists(const std::string& path) const { return group(path) != nullptr || dataset(path) != nullptr; }

Dataset& File::createDataset(const std::string& path)
{
    std::vector<std::string> parts = split(path);
    if (parts.empty())
        throw Err("h5mini: empty dataset path");
    const std::string leaf = parts.back();
    parts.pop_back();
    Group* g = &root;
    for (const std::string& name : parts) {
        auto it = g->groups.find(name);
        if (it == g->groups.end())
            it = g->groups.emplace(name, std::make_unique<Group>()).first;
        g = it->second.get();
    }
    return g->datasets[leaf];
}

bool File::save(const std::string& path, std::string* error) const
{
    try {
        Out out;
        out.b.resize(96, 0); // superblock, filled in last
        const GroupAddr r = writeGroup(out, root);
        out.align();
        Buf s;
        s.bytes(kSignature, 8);
        s.u8(0); // superblock version
        s.u8(0); // free-space storage version
        s.u8(0); // root group symbol table entry version
        s.u8(0);
        s.u8(0); // shared header message format version
        s.u8(8); // size of offsets
        s.u8(8); // size of lengths
        s.u8(0);
        s.u16(kGroupLeafK);
        s.u16(kGroupInternalK);
        s.u32(0); // file consistency flags
        s.u64(0); // base address
        s.u64(kUndef); // free-space info
        s.u64(out.size()); // end of file
        s.u64(kUndef); // driver information
        s.u64(0);        // root entry: link name offset
        s.u64(r.header); // object header
        s.u32(1);        // cache type: group
        s.u32(0);
        s.u64(r.btree);
        s.u64(r.heap);
        out.put(0, s.b.data(), s.b.size());
        std::ofstream f(path, std::ios::binary | std::ios::trunc);
        if (!f)
            throw Err("h5mini: cannot open '" + path + "' for writing");
        f.write(reinterpret_cast<const char*>(out.b.data()), static_cast<std::streamsize>(out.b.size()));
        if (!f)
            throw Err("h5mini: write failed");
        return true;
    } catch (const std::exception& e) {
        if (error)
            *error = e.what();
        return false;
    }
}

std::unique_ptr<File> File::load(const std::string& path, std::string* error)
{
    try {
        std::ifstream f(path, std::ios::binary);
        if (!f)
            throw Err("h5mini: cannot open '" + path + "'");
        std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        In in(bytes);
        // the superblock sits at 0 or at 512, 1024, 2048, ... (behind a user block)
        uint64_t sb = kUndef;
        for (uint64_t at = 0; at + 8 <= bytes.size(); at = at ? at * 2 : 512) {
            if (std::memcmp(bytes.data() + at, kSignature, 8) == 0) {
                sb = at;
                break;
            }
        }
        if (sb == kUndef)
            throw Err("h5mini: not an HDF5 file");
        const int version = static_cast<int>(in.le(sb + 8, 1));
        if (version > 1)
            throw Err("h5mini: superblock version " + std::to_string(version) + " is not supported (versions 0 and 1 are)");
        in.so = static_cast<int>(in.le(sb + 13, 1));
        in.sl = static_cast<int>(in.le(sb + 14, 1));
        if ((in.so != 8 && in.so != 4) || (in.sl != 8 && in.sl != 4))
            throw Err("h5mini: unsupported offset / length size");
        uint64_t p = sb + 24 + (version == 1 ? 4 : 0);
        const uint64_t baseField = in.le(p, in.so);
        in.base = baseField + (baseField == 0 ? sb : 0); // files with a user block store base address 0 relative to ... the superblock
        p += 4 * in.so; // base, free-space info, end of file, driver info
        // root group symbol table entry
        const uint64_t rootHeader = in.off(p + in.so);
        auto file = std::make_unique<File>();
        readGroupInto(in, rootHeader, file->root, 0);
        return file;
    } catch (const std::exception& e) {
        if (error)
            *error = e.what();
        return nullptr;
    }
}

} // namespace h5mini
