// icrp.cpp — ICRP 110 / 143 voxel-phantom import rules, host side.
//
// Restates what OpenDXMC's ICRPPhantomImportPipeline::importPhantom does with the phantom's three files
// (R:src/libopendxmc/icrpphantomimportpipeline.cpp:209-351): the organ table (`*_organs.dat`, :59-125) and the media
// table (`*_media.dat`, :133-205) are parsed line by line with the reference's own, rather particular, rules; "Air" is
// appended as organ 0 / medium 0 (rho 0.001, {N: 0.8, O: 0.2}, :225, :252); with remove_arms the organs whose name
// contains "arm", "hand", "Humeri" or "Ulnae" are turned into air (:226-246); organs that do not occur in the voxel
// array are dropped and the remaining ids made consecutive (:215-235 of pruneOrganArray), media no organ refers to are
// dropped and renumbered (:237-263); material and density arrays follow by organ lookup (:275-300).
//
// Everything the voxel array needs is folded into three 256-entry look-up tables (organ id -> new organ id, -> medium,
// -> density); the O(N) part - one gather per voxel over 7 - 55 M voxels - is the device kernel in import_kernels.cu.
#include "icrp.hpp"

#include <algorithm>
#include <charconv>
#include <cctype>

namespace dxb {

namespace {

struct OrganRow {
    double density = 0;
    uint8_t id = 0, medium = 0;
    std::string name;
};
struct MediaRow {
    uint8_t id = 0;
    std::vector<std::pair<uint32_t, double>> composition;
    std::string name;
};

std::string trimmed(const char* a, const char* b)
{
    while (a != b && std::isspace(static_cast<unsigned char>(*a)))
        ++a;
    while (a != b && std::isspace(static_cast<unsigned char>(*(b - 1))))
        --b;
    return std::string(a, b);
}

// one line of *_organs.dat: "<id> <name ... at least 50 characters wide> <tissue number> <density>"
bool organFromLine(const std::string& line, OrganRow& o)
{
    const char* p = line.data();
    const char* const end = p + line.size();
    auto r = std::from_chars(p, end, o.id);
    if (r.ec != std::errc {})
        return false; // header / blank lines do not start with a number
    const char* const nameBegin = r.ptr;
    // the name may contain digits ("Humeri, upper half"...): the tissue number is searched from 50 characters on
    if (end - r.ptr < 50)
        return false; // (the reference runs off such a line without finding a density)
    p = r.ptr + 50;
    const char* nameEnd = p;
    bool found = false;
    while (!found && p < end) {
        r = std::from_chars(p, end, o.medium);
        if (r.ec != std::errc {}) {
            ++p;
        } else {
            found = true;
            nameEnd = p;
            p = r.ptr;
        }
    }
    o.density = 0.0;
    while (p < end) {
        const auto d = std::from_chars(p++, end, o.density);
        if (d.ec == std::errc {})
            break;
    }
    if (o.density == 0.0)
        return false;
    if (nameEnd > end)
        return false;
    o.name = trimmed(nameBegin, nameEnd);
    return !o.name.empty();
}

// one line of *_media.dat: "<id> <name> <13 mass percentages of H C N O Na Mg P S Cl K Ca Fe I>"
bool mediumFromLine(const std::string& line, MediaRow& m)
{
    static const uint32_t kZ[13] = { 1, 6, 7, 8, 11, 12, 15, 16, 17, 19, 20, 26, 53 };
    const char* p = line.data();
    const char* const end = p + line.size();
    auto r = std::from_chars(p, end, m.id);
    if (r.ec != std::errc {})
        return false;
    const char* const nameBegin = r.ptr;
    p = r.ptr;
    while (p != end && !std::isdigit(static_cast<unsigned char>(*p)))
        ++p;
    const char* const nameEnd = p;
    if (p == end)
        return false;
    m.composition.clear();
    for (uint32_t z : kZ) {
        while (p != end && std::isspace(static_cast<unsigned char>(*p)))
            ++p;
        if (p == end)
            return false;
        double w = 0;
        const auto d = std::from_chars(p, end, w);
        if (d.ec != std::errc {})
            return false;
        p = d.ptr;
        m.composition.emplace_back(z, w);
    }
    m.name = trimmed(nameBegin, nameEnd);
    return true;
}

template <typename Row, typename Parser>
std::vector<Row> rowsOf(const char* text, Parser parse)
{
    std::vector<Row> rows;
    const char* p = text;
    while (*p) {
        const char* q = p;
        while (*q && *q != '\n')
            ++q;
        std::string line(p, q);
        Row row;
        if (parse(line, row))
            rows.push_back(std::move(row));
        p = *q ? q + 1 : q;
    }
    return rows;
}

} // namespace

// `present[v]` != 0: value v occurs in the organ array.  Returns false when a table has no valid line.
bool icrpPlan(const char* organsText, const char* mediaText, bool removeArms, const uint8_t present[256], IcrpPlan& plan)
{
    std::vector<OrganRow> organs = rowsOf<OrganRow>(organsText, organFromLine);
    if (organs.empty())
        return false;
    std::vector<MediaRow> media = rowsOf<MediaRow>(mediaText, mediumFromLine);
    if (media.empty())
        return false;
    {
        OrganRow air;
        air.density = 0.001;
        air.name = "Air";
        organs.push_back(air);
        MediaRow airM;
        airM.composition = { { 7, 0.8 }, { 8, 0.20 } };
        airM.name = "Air";
        media.push_back(airM);
    }
    // The reference edits the voxel array with a sequence of std::replace(old -> new) passes.  The same sequence applied
    // to a 256-entry identity table gives the net map value -> value; `has[v]` follows which values still occur.
    uint8_t map[256];
    bool has[256];
    for (int v = 0; v < 256; ++v) {
        map[v] = static_cast<uint8_t>(v);
        has[v] = present[v] != 0;
    }
    auto replaceAll = [&](uint8_t from, uint8_t to) {
        if (from == to)
            return;
        for (int v = 0; v < 256; ++v)
            if (map[v] == from)
                map[v] = to;
        if (has[from]) {
            has[from] = false;
            has[to] = true;
        }
    };
    if (removeArms) {
        for (const OrganRow& o : organs)
            for (const char* key : { "arm", "hand", "Humeri", "Ulnae" })
                if (o.name.find(key) != std::string::npos)
                    replaceAll(o.id, 0);
    }
    // organs that do not occur are dropped, the others numbered consecutively in id order
    std::stable_sort(organs.begin(), organs.end(), [](const OrganRow& a, const OrganRow& b) { return a.id < b.id; });
    organs.erase(std::remove_if(organs.begin(), organs.end(), [&](const OrganRow& o) { return !has[o.id]; }), organs.end());
    for (size_t i = 0; i < organs.size() && i < 256; ++i)
        if (organs[i].id != i) {
            replaceAll(organs[i].id, static_cast<uint8_t>(i));
            organs[i].id = static_cast<uint8_t>(i);
        }
    // media: keep those an organ refers to, number them consecutively in id order
    std::stable_sort(media.begin(), media.end(), [](const MediaRow& a, const MediaRow& b) { return a.id < b.id; });
    media.erase(std::remove_if(media.begin(), media.end(),
                    [&](const MediaRow& m) {
                        return std::none_of(organs.begin(), organs.end(), [&](const OrganRow& o) { return o.medium == m.id; });
                    }),
        media.end());
    for (size_t i = 0; i < media.size() && i < 256; ++i)
        if (media[i].id != i) {
            for (OrganRow& o : organs)
                if (o.medium == media[i].id)
                    o.medium = static_cast<uint8_t>(i);
            media[i].id = static_cast<uint8_t>(i);
        }
    // voxel look-up tables: organ value -> new organ value, and new organ value -> (medium, density); values that are
    // no organ id map to medium 0 and density 0 (the reference's `contains` tests, :281-296)
    uint8_t mediumOf[256];
    double densityOf[256];
    for (int v = 0; v < 256; ++v) {
        mediumOf[v] = 0;
        densityOf[v] = 0.0;
    }
    for (const OrganRow& o : organs) { // later rows win, like repeated map assignment
        mediumOf[o.id] = o.medium;
        densityOf[o.id] = o.density;
    }
    for (int v = 0; v < 256; ++v) {
        plan.organLut[v] = map[v];
        plan.materialLut[v] = mediumOf[map[v]];
        plan.densityLut[v] = densityOf[map[v]];
    }
    plan.organNames.clear();
    plan.organDensity.clear();
    plan.organMedium.clear();
    for (const OrganRow& o : organs) {
        plan.organNames.push_back(o.name);
        plan.organDensity.push_back(o.density);
        plan.organMedium.push_back(o.medium);
    }
    plan.mediaNames.clear();
    plan.mediaComposition.clear();
    for (const MediaRow& m : media) {
        plan.mediaNames.push_back(m.name);
        plan.mediaComposition.push_back(m.composition);
    }
    return true;
}

} // namespace dxb
