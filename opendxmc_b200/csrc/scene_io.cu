// scene_io.cu — OpenDXMC save files <-> a context.  The file layout is the one
// R:src/libopendxmc/hdf5wrapper.cpp:384-459 writes and :1070-1150 reads: root datasets "dimensions" (u64[3]: nx, ny, nz),
// "spacing" (f64[3], cm), "densityarray" (f64) / "materialarray" (u8) / "organarray" (u8) with HDF5 dims (nz, ny, nx)
// (:121-124: the x index varies fastest, like the arrays handed to AAVoxelGrid::setData), "materialnames" and
// "materialcomposition" (variable-length strings, e.g. "H0.111894O0.888106": element symbol + weight, :425-434), and after a
// simulation "dosearray", "dosevariancearray", "doseeventcountarray" (all f64, DataContainer's arrays).
#include "context_types.hpp"
#include "h5mini.hpp"

#include <cctype>
#include <map>

using namespace dxb;

namespace {

// "H0.111894O0.888106" -> {1: 0.111894, 8: 0.888106}: what dxmc::Material<5>::parseCompoundStr returns and the worker
// hands to Material::byWeight (R:src/libopendxmc/hdf5wrapper.cpp:1100, simulationpipeline.cpp:136)
std::map<uint32_t, double> parseComposition(const std::string& str)
{
    std::map<uint32_t, double> res;
    size_t i = 0;
    while (i < str.size()) {
        if (!std::isupper(static_cast<unsigned char>(str[i]))) {
            ++i;
            continue;
        }
        std::string sym(1, str[i++]);
        while (i < str.size() && std::islower(static_cast<unsigned char>(str[i])))
            sym.push_back(str[i++]);
        std::string num;
        while (i < str.size() && (std::isdigit(static_cast<unsigned char>(str[i])) || str[i] == '.'))
            num.push_back(str[i++]);
        for (uint32_t z = 1; z <= 92; ++z) {
            const Element* el = getElement(z);
            if (el && sym == el->symbol) {
                res[z] += num.empty() ? 1.0 : std::stod(num);
                break;
            }
        }
    }
    return res;
}

template <typename T>
bool numeric(const h5mini::Dataset* d, h5mini::Type t, size_t count, std::vector<T>& out)
{
    if (!d || d->type != t || d->count() != count || d->data.size() != count * sizeof(T))
        return false;
    out.resize(count);
    std::memcpy(out.data(), d->data.data(), d->data.size());
    return true;
}

void putArray(h5mini::File& f, const char* name, h5mini::Type t, const uint64_t dim[3], const void* data, size_t bytes)
{
    h5mini::Dataset& d = f.createDataset(name);
    d = h5mini::Dataset();
    d.type = t;
    d.dims = { dim[2], dim[1], dim[0] }; // z, y, x
    d.deflate = true;
    d.data.assign(static_cast<const uint8_t*>(data), static_cast<const uint8_t*>(data) + bytes);
}

} // namespace

extern "C" {

int dxb_load_scene(dxb_ctx* c, const char* path, uint64_t dim_out[3], double spacing_out[3], uint32_t* n_materials_out)
{
    if (!c || !path)
        return DXB_EINVAL;
    std::string err;
    auto f = h5mini::File::load(path, &err);
    if (!f)
        return fail(c, DXB_EINVAL, "load_scene: " + err);
    std::vector<uint64_t> dim;
    std::vector<double> spacing, density;
    std::vector<uint8_t> material;
    if (!numeric(f->dataset("dimensions"), h5mini::Type::U64, 3, dim) || !numeric(f->dataset("spacing"), h5mini::Type::F64, 3, spacing))
        return fail(c, DXB_EINVAL, "load_scene: no dimensions / spacing");
    const size_t n = static_cast<size_t>(dim[0]) * dim[1] * dim[2];
    if (!numeric(f->dataset("materialarray"), h5mini::Type::U8, n, material) || !numeric(f->dataset("densityarray"), h5mini::Type::F64, n, density))
        return fail(c, DXB_EINVAL, "load_scene: material / density array missing or of the wrong size");
    const h5mini::Dataset* names = f->dataset("materialnames");
    const h5mini::Dataset* comps = f->dataset("materialcomposition");
    if (!names || !comps || names->type != h5mini::Type::String || comps->type != h5mini::Type::String
        || names->strings.size() != comps->strings.size() || comps->strings.empty())
        return fail(c, DXB_EINVAL, "load_scene: material names / compositions missing");
    std::vector<dxb_material> owners;
    std::vector<const dxb_material*> handles;
    for (const std::string& s : comps->strings) {
        auto m = Material::byWeight(parseComposition(s));
        if (!m)
            return fail(c, DXB_EMATERIAL, "load_scene: cannot build material '" + s + "'"); // Material::byWeight -> nullopt
        owners.push_back(dxb_material { m });
    }
    for (const dxb_material& m : owners)
        handles.push_back(&m);
    int rc = dxb_set_materials(c, static_cast<uint32_t>(handles.size()), handles.data());
    if (rc != DXB_OK)
        return rc;
    rc = dxb_set_grid(c, dim.data(), spacing.data(), density.data(), material.data());
    if (rc != DXB_OK)
        return rc;
    // the scene stays with the context: dxb_save_dose writes it back next to the result
    c->scene = std::shared_ptr<void>(f.release(), [](void* p) { delete static_cast<h5mini::File*>(p); });
    for (int i = 0; i < 3; ++i) {
        if (dim_out)
            dim_out[i] = dim[i];
        if (spacing_out)
            spacing_out[i] = spacing[i];
    }
    if (n_materials_out)
        *n_materials_out = static_cast<uint32_t>(handles.size());
    return DXB_OK;
}

int dxb_save_dose(dxb_ctx* c, const char* path, int delete_air_dose, char units_out[4])
{
    if (!c || !path)
        return DXB_EINVAL;
    if (c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "save_dose: no grid");
    const World& w = c->devs[0]->world;
    const size_t n = w.nvox;
    std::vector<double> dose(n), variance(n), events(n);
    char units[4] = { 0, 0, 0, 0 };
    // DataContainer holds the arrays AFTER the driver's post-processing: air mask, uGy rule
    // (R:src/libopendxmc/simulationpipeline.cpp:174-232)
    const int rc = dxb_get_dose_postprocessed(c, delete_air_dose, dose.data(), variance.data(), events.data(), units);
    if (rc != DXB_OK)
        return rc;
    h5mini::File out;
    h5mini::File* f = c->scene ? static_cast<h5mini::File*>(c->scene.get()) : &out;
    const uint64_t dim[3] = { w.dim[0], w.dim[1], w.dim[2] };
    if (!c->scene) {
        // no scene file behind this context: dimensions, spacing and the result alone (the reference's load() also
        // wants material and density arrays; dxb_load_scene keeps them)
        h5mini::Dataset& d = f->createDataset("dimensions");
        d.type = h5mini::Type::U64;
        d.dims = { 3 };
        d.data.assign(reinterpret_cast<const uint8_t*>(dim), reinterpret_cast<const uint8_t*>(dim) + 24);
        h5mini::Dataset& s = f->createDataset("spacing");
        s.type = h5mini::Type::F64;
        s.dims = { 3 };
        s.data.assign(reinterpret_cast<const uint8_t*>(w.spacing), reinterpret_cast<const uint8_t*>(w.spacing) + 24);
    }
    putArray(*f, "dosearray", h5mini::Type::F64, dim, dose.data(), n * 8);
    putArray(*f, "dosevariancearray", h5mini::Type::F64, dim, variance.data(), n * 8);
    putArray(*f, "doseeventcountarray", h5mini::Type::F64, dim, events.data(), n * 8);
    std::string err;
    if (!f->save(path, &err))
        return fail(c, DXB_EINVAL, "save_dose: " + err);
    if (units_out)
        std::memcpy(units_out, units, 4);
    return DXB_OK;
}

} // extern "C"
