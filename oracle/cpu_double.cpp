// cpu_double.cpp — TEST DOUBLE (oracle/, test infrastructure; never shipped, never loaded by the product).
//
// The nine context-level entry points of include/dxb.h that the C++ shims call from World / AAVoxelGrid / Transport
//   dxb_create  dxb_destroy  dxb_last_error  dxb_set_materials  dxb_set_grid  dxb_set_grid_center  dxb_clear_dose
//   dxb_run  dxb_get_dose
// implemented on the CPU oracle (oracle.cpp).  Linked AHEAD of libdxmc_b200.so into oracle/_ref/opendxmc_ref_cpu, it lets
// OpenDXMC's own SimulationPipeline / worker<CORRECTION>() (compiled unmodified, oracle/Makefile.ref) run end to end on
// a box without a GPU: reference driver -> include/dxmc shims -> these functions -> oracle.  Everything host-side
// (materials, tubes, beams, exposures) still comes from libdxmc_b200.so.  The point is the plumbing and the
// reference's post-processing (R:src/libopendxmc/simulationpipeline.cpp:174-232), not the physics.
#include "oracle.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct dxb_ctx {
    std::vector<dxb_material_tables> tables;
    orc_world* world = nullptr;
    uint64_t n = 0;
    std::vector<double> dose, variance;
    std::vector<uint64_t> events;
    std::string error;
    uint64_t seed = 0x0DDC0FFEEull;
    uint64_t beamCounter = 0; // like the library: beam k runs on key seed + k * DXB_BEAM_KEY_STRIDE (include/dxb.h)
    uint64_t calibrationHistories = 720000; // small: this is a CPU run (env DXB_DOUBLE_CALIB overrides)
};

extern "C" {

int dxb_create(dxb_ctx** out, const int*, int)
{
    if (!out)
        return DXB_EINVAL;
    *out = new dxb_ctx();
    if (const char* e = std::getenv("DXB_DOUBLE_CALIB"))
        (*out)->calibrationHistories = std::strtoull(e, nullptr, 10);
    return DXB_OK;
}

void dxb_destroy(dxb_ctx* c)
{
    if (!c)
        return;
    if (c->world)
        orc_world_destroy(c->world);
    delete c;
}

const char* dxb_last_error(const dxb_ctx* c) { return c ? c->error.c_str() : "null context"; }

int dxb_set_materials(dxb_ctx* c, uint32_t n, const dxb_material* const* materials)
{
    if (!c || !materials || n == 0)
        return DXB_EINVAL;
    c->tables.resize(n);
    for (uint32_t i = 0; i < n; ++i)
        if (dxb_material_tables_get(materials[i], &c->tables[i]) != DXB_OK) {
            c->error = "set_materials: no tables";
            return DXB_EMATERIAL;
        }
    return DXB_OK;
}

int dxb_set_grid(dxb_ctx* c, const uint64_t dim[3], const double spacing_cm[3], const double* density, const uint8_t* material)
{
    if (!c || !dim || !spacing_cm || !density || !material || c->tables.empty())
        return DXB_EINVAL;
    if (c->world)
        orc_world_destroy(c->world);
    c->n = dim[0] * dim[1] * dim[2];
    c->world = orc_world_create(dim, spacing_cm, density, material, static_cast<uint32_t>(c->tables.size()), c->tables.data());
    if (!c->world) {
        c->error = "set_grid: oracle world";
        return DXB_EINVAL;
    }
    dxb_material *air = nullptr, *pmma = nullptr;
    const char* airName = "Air, Dry (near sea level)";
    const char* pmmaName = "Polymethyl Methacralate (Lucite, Perspex)";
    if (dxb_material_by_nist_name(&air, airName) != DXB_OK || dxb_material_by_nist_name(&pmma, pmmaName) != DXB_OK)
        return DXB_EMATERIAL;
    dxb_material_tables ta, tp;
    dxb_material_tables_get(air, &ta);
    dxb_material_tables_get(pmma, &tp);
    orc_world_set_reference_materials(c->world, &ta, &tp, dxb_nist_density(airName), dxb_nist_density(pmmaName)); // deep copies
    dxb_material_destroy(air);
    dxb_material_destroy(pmma);
    return dxb_clear_dose(c);
}

int dxb_set_grid_center(dxb_ctx* c, const double center_cm[3])
{
    if (!c || !center_cm)
        return DXB_EINVAL;
    if (center_cm[0] != 0.0 || center_cm[1] != 0.0 || center_cm[2] != 0.0) {
        c->error = "cpu double: the grid is centred on the origin";
        return DXB_EINVAL;
    }
    return DXB_OK;
}

int dxb_clear_dose(dxb_ctx* c)
{
    if (!c)
        return DXB_EINVAL;
    c->dose.assign(c->n, 0.0);
    c->variance.assign(c->n, 0.0);
    c->events.assign(c->n, 0);
    return DXB_OK;
}

int dxb_run(dxb_ctx* c, const dxb_beam_desc* beam, int physics_mode, int use_beam_calibration, dxb_progress* progress)
{
    if (!c || !beam || !c->world)
        return DXB_ESTATE;
    if (progress && !dxb_progress_continue(progress))
        return DXB_ECANCELLED;
    orc_stats st;
    // accumulates into dose / variance / events like repeated transport() calls on one world
    const uint64_t key = c->seed + c->beamCounter++ * DXB_BEAM_KEY_STRIDE;
    const int rc = orc_transport(c->world, beam, physics_mode, use_beam_calibration, key, c->calibrationHistories, 0, c->dose.data(),
        c->variance.data(), c->events.data(), &st);
    if (rc != 0) {
        c->error = "run: oracle transport failed";
        return DXB_EINVAL;
    }
    return DXB_OK;
}

int dxb_get_dose(dxb_ctx* c, double* dose, double* variance, uint64_t* n_events)
{
    if (!c || !c->world)
        return DXB_ESTATE;
    if (dose)
        std::memcpy(dose, c->dose.data(), c->n * sizeof(double));
    if (variance)
        std::memcpy(variance, c->variance.data(), c->n * sizeof(double));
    if (n_events)
        std::memcpy(n_events, c->events.data(), c->n * sizeof(uint64_t));
    return DXB_OK;
}

} // extern "C"
