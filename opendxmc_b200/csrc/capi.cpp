// capi.cpp — host-only part of the C ABI (include/dxb.h): materials, NIST table, tube, filters, beams.
#include "internal.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>

using namespace dxb;

extern "C" {

int dxb_abi_version(void) { return DXB_ABI_VERSION; }

// ----------------------------------------------------------------------- materials
int dxb_material_by_weight(dxb_material** out, uint32_t n, const uint32_t* Z, const double* weight)
{
    if (!out)
        return DXB_EINVAL;
    *out = nullptr;
    if (n == 0 || !Z || !weight)
        return DXB_EMATERIAL;
    std::map<uint32_t, double> w;
    for (uint32_t i = 0; i < n; ++i) {
        if (Z[i] < 1 || Z[i] > 92 || !(weight[i] >= 0) || !std::isfinite(weight[i]))
            return DXB_EMATERIAL;
        w[Z[i]] += weight[i];
    }
    auto m = Material::byWeight(w);
    if (!m)
        return DXB_EMATERIAL;
    *out = new dxb_material { m };
    return DXB_OK;
}

int dxb_material_by_nist_name(dxb_material** out, const char* name)
{
    if (!out)
        return DXB_EINVAL;
    *out = nullptr;
    if (!name)
        return DXB_EMATERIAL;
    auto m = Material::byNistName(name);
    if (!m)
        return DXB_EMATERIAL;
    *out = new dxb_material { m };
    return DXB_OK;
}

int dxb_material_by_chemical_formula(dxb_material** out, const char* formula)
{
    if (!out)
        return DXB_EINVAL;
    *out = nullptr;
    if (!formula)
        return DXB_EMATERIAL;
    auto m = Material::byChemicalFormula(formula);
    if (!m)
        return DXB_EMATERIAL;
    *out = new dxb_material { m };
    return DXB_OK;
}

void dxb_material_destroy(dxb_material* m) { delete m; }

int dxb_material_attenuation(const dxb_material* m, double energy_kev, double out_pic[3])
{
    if (!m || !out_pic)
        return DXB_EINVAL;
    m->m->attenuation(energy_kev, out_pic);
    return DXB_OK;
}
double dxb_material_mass_energy_transfer(const dxb_material* m, double e) { return m ? m->m->massEnergyTransfer(e) : 0.0; }
double dxb_material_effective_z(const dxb_material* m) { return m ? m->m->effectiveZ : 0.0; }
double dxb_material_form_factor(const dxb_material* m, double x) { return m ? m->m->formFactor(x) : 0.0; }
double dxb_material_scatter_factor(const dxb_material* m, double x) { return m ? m->m->scatterFactor(x) : 0.0; }

int dxb_material_tables_get(const dxb_material* m, dxb_material_tables* t)
{
    if (!m || !t)
        return DXB_EINVAL;
    const Material& M = *m->m;
    std::memset(t, 0, sizeof(*t));
    t->n_energy = kNEnergy;
    t->e_min_kev = kEMin;
    t->e_max_kev = energyNode(kNEnergy - 1);
    t->photo = M.photo.data();
    t->incoh = M.incoh.data();
    t->coh = M.coh.data();
    t->incoh_kn = M.incoh_kn.data();
    t->coh_thomson = M.coh.data();
    t->etr = M.etr.data();
    t->n_x = kNX;
    t->x_min = kXMin;
    t->x_max = xNode(kNX - 1);
    t->ff_cdf = M.ffCdf.data();
    t->sf = M.sf.data();
    t->n_shells = M.nShells;
    for (uint32_t i = 0; i < M.nShells; ++i)
        t->shells[i] = M.shells[i];
    t->rest_electrons_fraction = M.restElectronsFraction;
    t->rest_compton_j0 = M.restComptonJ0;
    t->electrons_per_gram = M.electronsPerGram;
    t->effective_z = M.effectiveZ;
    t->nodes_per_octave_e = kENodesPerOctave;
    t->nodes_per_octave_x = kXNodesPerOctave;
    return DXB_OK;
}
// The drop-in route for externally built physics data (e.g. EPICS2014-derived arrays resampled on the library's grids):
// the tables are taken as they are, nothing is recomputed from the analytic atom model.
int dxb_material_from_tables(dxb_material** out, const dxb_material_tables* t)
{
    if (!out || !t)
        return DXB_EINVAL;
    *out = nullptr;
    if (t->n_energy != kNEnergy || t->n_x != kNX || t->nodes_per_octave_e != kENodesPerOctave || t->nodes_per_octave_x != kXNodesPerOctave
        || std::fabs(t->e_min_kev - kEMin) > 1e-12 * kEMin || std::fabs(t->x_min - kXMin) > 1e-12 * kXMin)
        return DXB_EINVAL; // wrong grid geometry: resample on node(i) = min * 2^(i/P) * (1 + (i%P)/P), see dxb_material_tables
    if (!t->photo || !t->incoh || !t->coh || !t->etr || !t->ff_cdf || !t->sf || t->n_shells > DXB_MAX_SHELLS)
        return DXB_EINVAL;
    auto ok = [](const double* a, uint32_t n) {
        for (uint32_t i = 0; i < n; ++i)
            if (!(a[i] >= 0.0) || !std::isfinite(a[i]))
                return false;
        return true;
    };
    if (!ok(t->photo, kNEnergy) || !ok(t->incoh, kNEnergy) || !ok(t->coh, kNEnergy) || !ok(t->etr, kNEnergy) || !ok(t->ff_cdf, kNX) || !ok(t->sf, kNX))
        return DXB_EMATERIAL;
    for (uint32_t i = 1; i < kNX; ++i)
        if (t->ff_cdf[i] < t->ff_cdf[i - 1])
            return DXB_EMATERIAL; // a cumulative distribution
    auto m = std::make_shared<Material>();
    m->photo.assign(t->photo, t->photo + kNEnergy);
    m->incoh.assign(t->incoh, t->incoh + kNEnergy);
    m->coh.assign(t->coh, t->coh + kNEnergy);
    m->incoh_kn.assign(t->incoh_kn ? t->incoh_kn : t->incoh, (t->incoh_kn ? t->incoh_kn : t->incoh) + kNEnergy);
    m->etr.assign(t->etr, t->etr + kNEnergy);
    m->ffCdf.assign(t->ff_cdf, t->ff_cdf + kNX);
    m->sf.assign(t->sf, t->sf + kNX);
    // F^2 / Z^2 on the x grid from the slope of the cumulative A(x^2) (only the F(x) getter uses it)
    m->ff2.assign(kNX, 0.0);
    for (uint32_t i = 0; i < kNX; ++i) {
        const uint32_t a = i == 0 ? 0 : i - 1, b = i + 1 < kNX ? i + 1 : i;
        const double ta = xNode(a) * xNode(a), tb = xNode(b) * xNode(b);
        m->ff2[i] = tb > ta ? std::max(0.0, (t->ff_cdf[b] - t->ff_cdf[a]) / (tb - ta)) : 0.0;
    }
    m->nShells = t->n_shells;
    for (uint32_t i = 0; i < t->n_shells; ++i)
        m->shells[i] = t->shells[i];
    m->restElectronsFraction = t->rest_electrons_fraction;
    m->restComptonJ0 = t->rest_compton_j0;
    m->electronsPerGram = t->electrons_per_gram;
    m->effectiveZ = t->effective_z;
    m->sumAZ = m->sumAZ2 = 1.0;
    *out = new dxb_material { m };
    return DXB_OK;
}
uint32_t dxb_table_n_energy(void) { return kNEnergy; }
double dxb_table_e_min(void) { return kEMin; }
double dxb_table_e_max(void) { return kEMax; }

// ----------------------------------------------------------------------- NIST / atoms
int dxb_nist_count(void) { return static_cast<int>(nistTable().size()); }
const char* dxb_nist_name(int i)
{
    const auto& t = nistTable();
    return (i >= 0 && i < static_cast<int>(t.size())) ? t[i].name : nullptr;
}
double dxb_nist_density(const char* name)
{
    const NistEntry* e = name ? nistFind(name) : nullptr;
    return e ? e->density : -1.0;
}
int dxb_nist_composition(const char* name, uint32_t* Z, double* weight, int cap)
{
    const NistEntry* e = name ? nistFind(name) : nullptr;
    if (!e)
        return 0;
    const int n = static_cast<int>(e->w.size());
    for (int i = 0; i < std::min(n, cap); ++i) {
        if (Z)
            Z[i] = e->w[i].first;
        if (weight)
            weight[i] = e->w[i].second;
    }
    return n;
}
const char* dxb_atom_symbol(uint32_t Z)
{
    const Element* e = getElement(Z);
    return e ? e->symbol : "";
}
double dxb_atom_weight(uint32_t Z)
{
    const Element* e = getElement(Z);
    return e ? e->A : 0.0;
}
double dxb_atom_standard_density(uint32_t Z)
{
    const Element* e = getElement(Z);
    return e ? e->density : 0.0;
}

// ----------------------------------------------------------------------- tube
int dxb_tube_energies(const dxb_tube_desc* t, double* energy, int cap)
{
    if (!t)
        return 0;
    const auto e = tubeEnergies(*t);
    if (energy)
        for (int i = 0; i < std::min<int>(cap, static_cast<int>(e.size())); ++i)
            energy[i] = e[i];
    return static_cast<int>(e.size());
}
int dxb_tube_spectrum(const dxb_tube_desc* t, const double* energy, int n, int normalize, double* weight)
{
    if (!t || !energy || !weight || n <= 0)
        return DXB_EINVAL;
    const std::vector<double> e(energy, energy + n);
    const auto w = tubeSpectrum(*t, e, normalize != 0);
    std::copy(w.begin(), w.end(), weight);
    return DXB_OK;
}
double dxb_tube_mean_energy(const dxb_tube_desc* t) { return t ? tubeMeanEnergy(*t) : 0.0; }
double dxb_tube_al_half_value_layer_mm(const dxb_tube_desc* t) { return t ? tubeAlHVLmm(*t) : 0.0; }

// ----------------------------------------------------------------------- beams
void dxb_beam_desc_init(dxb_beam_desc* b, int type)
{
    if (!b)
        return;
    std::memset(b, 0, sizeof(*b));
    b->type = type;
    b->n_exposures = 1;
    b->particles_per_exposure = 1000000; // R:src/libopendxmc/beamsettingsmodel.cpp:470-474,1171
    b->direction[2] = 1.0;
    b->stop[2] = 1.0;
    b->sdd = 119.0;                      // R:...beamsettingsmodel.cpp:1168
    b->fov = 50.0;
    b->fov_b = 33.0;
    b->collimation = 3.84;               // R:...beamsettingsmodel.cpp:1169
    b->pitch = 1.0;
    b->step_angle = 5.0 * kPi / 180.0;   // R:...beamsettingsmodel.cpp:1170
    b->stop_angle = 2.0 * kPi;
    b->n_slices = 1;
    b->slice_spacing = 3.84;
    b->relative_mas_a = b->relative_mas_b = 1.0;
    b->ctdi = 1.0;
    b->ctdi_diameter = 32.0;
    b->dap = 1.0;
    b->air_kerma = 1.0;
    b->energy = 60.0;
    b->organ_aec.low_weight = 0.6;
    b->organ_aec.ramp_angle = 20.0 * kPi / 180.0;
    b->organ_aec.start_angle = 0.0;
    b->organ_aec.stop_angle = kPi;
    if (type == DXB_BEAM_DX) {
        // OpenDXMC's DXBeam: cosines {1,0,0},{0,-1,0}; 20 x 20 cm at SDD 100 stored as tan(half size / SDD)
        // (R:src/libopendxmc/dxmc_specialization.cpp:22-26,46-52)
        b->cosines[0][0] = 1.0;
        b->cosines[1][1] = -1.0;
        b->sdd = 100.0;
        b->half_angles[0] = b->half_angles[1] = std::tan(0.5 * 20.0 / 100.0);
    }
    if (type == DXB_BEAM_CBCT) {
        b->sdd = 100.0;
        b->step_angle = kPi / 180.0;
        b->half_angles[0] = b->half_angles[1] = 5.0 * kPi / 180.0;
    }
}

uint64_t dxb_beam_number_of_exposures(const dxb_beam_desc* b) { return b ? beamNumberOfExposures(*b) : 0; }
uint64_t dxb_beam_number_of_particles(const dxb_beam_desc* b)
{
    return b ? beamNumberOfExposures(*b) * b->particles_per_exposure : 0;
}
int dxb_beam_exposure(const dxb_beam_desc* b, uint64_t index, dxb_exposure* out)
{
    if (!b || !out)
        return DXB_EINVAL;
    const AecTable aec = makeAec(b->aec);
    return beamExposure(*b, index, aec, *out);
}
double dxb_bowtie_weight(const dxb_bowtie* b, double angle) { return b ? makeBowtie(*b)(angle) : 1.0; }
double dxb_aec_weight(const dxb_aec* a, const double position[3]) { return (a && position) ? makeAec(*a)(position) : 1.0; }
double dxb_organ_aec_weight(const dxb_organ_aec* o, double angle) { return o ? organAecWeight(*o, angle) : 1.0; }
double dxb_organ_aec_max_weight(const dxb_organ_aec* o) { return o ? organAecMaxWeight(*o) : 1.0; }
double dxb_beam_analytic_calibration(const dxb_beam_desc* b) { return b ? beamAnalyticCalibration(*b) : 0.0; }

} // extern "C"
