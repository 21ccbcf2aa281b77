// physics.hpp — host-side physics data model of libdxmc_b200.
//
// Replaces DXMClib's AtomHandler / Material<5> / NISTMaterials (absent from
// /root/reference).  DXMClib reads EPICS2014 cross sections
// from `physicslists.bin`; no cross-section data exists in this build
// environment, so every table here is GENERATED from an analytic atomic model
// (documented in DESIGN.md §Physics data).  The table FORMAT (dxb_material_tables
// in include/dxb.h) is source-agnostic: EPICS-derived arrays can be dropped in
// without touching the transport kernels.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/dxb.h"

namespace dxb {

// ---- constants (CODATA) ----------------------------------------------------
constexpr double kElectronRestMassKeV = 510.99895;
constexpr double kHcKeVAngstrom = 12.398419843;       // h*c
constexpr double kClassicalElectronRadiusSq = 7.9407877e-26; // cm^2
constexpr double kAvogadro = 6.02214076e23;
constexpr double kBohrRadiusAngstrom = 0.529177210903;
constexpr double kRydbergKeV = 0.013605693123;
constexpr double kFineStructure = 1.0 / 137.035999084;
constexpr double kKeVperGramToMilliGray = 1.602176634e-10; // 1 keV/g = 1.602e-16 J / 1e-3 kg = 1.602e-13 Gy
constexpr double kPi = 3.14159265358979323846;

// ---- table geometry (shared by every material, the device and the oracle) ---
// Both grids are semi-log: a power-of-two number P of nodes per octave, uniformly spaced inside the octave,
// node(i) = vmin * 2^(i/P) * (1 + (i%P)/P).  The device reads the grid coordinate straight from the float bit
// pattern (exponent + top mantissa bits = index, remaining mantissa bits = fraction), which is exact, so f32
// lookups stay within 1e-6 of the f64 ones and the hot loop needs no logarithm.
constexpr uint32_t kENodesPerOctave = 64;
constexpr double kEMin = 1.0;        // keV  (DXMClib MIN_ENERGY, SURVEY.md §8c "1 keV cutoff")
constexpr double kEMax = 150.0;      // keV  (DXMClib MAX_ENERGY); grid reaches 128 * (1 + 16/64) = 160
constexpr uint32_t kNEnergy = 7 * 64 + 16 + 1; // 465 nodes: 1 .. 160 keV
constexpr uint32_t kXNodesPerOctave = 32;
constexpr double kXMin = 1.0 / 128.0; // 1/Angstrom
constexpr uint32_t kNX = 11 * 32 + 1; // 353 nodes: 2^-7 .. 2^4 = 16 > kEMax/hc = 12.1
constexpr double kXMax = 16.0;

double energyNode(uint32_t i);
double xNode(uint32_t i);

// ---- elements ----------------------------------------------------------------
struct SlaterGroup {
    int n;            // principal quantum number
    int kind;         // 0 = s/p group, 1 = d, 2 = f
    double electrons;
    double nstar;     // effective principal quantum number
    double zeta;      // orbital exponent (Z - s)/n*   [1/a0]
    double binding_kev;
};

struct Element {
    uint32_t Z = 0;
    double A = 0;                 // g/mol
    double density = 0;           // g/cm3 (standard state), for filters
    const char* symbol = "";
    std::vector<SlaterGroup> groups;
    double edgeK = 0, edgeL1 = 0, edgeL2 = 0, edgeL3 = 0; // keV (0: below table range / not modelled)
    // cross sections per atom [cm^2] on the common energy grid
    std::vector<double> photo, incoh, coh, incoh_kn, etr_incoh; // etr_incoh: incoh * mean fraction of E transferred
    // form factor F(x) and incoherent scattering function S(x) on a fine x grid
    std::vector<double> ffx, sfx;   // on fineX()
    double formFactor(double x) const;   // F(x)
    double scatterFunction(double x) const; // S(x) in [0, Z]
    double photoelectric(double energy_kev) const; // cm^2 / atom, analytic (edges included)
    double jumpK() const;  // K-edge jump ratio
    double fluorYieldK() const;
    double kAlphaEnergy() const;
};

const Element* getElement(uint32_t Z);   // nullptr if Z not in 1..92; cached, thread safe
double kleinNishinaTotal(double energy_kev); // cm^2 / electron

// ---- materials ---------------------------------------------------------------
struct Material {
    std::map<uint32_t, double> massFraction;   // normalised
    std::vector<double> photo, incoh, coh, incoh_kn, etr; // [kNEnergy], cm^2/g
    std::vector<double> ffCdf, sf;             // [kNX]
    std::vector<double> ff2;                   // [kNX]  F^2 per average atom / sum(a Z^2)
    uint32_t nShells = 0;
    dxb_shell shells[DXB_MAX_SHELLS] = {};
    double restElectronsFraction = 1.0;
    double restComptonJ0 = 0.0; // J(0) of the electrons not covered by `shells`
    double electronsPerGram = 0;
    double effectiveZ = 0;
    double sumAZ = 0, sumAZ2 = 0;              // per average atom
    double meanAtomicWeight = 0;

    static std::shared_ptr<Material> byWeight(const std::map<uint32_t, double>& w);
    static std::shared_ptr<Material> byNistName(const std::string& name);
    static std::shared_ptr<Material> byChemicalFormula(const std::string& formula);

    // double-precision evaluation with the SAME interpolation rule the device uses
    void attenuation(double e, double out[3]) const;
    double total(double e) const;
    double massEnergyTransfer(double e) const;
    double formFactor(double x) const;     // sqrt(sum a_i F_i^2)
    double scatterFactor(double x) const;  // S/Z in [0,1]
};

// interpolation helpers on the common grids (value tables, lin-interp in log coordinate)
struct GridPos { uint32_t i; double f; };
GridPos energyPos(double e);
GridPos xPos(double x);
double lerpTable(const std::vector<double>& t, GridPos p);

// ---- NIST compounds -----------------------------------------------------------
struct NistEntry { const char* name; double density; std::vector<std::pair<uint32_t, double>> w; };
const std::vector<NistEntry>& nistTable();
const NistEntry* nistFind(const std::string& name);

// ---- tube ---------------------------------------------------------------------
std::vector<double> tubeEnergies(const dxb_tube_desc& t);
std::vector<double> tubeSpectrum(const dxb_tube_desc& t, const std::vector<double>& energies, bool normalize);
double tubeMeanEnergy(const dxb_tube_desc& t);
double tubeAlHVLmm(const dxb_tube_desc& t);

// ---- filters ------------------------------------------------------------------
struct BowtieTable {           // sorted, normalised, piecewise-linear in |angle|
    std::vector<double> angle, weight;
    bool empty() const { return angle.size() < 2; }
    double operator()(double a) const;
};
BowtieTable makeBowtie(const dxb_bowtie& b);

struct AecTable {
    std::vector<double> weights;     // normalised to mean 1
    double start[3] = {0, 0, 0}, dir[3] = {0, 0, 1}, length = 0;
    bool empty() const { return weights.size() < 2 || length <= 0; }
    double operator()(const double pos[3]) const;
};
AecTable makeAec(const dxb_aec& a);
double organAecWeight(const dxb_organ_aec& o, double angle);
double organAecMaxWeight(const dxb_organ_aec& o);

// ---- beams -> exposures ---------------------------------------------------------
uint64_t beamNumberOfExposures(const dxb_beam_desc& b);
int beamExposure(const dxb_beam_desc& b, uint64_t i, const AecTable& aec, dxb_exposure& out);
double beamAnalyticCalibration(const dxb_beam_desc& b);

// ---- small vector helpers (dxmc::vectormath, R:src/libopendxmc/dxmc_specialization.cpp:64-88) ----
namespace vm {
inline void cross(const double a[3], const double b[3], double r[3])
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
double norm(const double a[3]);
void normalize(double a[3]);
void rotate(const double v[3], const double axis[3], double angle, double r[3]); // Rodrigues
int argminAbs(const double a[3]);
}

} // namespace dxb
