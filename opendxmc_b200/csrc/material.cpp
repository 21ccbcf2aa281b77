// material.cpp — dxmc::Material<5> / dxmc::NISTMaterials replacement (host side).
// Consumed at R:src/libopendxmc/simulationpipeline.cpp:134-144 (byWeight),
// R:src/libopendxmc/ctsegmentationpipeline.cpp:73-85 (byNistName, attenuationValues),
// R:src/libopendxmc/otherphantomimportpipeline.cpp:87-93 (NISTMaterials::density).
#include "physics.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>

namespace dxb {

// ---------------------------------------------------------------------------
// NIST compounds (mass fractions).  Only the names OpenDXMC uses are required
// (SURVEY.md §8a row a16); a few more are included for phantoms and tests.
// ---------------------------------------------------------------------------
const std::vector<NistEntry>& nistTable()
{
    static const std::vector<NistEntry> t = {
        { "Air, Dry (near sea level)", 1.20479e-3, { { 6, 0.000124 }, { 7, 0.755268 }, { 8, 0.231781 }, { 18, 0.012827 } } },
        { "Water, Liquid", 1.0, { { 1, 0.111894 }, { 8, 0.888106 } } },
        { "Adipose Tissue (ICRP)", 0.92,
            { { 1, 0.119477 }, { 6, 0.637240 }, { 7, 0.007970 }, { 8, 0.232333 }, { 11, 0.000500 }, { 12, 0.000020 },
                { 15, 0.000160 }, { 16, 0.000730 }, { 17, 0.001190 }, { 19, 0.000320 }, { 20, 0.000020 }, { 26, 0.000020 },
                { 30, 0.000020 } } },
        { "Tissue, Soft (ICRP)", 1.0,
            { { 1, 0.104472 }, { 6, 0.232190 }, { 7, 0.024880 }, { 8, 0.630238 }, { 11, 0.001130 }, { 12, 0.000130 },
                { 15, 0.001330 }, { 16, 0.001990 }, { 17, 0.001340 }, { 19, 0.001990 }, { 20, 0.000230 }, { 26, 0.000050 },
                { 30, 0.000030 } } },
        { "Muscle, Skeletal", 1.04,
            { { 1, 0.100637 }, { 6, 0.107830 }, { 7, 0.027680 }, { 8, 0.754773 }, { 11, 0.000750 }, { 12, 0.000190 },
                { 15, 0.001800 }, { 16, 0.002410 }, { 17, 0.000790 }, { 19, 0.003020 }, { 20, 0.000030 }, { 26, 0.000040 },
                { 30, 0.000050 } } },
        { "Bone, Cortical (ICRP)", 1.85,
            { { 1, 0.047234 }, { 6, 0.144330 }, { 7, 0.041990 }, { 8, 0.446096 }, { 12, 0.002200 }, { 15, 0.104970 },
                { 16, 0.003150 }, { 20, 0.209930 }, { 30, 0.000100 } } },
        { "Polymethyl Methacralate (Lucite, Perspex)", 1.19, { { 1, 0.080538 }, { 6, 0.599848 }, { 8, 0.319614 } } },
        { "Lung (ICRP)", 1.05,
            { { 1, 0.101278 }, { 6, 0.102310 }, { 7, 0.028650 }, { 8, 0.757072 }, { 11, 0.001840 }, { 12, 0.000730 },
                { 15, 0.000800 }, { 16, 0.002250 }, { 17, 0.002660 }, { 19, 0.001940 }, { 20, 0.000090 }, { 26, 0.000370 },
                { 30, 0.000010 } } },
        { "Bone, Compact (ICRU)", 1.85,
            { { 1, 0.063984 }, { 6, 0.278000 }, { 7, 0.027000 }, { 8, 0.410016 }, { 12, 0.002000 }, { 15, 0.070000 },
                { 16, 0.002000 }, { 20, 0.147000 } } },
        { "Polyethylene", 0.94, { { 1, 0.143711 }, { 6, 0.856289 } } },
        { "Aluminum", 2.699, { { 13, 1.0 } } },
    };
    return t;
}

const NistEntry* nistFind(const std::string& name)
{
    for (const auto& e : nistTable())
        if (name == e.name)
            return &e;
    return nullptr;
}

// ---------------------------------------------------------------------------
namespace {

void buildShells(Material& m, const std::vector<std::pair<const Element*, double>>& atoms /* (element, atom fraction) */)
{
    // Candidate shells: every Slater group of every element, weighted by atom fraction.
    struct Cand {
        double binding, electrons, j0, fluorYield, fluorE, photoShare;
        bool isK;
    };
    std::vector<Cand> cands;
    double totalElectrons = 0;
    for (const auto& [el, a] : atoms) {
        for (const auto& g : el->groups) {
            Cand c;
            c.binding = g.binding_kev;
            c.electrons = g.electrons * a;
            // Compton profile at pz = 0 of a Slater-type orbital, J(0) ~ c_n / zeta (1s exact: 8/(3 pi zeta));
            // atomic units of momentum.
            c.j0 = 8.0 / (3.0 * kPi * g.zeta) * (g.n == 1 ? 1.0 : 1.0 + 0.25 * (g.n - 1));
            c.isK = (g.n == 1);
            c.fluorYield = c.isK ? el->fluorYieldK() : 0.0;
            c.fluorE = c.isK ? el->kAlphaEnergy() : 0.0;
            c.photoShare = 0;
            cands.push_back(c);
            totalElectrons += c.electrons;
        }
    }
    // keep the most tightly bound shells that still hold a noticeable share of electrons
    std::sort(cands.begin(), cands.end(), [](const Cand& x, const Cand& y) { return x.binding > y.binding; });
    m.nShells = 0;
    double covered = 0, restElectrons = 0, restJ = 0;
    bool tableClosed = false;
    for (const auto& c : cands) {
        // the table keeps the most tightly bound groups above the transport cut-off; everything else forms one
        // unbound group whose profile height is the electron-weighted mean (per-electron profiles add linearly)
        const bool keep = !tableClosed && m.nShells < DXB_MAX_SHELLS && c.binding >= kEMin && c.electrons / totalElectrons >= 1e-5;
        if (!keep) {
            if (c.binding < kEMin || m.nShells >= DXB_MAX_SHELLS)
                tableClosed = true;
            restElectrons += c.electrons;
            restJ += c.electrons * c.j0;
            continue;
        }
        dxb_shell& s = m.shells[m.nShells++];
        s.binding_energy_kev = c.binding;
        s.n_electrons = c.electrons;
        s.n_electrons_fraction = c.electrons / totalElectrons;
        s.fluor_yield = c.fluorYield;
        s.fluor_energy_kev = c.fluorE;
        s.compton_j0 = c.j0;
        s.photo_fraction_above = 0; // filled by the caller (needs cross sections)
        covered += s.n_electrons_fraction;
    }
    m.restComptonJ0 = restElectrons > 0 ? restJ / restElectrons : 0.0;
    m.restElectronsFraction = std::max(0.0, 1.0 - covered);
}

} // namespace

std::shared_ptr<Material> Material::byWeight(const std::map<uint32_t, double>& w)
{
    if (w.empty())
        return nullptr;
    double total = 0;
    for (const auto& [Z, x] : w) {
        if (Z < 1 || Z > 92 || !(x >= 0) || !std::isfinite(x))
            return nullptr;
        total += x;
    }
    if (!(total > 0))
        return nullptr;

    auto m = std::make_shared<Material>();
    std::vector<std::pair<const Element*, double>> atoms; // atom-number fractions
    double molSum = 0;
    for (const auto& [Z, x] : w) {
        if (x <= 0)
            continue;
        const Element* el = getElement(Z);
        const double wf = x / total;
        m->massFraction[Z] = wf;
        atoms.emplace_back(el, wf / el->A);
        molSum += wf / el->A;
    }
    for (auto& a : atoms)
        a.second /= molSum;
    m->meanAtomicWeight = 1.0 / molSum;

    m->photo.assign(kNEnergy, 0.0);
    m->incoh.assign(kNEnergy, 0.0);
    m->coh.assign(kNEnergy, 0.0);
    m->incoh_kn.assign(kNEnergy, 0.0);
    m->etr.assign(kNEnergy, 0.0);
    m->electronsPerGram = 0;
    for (const auto& [Z, wf] : m->massFraction) {
        const Element* el = getElement(Z);
        const double nPerGram = kAvogadro * wf / el->A; // atoms of Z per gram
        m->electronsPerGram += nPerGram * Z;
        for (uint32_t i = 0; i < kNEnergy; ++i) {
            m->photo[i] += nPerGram * el->photo[i];
            m->incoh[i] += nPerGram * el->incoh[i];
            m->coh[i] += nPerGram * el->coh[i];
            m->incoh_kn[i] += nPerGram * el->incoh_kn[i];
            m->etr[i] += nPerGram * (el->photo[i] + el->etr_incoh[i]);
        }
    }
    // per-average-atom sums
    m->sumAZ = 0;
    m->sumAZ2 = 0;
    double zeffNum = 0, zeffDen = 0;
    for (const auto& [el, a] : atoms) {
        m->sumAZ += a * el->Z;
        m->sumAZ2 += a * el->Z * el->Z;
        zeffNum += a * el->Z * std::pow(static_cast<double>(el->Z), 2.94);
        zeffDen += a * el->Z;
    }
    m->effectiveZ = std::pow(zeffNum / zeffDen, 1.0 / 2.94);

    // form factor / scatter function of the mixture (independent-atom approximation)
    m->ff2.resize(kNX);
    m->sf.resize(kNX);
    for (uint32_t i = 0; i < kNX; ++i) {
        const double x = xNode(i);
        double f2 = 0, s = 0;
        for (const auto& [el, a] : atoms) {
            const double F = el->formFactor(x);
            f2 += a * F * F;
            s += a * el->scatterFunction(x);
        }
        m->ff2[i] = f2 / m->sumAZ2;
        m->sf[i] = std::min(1.0, s / m->sumAZ);
    }
    // cumulative A(x_k) = int_0^{x_k^2} F^2(t)/Z^2 dt: below x_0 F^2 is flat; inside a bin use a fine
    // Simpson rule on the exact mixture form factor so the node values are accurate.
    m->ffCdf.resize(kNX);
    m->ffCdf[0] = m->ff2[0] * xNode(0) * xNode(0);
    for (uint32_t i = 1; i < kNX; ++i) {
        const double t0 = xNode(i - 1) * xNode(i - 1), t1 = xNode(i) * xNode(i);
        constexpr int NS = 16;
        double sum = 0;
        for (int j = 0; j <= NS; ++j) {
            const double t = t0 + (t1 - t0) * j / NS;
            const double x = std::sqrt(t);
            double f2 = 0;
            for (const auto& [el, a] : atoms) {
                const double F = el->formFactor(x);
                f2 += a * F * F;
            }
            const double wq = (j == 0 || j == NS) ? 1.0 : ((j & 1) ? 4.0 : 2.0);
            sum += wq * f2 / m->sumAZ2;
        }
        m->ffCdf[i] = m->ffCdf[i - 1] + sum * (t1 - t0) / NS / 3.0;
    }

    buildShells(*m, atoms);
    // photoelectric shell shares (used only in mode 2): K shells take (1 - 1/J_K) of their element's
    // photo cross section above the edge; evaluated at a representative 1.5 x binding energy.
    for (uint32_t s = 0; s < m->nShells; ++s) {
        dxb_shell& sh = m->shells[s];
        double share = 0;
        if (sh.fluor_yield > 0 || sh.binding_energy_kev > 0) {
            for (const auto& [el, a] : atoms) {
                if (std::fabs(el->edgeK - sh.binding_energy_kev) < 1e-9) {
                    const double e = std::min(kEMax, 1.5 * el->edgeK);
                    const double nPerGram = kAvogadro * m->massFraction[el->Z] / el->A;
                    const double elPhoto = nPerGram * el->photoelectric(e);
                    double tot[3];
                    m->attenuation(e, tot);
                    share = tot[0] > 0 ? elPhoto / tot[0] * (1.0 - 1.0 / el->jumpK()) : 0.0;
                }
            }
        }
        sh.photo_fraction_above = share;
    }
    return m;
}

std::shared_ptr<Material> Material::byNistName(const std::string& name)
{
    const NistEntry* e = nistFind(name);
    if (!e)
        return nullptr;
    std::map<uint32_t, double> w;
    for (const auto& p : e->w)
        w[p.first] += p.second;
    return byWeight(w);
}

std::shared_ptr<Material> Material::byChemicalFormula(const std::string& formula)
{
    // e.g. "H2O", "C5O2H8", "Ca10P6O26H2"; no parentheses
    std::map<uint32_t, double> atoms;
    size_t i = 0;
    while (i < formula.size()) {
        if (!std::isupper(static_cast<unsigned char>(formula[i])))
            return nullptr;
        std::string sym(1, formula[i++]);
        while (i < formula.size() && std::islower(static_cast<unsigned char>(formula[i])))
            sym += formula[i++];
        std::string num;
        while (i < formula.size() && (std::isdigit(static_cast<unsigned char>(formula[i])) || formula[i] == '.'))
            num += formula[i++];
        uint32_t Z = 0;
        for (uint32_t z = 1; z <= 92; ++z)
            if (sym == getElement(z)->symbol) {
                Z = z;
                break;
            }
        if (!Z)
            return nullptr;
        atoms[Z] += num.empty() ? 1.0 : std::stod(num);
    }
    std::map<uint32_t, double> w;
    for (const auto& [Z, n] : atoms)
        w[Z] = n * getElement(Z)->A;
    return byWeight(w);
}

void Material::attenuation(double e, double out[3]) const
{
    const GridPos p = energyPos(e);
    out[0] = lerpTable(photo, p);
    out[1] = lerpTable(incoh, p);
    out[2] = lerpTable(coh, p);
}
double Material::total(double e) const
{
    double a[3];
    attenuation(e, a);
    return a[0] + a[1] + a[2];
}
double Material::massEnergyTransfer(double e) const { return lerpTable(etr, energyPos(e)); }
double Material::formFactor(double x) const
{
    const double v = x <= kXMin ? ff2[0] : lerpTable(ff2, xPos(x));
    return std::sqrt(std::max(0.0, v) * sumAZ2);
}
double Material::scatterFactor(double x) const
{
    if (x <= kXMin) {
        const double r = x / kXMin;
        return sf[0] * r * r;
    }
    return lerpTable(sf, xPos(x));
}

} // namespace dxb
