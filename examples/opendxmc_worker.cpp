// examples/opendxmc_worker.cpp — OpenDXMC's simulation driver retyped against the shim headers.
//
// `worker<CORRECTION>()` below is the body of R:src/libopendxmc/simulationpipeline.cpp:124-235 with the Qt/VTK
// DataContainer replaced by a plain struct of the same arrays (R:src/libopendxmc/datacontainer.hpp:98-105) and
// std::execution::par_unseq dropped (needs TBB).  Every dxmc:: call is spelled as in the reference, which is the point:
// it compiles and runs against include/dxmc/ + libdxmc_b200.so.  `main` builds the C1 phantom the way
// R:src/libopendxmc/otherphantomimportpipeline.cpp:32-110 does and runs a DXBeam subclass as in
// R:src/libopendxmc/dxmc_specialization.cpp:22-90 plus a CT spiral beam with the GUI defaults
// (R:src/libopendxmc/beamsettingsmodel.cpp:1152-1171).
#include <dxmc/transport.hpp>
#include <dxmc/world/world.hpp>
#include <dxmc/world/worlditems/aavoxelgrid.hpp>

#include "dxmc/beams/beamtype.hpp"
#include "dxmc/beams/cbctbeam.hpp"
#include "dxmc/beams/ctsequentialbeam.hpp"
#include "dxmc/beams/ctspiralbeam.hpp"
#include "dxmc/beams/ctspiraldualenergybeam.hpp"
#include "dxmc/beams/dxbeam.hpp"
#include "dxmc/beams/pencilbeam.hpp"
#include "dxmc/beams/tube/tube.hpp"
#include "dxmc/material/material.hpp"
#include "dxmc/material/nistmaterials.hpp"
#include "dxmc/transportprogress.hpp"

#include <algorithm>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <variant>
#include <vector>

// ---- R:src/libopendxmc/dxmc_specialization.hpp:35-77
using Material = dxmc::Material<5>;
using Tube = dxmc::Tube;
using NISTMaterials = dxmc::NISTMaterials;
using CTSequentialBeam = dxmc::CTSequentialBeam<false>;
using CTSpiralBeam = dxmc::CTSpiralBeam<false>;
using CTSpiralDualEnergyBeam = dxmc::CTSpiralDualEnergyBeam<false>;
using CBCTBeam = dxmc::CBCTBeam<false>;
using CTAECFilter = dxmc::CTAECFilter;
using BowtieFilter = dxmc::BowtieFilter;
using PencilBeam = dxmc::PencilBeam<false>;

class DXBeam : public dxmc::DXBeam<false> {
public:
    DXBeam(const std::map<std::size_t, double>& filtrationMaterials = {})
        : dxmc::DXBeam<false>({ 0, 0, 0 }, { { { 1, 0, 0 }, { 0, -1, 0 } } }, filtrationMaterials)
    {
        updatePosition();
        setCollimation({ 20, 20 });
    }
    void setRotationCenter(const std::array<double, 3>& c)
    {
        m_rotation_center = c;
        updatePosition();
    }
    void setSourcePatientDistance(double d)
    {
        m_SPD = std::abs(d);
        updatePosition();
    }
    void setCollimation(const std::array<double, 2>& coll)
    {
        const std::array<double, 2> ang = { std::tan(0.5 * std::abs(coll[0]) / m_SDD), std::tan(0.5 * std::abs(coll[1]) / m_SDD) };
        setCollimationHalfAngles(ang);
    }
    void setPrimaryAngleDeg(double ang)
    {
        m_rotAngles[0] = dxmc::DEG_TO_RAD<double>() * std::clamp(ang, -180.0, 180.0);
        updatePosition();
    }

protected:
    void updatePosition()
    {
        std::array<std::array<double, 3>, 2> cosines = { { { 0, 0, 1 }, { -1, 0, 0 } } };
        cosines[0] = dxmc::vectormath::rotate(cosines[0], { 0, 0, 1 }, m_rotAngles[0]);
        cosines[1] = dxmc::vectormath::rotate(cosines[1], { 0, 0, 1 }, m_rotAngles[0]);
        cosines[0] = dxmc::vectormath::rotate(cosines[0], { -1, 0, 0 }, m_rotAngles[1]);
        cosines[1] = dxmc::vectormath::rotate(cosines[1], { -1, 0, 0 }, m_rotAngles[1]);
        auto dir = dxmc::vectormath::cross(cosines[0], cosines[1]);
        auto ddist = dxmc::vectormath::scale(dir, -m_SPD);
        setPosition(dxmc::vectormath::add(m_rotation_center, ddist));
        setDirectionCosines(cosines);
    }

private:
    std::array<double, 3> m_rotation_center = { 0, 0, 0 };
    double m_SPD = 100, m_SDD = 100;
    std::array<double, 2> m_rotAngles = { 0, 0 };
};

using Beam = std::variant<DXBeam, CTSpiralBeam, CTSpiralDualEnergyBeam, CBCTBeam, CTSequentialBeam, PencilBeam>;

// ---- the arrays of DataContainer that the driver touches
struct DataContainer {
    struct MaterialTemplate {
        std::string name;
        std::map<std::size_t, double> Z;
    };
    std::array<std::size_t, 3> m_dimensions { 0, 0, 0 };
    std::array<double, 3> m_spacing { 1, 1, 1 }; // cm
    std::vector<double> density, dose, doseVariance, doseCount;
    std::vector<std::uint8_t> material;
    std::vector<MaterialTemplate> materials;
    std::string doseUnits;
    const std::array<std::size_t, 3>& dimensions() const { return m_dimensions; }
    const std::array<double, 3>& spacing() const { return m_spacing; }
    const std::vector<double>& getDensityArray() const { return density; }
    const std::vector<std::uint8_t>& getMaterialArray() const { return material; }
    const std::vector<MaterialTemplate>& getMaterials() const { return materials; }
};

// ---- R:src/libopendxmc/simulationpipeline.cpp:124-235
template <int CORRECTION = 1>
void worker(bool deleteAirDose, int nthreads, std::shared_ptr<DataContainer> data, std::vector<std::shared_ptr<Beam>> beams, dxmc::TransportProgress* progress)
{
    using VoxelGrid = dxmc::AAVoxelGrid<5, CORRECTION, 255>;
    using World = dxmc::World<VoxelGrid>;

    World world;
    auto& vgrid = world.template addItem<VoxelGrid>();
    {
        std::vector<Material> materials;
        for (const auto& materialTemplate : data->getMaterials()) {
            auto material = Material::byWeight(materialTemplate.Z);
            if (!material) {
                progress->setStopSimulation();
                return;
            } else {
                materials.push_back(material.value());
            }
        }
        const auto dims = data->dimensions();
        const auto spacing = data->spacing();
        const auto& densityArray = data->getDensityArray();
        const auto& materialArray = data->getMaterialArray();
        vgrid.setData(dims, densityArray, materialArray, materials);
        vgrid.setSpacing(spacing);
    }
    world.build();

    dxmc::Transport transport;
    if (nthreads > 0)
        transport.setNumberOfThreads(nthreads);

    const int Njobs = beams.size();
    for (int jobIdx = 0; jobIdx < Njobs; jobIdx++) {
        const auto& currentbeam = *(beams[jobIdx]);
        std::visit([&](auto&& beam) { transport(world, beam, progress, true); }, currentbeam);
        if (!progress->continueSimulation())
            return;
    }

    const auto N = vgrid.size();
    {
        std::vector<double> dose(N);
        for (std::size_t i = 0; i < N; ++i)
            dose[i] = vgrid.doseScored(i).dose();
        if (deleteAirDose) {
            const auto& matarr = data->getMaterialArray();
            std::transform(dose.cbegin(), dose.cend(), matarr.cbegin(), dose.begin(), [](const auto d, const auto m) { return m > 0 ? d : 0.0; });
        }
        const auto& max_idx = std::max_element(dose.cbegin(), dose.cend());
        if (*max_idx < 1) {
            std::transform(dose.cbegin(), dose.cend(), dose.begin(), [](const auto d) { return d * 1e3; });
            data->doseUnits = "uGy";
        } else {
            data->doseUnits = "mGy";
        }
        data->dose = dose;
    }
    {
        std::vector<double> dose_count_array(N, 0);
        for (std::size_t i = 0; i < N; ++i)
            dose_count_array[i] = static_cast<double>(vgrid.doseScored(i).numberOfEvents());
        if (deleteAirDose) {
            const auto& matarr = data->getMaterialArray();
            std::transform(dose_count_array.cbegin(), dose_count_array.cend(), matarr.cbegin(), dose_count_array.begin(), [](const auto d, const auto m) { return m > 0 ? d : 0; });
        }
        data->doseCount = dose_count_array;
    }
    {
        std::vector<double> dose_var(N, 0.0);
        for (std::size_t i = 0; i < N; ++i)
            dose_var[i] = vgrid.doseScored(i).variance();
        if (deleteAirDose) {
            const auto& matarr = data->getMaterialArray();
            std::transform(dose_var.cbegin(), dose_var.cend(), matarr.cbegin(), dose_var.begin(), [](const auto d, const auto m) { return m > 0 ? d : 0.0; });
        }
        if (data->doseUnits[0] == 'u')
            std::for_each(dose_var.begin(), dose_var.end(), [](auto& v) { v *= 1e6; });
        data->doseVariance = dose_var;
    }
    progress->setStopSimulation();
}

// ---- R:src/libopendxmc/otherphantomimportpipeline.cpp:32-110 (cylinder phantom)
static std::shared_ptr<DataContainer> cylinderPhantom(std::size_t n, double spacing)
{
    auto vol = std::make_shared<DataContainer>();
    vol->m_dimensions = { n, n, n };
    vol->m_spacing = { spacing, spacing, spacing };
    const std::size_t N = n * n * n;
    vol->material.resize(N);
    const auto cx = n / 2.0, cy = n / 2.0;
    const auto r = std::min(cx, cy) * 16.0 / (0.5 * n * spacing); // r = 16 cm
    for (std::size_t k = 0; k < n; ++k)
        for (std::size_t j = 0; j < n; ++j)
            for (std::size_t i = 0; i < n; ++i) {
                const auto x = i - cx, y = j - cy;
                vol->material[i + j * n + k * n * n] = x * x + y * y <= r * r ? 1 : 0;
            }
    const std::vector<std::string> names = { "Air, Dry (near sea level)", "Polymethyl Methacralate (Lucite, Perspex)" };
    for (const auto& nm : names)
        vol->materials.push_back({ nm, dxmc::NISTMaterials::Composition(nm) });
    const double air_dens = dxmc::NISTMaterials::density(names[0]), pmma_dens = dxmc::NISTMaterials::density(names[1]);
    vol->density.resize(N);
    std::transform(vol->material.cbegin(), vol->material.cend(), vol->density.begin(), [=](const auto m) { return m == 1 ? pmma_dens : air_dens; });
    return vol;
}

int main(int argc, char** argv)
{
    const std::uint64_t histories = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 2000000ull;
    auto data = cylinderPhantom(64, 36.0 / 64);

    std::vector<std::shared_ptr<Beam>> beams;
    {
        CTSpiralBeam ct({ 0, 0, -10 }, { 0, 0, 10 }, { { 13, 9.0 } });
        ct.setTubeVoltage(120);
        ct.setStepAngleDeg(5);
        ct.setCTDIvol(10.0);
        ct.setNumberOfParticlesPerExposure(std::max<std::uint64_t>(1, histories / ct.numberOfExposures()));
        beams.push_back(std::make_shared<Beam>(ct));
    }
    {
        DXBeam dx({ { 13, 2.0 }, { 29, 0.1 } });
        dx.setTubeVoltage(80);
        dx.setRotationCenter({ 0, 0, 0 });
        dx.setSourcePatientDistance(100);
        dx.setCollimation({ 30, 30 });
        dx.setDAPvalue(1.0);
        dx.setNumberOfExposures(16);
        dx.setNumberOfParticlesPerExposure(histories / 16);
        beams.push_back(std::make_shared<Beam>(dx));
    }
    dxmc::TransportProgress progress;
    try {
        worker<1>(true, 0, data, beams, &progress);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "worker failed: %s\n", e.what());
        return 2;
    }
    if (data->dose.empty()) {
        std::fprintf(stderr, "no dose produced\n");
        return 1;
    }
    double sum = 0, mx = 0, ev = 0;
    for (std::size_t i = 0; i < data->dose.size(); ++i) {
        sum += data->dose[i];
        mx = std::max(mx, data->dose[i]);
        ev += data->doseCount[i];
    }
    const auto [done, total] = progress.progress();
    std::printf("opendxmc_worker: voxels=%zu units=%s mean_dose=%.6g max_dose=%.6g events=%.0f progress=%llu/%llu finished=%d\n", data->dose.size(),
        data->doseUnits.c_str(), sum / data->dose.size(), mx, ev, (unsigned long long)done, (unsigned long long)total, !progress.continueSimulation());
    return (sum > 0 && ev > 0) ? 0 : 1;
}
