// ref_driver.cpp — drives OpenDXMC's OWN sources for this boundary against include/dxmc/ + libdxmc_b200.so.
// TEST INFRASTRUCTURE (oracle/): built by oracle/Makefile.ref into oracle/_ref/opendxmc_ref from the reference's
// translation units WHERE THEY LIE under /root/reference/src/libopendxmc (nothing is copied):
//     dxmc_specialization.cpp  beamactorcontainer.cpp  datacontainer.cpp  basepipeline.cpp
//     otherphantomimportpipeline.cpp  ctsegmentationpipeline.cpp  simulationpipeline.cpp
// with the tests-only Qt / VTK stand-ins of tests/stubs/ (Qt's moc is replaced by the signal bodies below).
//
//   opendxmc_ref host
//       CPU only.  JSON lines: the app's DXBeam pose / collimation (R:dxmc_specialization.cpp:22-90), the beam outline
//       geometry BeamActorContainer::update builds from exposure(i) for all six beam types
//       (R:beamactorcontainer.cpp:104-203), the water-equivalent-diameter AEC profile of DataContainer
//       (R:datacontainer.cpp:42-100) on the reference's own PMMA cylinder (R:otherphantomimportpipeline.cpp:32-128), the
//       HU -> (material, density) segmentation of CTSegmentationPipeline (R:ctsegmentationpipeline.cpp:61-169) on a HU ramp.
//   opendxmc_ref run <mode 0|1|2> <delete_air 0|1> <histories per exposure> <out prefix>
//       Needs a GPU.  The reference's SimulationPipeline (worker<CORRECTION>, R:simulationpipeline.cpp:124-235) runs a
//       CT sequential beam on that cylinder; writes <prefix>.json (geometry, units) and raw little-endian arrays
//       <prefix>.{density,material,dose,variance,count}.bin for the Python side to rebuild the same world and compare.
#include <beamactorcontainer.hpp>
#include <ctsegmentationpipeline.hpp>
#include <datacontainer.hpp>
#include <dxmc_specialization.hpp>
#include <otherphantomimportpipeline.hpp>
#include <simulationpipeline.hpp>

#include <vtkActor.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <variant>
#include <vector>

// ---- what moc would generate: the signals.  They record what the test wants to see.
static std::shared_ptr<DataContainer> g_imported, g_simulated;
static std::atomic<int> g_running { -1 };
static std::shared_ptr<DataContainer> g_segmented;
void BasePipeline::imageDataChanged(std::shared_ptr<DataContainer> d)
{
    if (dynamic_cast<SimulationPipeline*>(this))
        g_simulated = d;
    else if (dynamic_cast<CTSegmentationPipeline*>(this))
        g_segmented = d;
    else
        g_imported = d;
}
void BasePipeline::dataProcessingStarted(ProgressWorkType) { }
void BasePipeline::dataProcessingFinished(ProgressWorkType) { }
void SimulationPipeline::simulationReady(bool) { }
void SimulationPipeline::simulationRunning(bool on) { g_running = on ? 1 : 0; }
void SimulationPipeline::simulationProgress(QString, int) { }
void OtherPhantomImportPipeline::errorMessage(QString) { }

static void printVec(const char* key, const std::array<double, 3>& v, bool comma = true)
{
    std::printf("\"%s\": [%.17g, %.17g, %.17g]%s", key, v[0], v[1], v[2], comma ? ", " : "");
}

static void dumpDx(const char* tag, const DXBeam& b)
{
    const auto& c = b.directionCosines();
    const auto& h = b.collimationHalfAngles();
    const auto coll = b.collimation();
    std::printf("{\"kind\": \"dxbeam\", \"tag\": \"%s\", ", tag);
    printVec("pos", b.position());
    printVec("c0", c[0]);
    printVec("c1", c[1]);
    std::printf("\"half\": [%.17g, %.17g], \"coll\": [%.17g, %.17g], \"prim\": %.17g, \"sec\": %.17g}\n", h[0], h[1], coll[0], coll[1],
        b.primaryAngleDeg(), b.secondaryAngleDeg());
}

static void dumpOutline(const char* tag, Beam beam)
{
    auto ptr = std::make_shared<Beam>(std::move(beam));
    BeamActorContainer c(ptr);
    c.update();
    auto actor = c.createActor();
    const vtkPolyData* pd = actor->mapperInput;
    std::printf("{\"kind\": \"outline\", \"tag\": \"%s\", \"points\": [", tag);
    for (std::size_t i = 0; i < pd->points->pts.size(); ++i) {
        const auto& p = pd->points->pts[i];
        std::printf("%s[%.17g, %.17g, %.17g]", i ? ", " : "", p[0], p[1], p[2]);
    }
    std::printf("], \"cells\": [");
    for (std::size_t i = 0; i < pd->lines->cells.size(); ++i) {
        std::printf("%s[", i ? ", " : "");
        for (std::size_t k = 0; k < pd->lines->cells[i].size(); ++k)
            std::printf("%s%lld", k ? ", " : "", pd->lines->cells[i][k]);
        std::printf("]");
    }
    std::printf("]}\n");
}

static std::shared_ptr<DataContainer> makeCylinder(double d, int n, int nz)
{
    OtherPhantomImportPipeline imp;
    imp.importPhantom(0, d, d, d, n, n, nz); // type 0 = cylinder: PMMA in air (R:otherphantomimportpipeline.cpp:61-128)
    return g_imported;
}

static int hostMode()
{
    DXBeam b;
    dumpDx("default", b);
    b.setRotationCenter({ 1.0, 2.0, 3.0 });
    b.setSourcePatientDistance(80.0);
    dumpDx("moved", b);
    b.setPrimaryAngleDeg(35.0);
    b.setSecondaryAngleDeg(-20.0);
    dumpDx("rotated", b);
    b.setSourceDetectorDistance(120.0);
    b.setCollimation({ 35.0, 43.0 });
    dumpDx("collimated", b);
    b.setPrimaryAngleDeg(400.0);
    b.setSecondaryAngleDeg(-120.0);
    dumpDx("clamped", b);

    dumpOutline("dx", b);
    CTSpiralBeam spiral({ 0, 0, -4 }, { 0, 0, 4 }, { { 13, 9.0 } });
    spiral.setStepAngleDeg(30.0);
    dumpOutline("spiral", spiral);
    CTSpiralDualEnergyBeam dual({ 0, 0, -4 }, { 0, 0, 4 }, { { 13, 9.0 } });
    dual.setStepAngleDeg(30.0);
    dumpOutline("dual", dual);
    CBCTBeam cbct({ 1, 2, 3 }, { 0, 0, 1 }, { { 13, 2.0 } });
    cbct.setStepAngleDeg(20.0);
    dumpOutline("cbct", cbct);
    CTSequentialBeam seq({ 0, 0, -2 }, { 0, 0, 1 }, { { 13, 9.0 } });
    seq.setStepAngleDeg(45.0);
    seq.setNumberOfSlices(2);
    dumpOutline("sequential", seq);
    PencilBeam pencil;
    pencil.setPosition({ 0, -30, 0 });
    pencil.setDirection({ 0, 1, 0 });
    dumpOutline("pencil", pencil);

    auto vol = makeCylinder(0.5, 40, 6);
    const auto wed = vol->calculateWaterEquivalentDiameter(true);
    const auto aec = vol->calculateAECfilterFromWaterEquivalentDiameter(true);
    std::printf("{\"kind\": \"wed\", \"dim\": [%zu, %zu, %zu], \"spacing\": [%.17g, %.17g, %.17g], \"pmma_voxels_per_slice\": %zu, "
                "\"density\": [%.17g, %.17g], \"wed\": [",
        vol->dimensions()[0], vol->dimensions()[1], vol->dimensions()[2], vol->spacing()[0], vol->spacing()[1], vol->spacing()[2],
        static_cast<std::size_t>(std::count(vol->getMaterialArray().begin(), vol->getMaterialArray().begin() + 1600, std::uint8_t { 1 })),
        dxmc::NISTMaterials::density("Air, Dry (near sea level)"), dxmc::NISTMaterials::density("Polymethyl Methacralate (Lucite, Perspex)"));
    for (std::size_t i = 0; i < wed.size(); ++i)
        std::printf("%s%.17g", i ? ", " : "", wed[i]);
    std::printf("], \"aec_weights\": [");
    const auto& w = aec.weights();
    for (std::size_t i = 0; i < w.size(); ++i)
        std::printf("%s%.17g", i ? ", " : "", w[i]);
    std::printf("], \"aec_start_z\": %.17g, \"aec_stop_z\": %.17g, \"aec_empty\": %s}\n", aec.start()[2], aec.stop()[2], aec.isEmpty() ? "true" : "false");

    // CT segmentation of a HU ramp, 120 kV with 9 mm Al (R:src/libopendxmc/ctsegmentationpipeline.cpp:114-169)
    auto ct = std::make_shared<DataContainer>();
    const std::size_t nhu = 1800;
    ct->setDimensions({ 30, 30, 2 });
    ct->setSpacing({ 0.1, 0.1, 0.1 });
    std::vector<double> hu(nhu);
    for (std::size_t i = 0; i < nhu; ++i)
        hu[i] = -1100.0 + 2.0 * static_cast<double>(i) + 0.37;
    ct->setImageArray(DataContainer::ImageType::CT, hu);
    CTSegmentationPipeline seg;
    seg.setAqusitionVoltage(120.0);
    seg.setAlFiltration(9.0);
    seg.updateImageData(ct);
    if (!g_segmented)
        return 4;
    std::printf("{\"kind\": \"segmentation\", \"hu0\": -1100.0, \"hu_step\": 2.0, \"hu_offset\": 0.37, \"n\": %zu, \"material\": \"", nhu);
    for (auto m : g_segmented->getMaterialArray())
        std::printf("%d", static_cast<int>(m));
    std::printf("\", \"density\": [");
    const auto& dd = g_segmented->getDensityArray();
    for (std::size_t i = 0; i < dd.size(); ++i)
        std::printf("%s%.17g", i ? ", " : "", dd[i]);
    std::printf("], \"materials\": [");
    for (std::size_t i = 0; i < g_segmented->getMaterials().size(); ++i)
        std::printf("%s\"%s\"", i ? ", " : "", g_segmented->getMaterials()[i].name.c_str());
    std::printf("]}\n");
    return 0;
}

template <typename T>
static void writeRaw(const std::string& path, const std::vector<T>& v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}

static int runMode(int mode, bool deleteAir, unsigned long long perExposure, const std::string& prefix)
{
    auto vol = makeCylinder(0.5, 64, 32);
    CTSequentialBeam seq({ 0, 0, 0 }, { 0, 0, 1 }, { { 13, 9.0 } });
    seq.setStepAngleDeg(10.0);
    seq.setNumberOfParticlesPerExposure(perExposure);
    auto beam = std::make_shared<Beam>(seq);
    auto actor = std::make_shared<BeamActorContainer>(beam);

    SimulationPipeline sim;
    sim.setLowEnergyCorrectionLevel(mode);
    sim.setDeleteAirDose(deleteAir);
    sim.setNumberOfThreads(0);
    sim.updateImageData(vol);
    sim.addBeamActor(actor);
    sim.startSimulation();
    if (g_running == 0) {
        std::fprintf(stderr, "ref_driver: the pipeline refused to start\n");
        return 2;
    }
    // Qt's 3 s timer, compressed: the pipeline publishes the result from timerEvent() once the worker has finished
    for (int i = 0; i < 60000 && !g_simulated; ++i) {
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        sim.timerEvent(nullptr);
    }
    if (!g_simulated) {
        std::fprintf(stderr, "ref_driver: no result\n");
        return 3;
    }
    const auto& d = *g_simulated;
    writeRaw(prefix + ".density.bin", d.getDensityArray());
    writeRaw(prefix + ".material.bin", d.getMaterialArray());
    writeRaw(prefix + ".dose.bin", d.getDoseArray());
    writeRaw(prefix + ".variance.bin", d.getDoseVarianceArray());
    writeRaw(prefix + ".count.bin", d.getDoseEventCountArray());
    std::ofstream j(prefix + ".json");
    j << "{\"dim\": [" << d.dimensions()[0] << ", " << d.dimensions()[1] << ", " << d.dimensions()[2] << "], \"spacing\": [" << d.spacing()[0] << ", "
      << d.spacing()[1] << ", " << d.spacing()[2] << "], \"mode\": " << mode << ", \"delete_air\": " << (deleteAir ? 1 : 0)
      << ", \"per_exposure\": " << perExposure << ", \"exposures\": " << seq.numberOfExposures() << ", \"dose_units\": \""
      << d.units(DataContainer::ImageType::Dose) << "\"}\n";
    return 0;
}

int main(int argc, char** argv)
{
    const std::string what = argc > 1 ? argv[1] : "host";
    if (what == "host")
        return hostMode();
    if (what == "run" && argc >= 6)
        return runMode(std::atoi(argv[2]), std::atoi(argv[3]) != 0, std::strtoull(argv[4], nullptr, 10), argv[5]);
    std::fprintf(stderr, "usage: opendxmc_ref host | run <mode> <delete_air> <histories per exposure> <out prefix>\n");
    return 1;
}
