// ref_driver.cpp — drives OpenDXMC's OWN sources for this boundary against include/dxmc/ + libdxmc_b200.so.
// TEST INFRASTRUCTURE (oracle/): built by oracle/Makefile.ref into oracle/_ref/opendxmc_ref from the reference's
// translation units WHERE THEY LIE under /root/reference/src/libopendxmc (nothing is copied):
//     dxmc_specialization.cpp  beamactorcontainer.cpp  datacontainer.cpp  basepipeline.cpp
//     otherphantomimportpipeline.cpp  icrpphantomimportpipeline.cpp  ctsegmentationpipeline.cpp  dosetablepipeline.cpp
//     beamsettingsmodel.cpp  hdf5wrapper.cpp  simulationpipeline.cpp
//     bowtiefilterreader.cpp (over the small JSON parser in tests/stubs/QJsonDocument)
// with the tests-only Qt / VTK stand-ins of tests/stubs/ (Qt's moc is replaced by the signal bodies below).
//
//   opendxmc_ref host
//       CPU only.  JSON lines: the app's DXBeam pose / collimation (R:dxmc_specialization.cpp:22-90), the beam outline
//       geometry BeamActorContainer::update builds from exposure(i) for all six beam types
//       (R:beamactorcontainer.cpp:104-203), the water-equivalent-diameter AEC profile of DataContainer
//       (R:datacontainer.cpp:42-100) on the reference's own PMMA cylinder (R:otherphantomimportpipeline.cpp:32-128), the
//       HU -> (material, density) segmentation of CTSegmentationPipeline (R:ctsegmentationpipeline.cpp:61-169) on a HU ramp.
//   opendxmc_ref icrp <organ array file> <organs.dat> <media.dat> <nx> <ny> <nz> <remove arms 0|1>
//       CPU only.  ICRPPhantomImportPipeline::importPhantom (R:icrpphantomimportpipeline.cpp:258-351) on a caller-made
//       organ array with the reference's real organ / media tables: organ, material and density arrays, names, compositions.
//   opendxmc_ref bowtie
//       CPU only, run with the reference tree as working directory.  BowtieFilterReader::read on the reference's
//       data/bowtiefilters/bowtiefilters.json (R:bowtiefilterreader.cpp:34-110): names, points, weights at sample angles.
//   opendxmc_ref beammodel
//       CPU only.  BeamSettingsModel (R:beamsettingsmodel.cpp, 1800 lines: every getter / setter of the six beam types, tube,
//       bowtie, AEC and organ-AEC the GUI offers) creates its six default beams; every (label, value) row of the settings
//       tree is printed, then a few values are edited through the model's setters and printed again.
//   opendxmc_ref h5roundtrip
//       CPU only.  HDF5Wrapper (R:hdf5wrapper.cpp, 1150 lines) saves the PMMA cylinder + its AEC profile and the five
//       savable beam types (after the edits of `beammodel`) and loads them back - against the in-memory stand-in for the
//       HDF5 C++ API in tests/stubs/H5Cpp.h, so this checks the wrapper's use of the dxmc:: accessors (radian forms,
//       filters, parseCompoundStr / AtomHandler), not the file format.  Prints the settings rows before and after.
//   opendxmc_ref dosetable <prefix> <nx> <ny> <nz> <dx> <dy> <dz> <n organs>
//       CPU only.  DoseTablePipeline::updateImageData (R:dosetablepipeline.cpp:36-95) on <prefix>.{organ,dose,density}.bin:
//       per organ the voxel count, volume, mass and dose the app's table shows.
//   opendxmc_ref run <mode 0|1|2> <delete_air 0|1> <histories per exposure> <out prefix> [level] [sequential|spiral|dual|dx|cbct|pencil]
//       Needs a GPU - or the build oracle/_ref/opendxmc_ref_cpu, which links the CPU test double of the context-level
//       C ABI (oracle/cpu_double.cpp, backed by the oracle) ahead of the library and runs anywhere.  The reference's SimulationPipeline (worker<CORRECTION>, R:simulationpipeline.cpp:124-235) runs a
//       CT sequential beam on that cylinder; writes <prefix>.json (geometry, units) and raw little-endian arrays
//       <prefix>.{density,material,dose,variance,count}.bin for the Python side to rebuild the same world and compare.
#include <beamactorcontainer.hpp>
#include <beamsettingsmodel.hpp>
#include <bowtiefilterreader.hpp>
#include <ctsegmentationpipeline.hpp>
#include <datacontainer.hpp>
#include <dosetablepipeline.hpp>
#include <dxmc_specialization.hpp>
#include <hdf5wrapper.hpp>
#include <icrpphantomimportpipeline.hpp>
#include <otherphantomimportpipeline.hpp>
#include <simulationpipeline.hpp>

#include <vtkActor.h>

#include <atomic>
#include <chrono>
#include <cmath>
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
void ICRPPhantomImportPipeline::errorMessage(QString) { }
static std::vector<std::shared_ptr<BeamActorContainer>> g_actors;
void BeamSettingsModel::beamActorAdded(std::shared_ptr<BeamActorContainer> a) { g_actors.push_back(a); }
void BeamSettingsModel::beamActorRemoved(std::shared_ptr<BeamActorContainer>) { }
void BeamSettingsModel::requestRender() { }
struct DoseRow {
    std::string name;
    int voxels = -1;
    double volume = 0, mass = 0, dose = 0;
};
static std::vector<DoseRow> g_table;
static std::vector<std::string> g_header;
void DoseTablePipeline::clearTable() { g_table.clear(); }
void DoseTablePipeline::enableSorting(bool) { }
void DoseTablePipeline::doseDataHeader(QStringList h)
{
    g_header.clear();
    for (const auto& q : h)
        g_header.push_back(q.toStdString());
}
void DoseTablePipeline::doseData(int col, int row, QVariant v)
{
    if (static_cast<std::size_t>(row) >= g_table.size())
        g_table.resize(static_cast<std::size_t>(row) + 1);
    auto& r = g_table[static_cast<std::size_t>(row)];
    if (col == 0)
        r.name = v.s.toStdString();
    else if (col == 1)
        r.voxels = v.i;
    else if (col == 2)
        r.volume = v.d;
    else if (col == 3)
        r.mass = v.d;
    else if (col == 4)
        r.dose = v.d;
}

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

static int icrpMode(char** a)
{
    ICRPPhantomImportPipeline imp;
    imp.setRemoveArms(std::atoi(a[8]) != 0);
    g_imported = nullptr;
    imp.importPhantom(QString(a[2]), QString(a[3]), QString(a[4]), 1.0, 1.0, 1.0, std::atoi(a[5]), std::atoi(a[6]), std::atoi(a[7]));
    if (!g_imported)
        return 5;
    const auto& d = *g_imported;
    std::printf("{\"kind\": \"icrp\", \"organ\": [");
    for (std::size_t i = 0; i < d.getOrganArray().size(); ++i)
        std::printf("%s%d", i ? ", " : "", static_cast<int>(d.getOrganArray()[i]));
    std::printf("], \"material\": [");
    for (std::size_t i = 0; i < d.getMaterialArray().size(); ++i)
        std::printf("%s%d", i ? ", " : "", static_cast<int>(d.getMaterialArray()[i]));
    std::printf("], \"density\": [");
    for (std::size_t i = 0; i < d.getDensityArray().size(); ++i)
        std::printf("%s%.17g", i ? ", " : "", d.getDensityArray()[i]);
    std::printf("], \"organ_names\": [");
    for (std::size_t i = 0; i < d.getOrganNames().size(); ++i)
        std::printf("%s\"%s\"", i ? ", " : "", d.getOrganNames()[i].c_str());
    std::printf("], \"materials\": [");
    for (std::size_t i = 0; i < d.getMaterials().size(); ++i) {
        std::printf("%s{\"name\": \"%s\", \"Z\": {", i ? ", " : "", d.getMaterials()[i].name.c_str());
        bool first = true;
        for (const auto& [z, w] : d.getMaterials()[i].Z) {
            std::printf("%s\"%llu\": %.17g", first ? "" : ", ", static_cast<unsigned long long>(z), w);
            first = false;
        }
        std::printf("}}");
    }
    std::printf("], \"spacing\": [%.17g, %.17g, %.17g]}\n", d.spacing()[0], d.spacing()[1], d.spacing()[2]);
    return 0;
}

static std::string jsonEscape(const std::string& in)
{
    std::string o;
    for (char ch : in) {
        if (ch == '"' || ch == '\\')
            o.push_back('\\');
        o.push_back(ch);
    }
    return o;
}

static void dumpItem(const QStandardItem* it, const std::string& path, bool& first)
{
    for (int r = 0; r < it->rowCount(); ++r) {
        const QStandardItem* label = it->child(r, 0);
        const QStandardItem* value = it->child(r, 1);
        const std::string p = path + "/" + label->data(Qt::DisplayRole).toString().toStdString();
        if (value) {
            const QVariant v = value->data(Qt::DisplayRole);
            const QVariant c = value->data(Qt::CheckStateRole);
            std::printf("%s\n  [\"%s\", ", first ? "" : ",", jsonEscape(p).c_str());
            first = false;
            if (v.kind == QVariant::Double || v.kind == QVariant::Int || v.kind == QVariant::ULongLong)
                std::printf("%.17g]", v.d);
            else if (v.kind == QVariant::String)
                std::printf("\"%s\"]", jsonEscape(v.s.toStdString()).c_str());
            else if (c.isValid())
                std::printf("%s]", c.i == Qt::Checked ? "true" : "false");
            else
                std::printf("null]");
        }
        dumpItem(label, p, first);
    }
}

// value item of the row addressed by a '/'-separated label path below `it` ("Tube B/Tube potential [kV]")
static QStandardItem* findValueItem(QStandardItem* it, const std::string& path)
{
    const auto cut = path.find('/');
    const std::string head = path.substr(0, cut);
    for (int r = 0; r < it->rowCount(); ++r) {
        QStandardItem* l = it->child(r, 0);
        if (l->data(Qt::DisplayRole).toString().toStdString() != head)
            continue;
        if (cut == std::string::npos)
            return it->child(r, 1);
        return findValueItem(l, path.substr(cut + 1));
    }
    return nullptr;
}

static void dumpModel(BeamSettingsModel& model, const char* tag, std::size_t nBeams)
{
    std::printf("{\"kind\": \"beammodel\", \"tag\": \"%s\", \"beams\": %zu, \"rows\": [", tag, nBeams);
    bool first = true;
    for (int b = 0; b < model.rowCount(); ++b) {
        const QStandardItem* root = model.item(b);
        dumpItem(root, root->data(Qt::DisplayRole).toString().toStdString(), first);
    }
    std::printf("\n]}\n");
}

// edit through the model, the way the GUI's delegate does (EditableItem::setData -> the setter lambdas)
static int applyEdits(BeamSettingsModel& model)
{
    struct Edit {
        int beam;
        const char* label;
        QVariant value;
    };
    const Edit edits[] = {
        { 0, "Tube/Tube potential [kV]", QVariant(80.0) }, { 0, "Tube/Tube Al filtration [mm]", QVariant(3.5) },
        { 0, "Collimation [cm x cm]", QVariant("30, 25") }, { 0, "Primary angle [deg]", QVariant(40.0) },
        { 0, "Source rotation center distance [cm]", QVariant(75.0) },
        { 1, "Set angle step [deg]", QVariant(2.0) }, { 1, "Set stop angle [deg]", QVariant(200.0) },
        { 2, "Photon energy [keV]", QVariant(75.0) }, { 2, "Direction normal (x, y, z)", QVariant("0, 1, 0") },
        { 3, "Pitch", QVariant(1.4) }, { 3, "Set angle step [deg]", QVariant(10.0) }, { 3, "Total collimation [cm]", QVariant(2.0) },
        { 3, "CTDIvol [mGy]", QVariant(7.5) }, { 3, "Stop position [cm]", QVariant("0, 0, 10") }, { 3, "Tube/Tube Sn filtration [mm]", QVariant(0.4) },
        { 3, "Organ AEC/Low weight [0-1]", QVariant(0.3) }, { 3, "Organ AEC/Stop angle [deg]", QVariant(120.0) },
        { 4, "Number of slices", QVariant(qulonglong { 7 }) }, { 4, "Slice spacing [cm]", QVariant(1.5) }, { 4, "CTDIw [mGy]", QVariant(3.0) },
        { 5, "Tube B/Tube potential [kV]", QVariant(140.0) }, { 5, "Tube B/Relative tube current", QVariant(2.5) }, { 5, "Pitch", QVariant(3.0) },
        { 5, "Tube B offset angle [deg]", QVariant(95.0) }, { 5, "Scan FOV Tube B [cm]", QVariant(30.0) },
    };
    int applied = 0;
    for (const auto& e : edits) {
        if (QStandardItem* it = findValueItem(model.item(e.beam), e.label)) {
            it->setData(e.value, Qt::EditRole);
            ++applied;
        } else {
            std::fprintf(stderr, "ref_driver: no row '%s' in beam %d\n", e.label, e.beam);
        }
    }
    std::fprintf(stderr, "ref_driver: %d of %zu edits applied\n", applied, sizeof(edits) / sizeof(edits[0]));
    return applied;
}

static void addDefaultBeams(BeamSettingsModel& model)
{
    model.addDXBeam();
    model.addCBCTBeam();
    model.addPencilBeam();
    model.addCTSpiralBeam();
    model.addCTSequentialBeam();
    model.addCTSpiralDualEnergyBeam();
}

static int bowtieMode()
{
    const auto filters = BowtieFilterReader::read(QString("data/bowtiefilters/bowtiefilters.json"));
    std::printf("{\"kind\": \"bowtie\", \"filters\": [");
    bool first = true;
    for (auto it = filters.cbegin(); it != filters.cend(); ++it) {
        std::printf("%s\n  {\"name\": \"%s\", \"data\": [", first ? "" : ",", jsonEscape(it->first.toStdString()).c_str());
        first = false;
        const auto data = it->second.data();
        for (std::size_t i = 0; i < data.size(); ++i)
            std::printf("%s[%.17g, %.17g]", i ? ", " : "", data[i].first, data[i].second);
        std::printf("], \"weights\": [");
        for (int k = 0; k <= 9; ++k)
            std::printf("%s%.17g", k ? ", " : "", it->second(0.05 * k));
        std::printf("]}");
    }
    std::printf("\n]}\n");
    return 0;
}

static int beamModelMode()
{
    BeamSettingsModel model;
    addDefaultBeams(model);
    dumpModel(model, "defaults", g_actors.size());
    applyEdits(model);
    dumpModel(model, "edited", g_actors.size());
    return 0;
}

static int h5RoundTripMode(const std::string& file)
{
    BeamSettingsModel model;
    addDefaultBeams(model);
    applyEdits(model);
    const auto saved = g_actors; // the six actors the model announced
    auto vol = makeCylinder(0.5, 24, 5);
    vol->setAecData(vol->calculateAECfilterFromWaterEquivalentDiameter(true));
    int nSaved = 0;
    {
        HDF5Wrapper out(file, HDF5Wrapper::FileOpenMode::WriteOver);
        if (!out.save(vol))
            return 7;
        for (const auto& a : saved)
            nSaved += out.save(a) ? 1 : 0; // the pencil beam has no save() overload (R:hdf5wrapper.cpp:1050-1068)
    }
    HDF5Wrapper in(file, HDF5Wrapper::FileOpenMode::ReadOnly);
    auto back = in.load();
    auto beams = in.loadBeams();
    if (!back)
        return 8;
    const bool sameGrid = back->dimensions() == vol->dimensions() && back->spacing() == vol->spacing() && back->getDensityArray() == vol->getDensityArray()
        && back->getMaterialArray() == vol->getMaterialArray() && back->getOrganArray() == vol->getOrganArray() && back->getOrganNames() == vol->getOrganNames();
    bool sameMaterials = back->getMaterials().size() == vol->getMaterials().size();
    double worstZ = 0;
    for (std::size_t i = 0; sameMaterials && i < vol->getMaterials().size(); ++i) {
        const auto& a = vol->getMaterials()[i];
        const auto& b = back->getMaterials()[i];
        sameMaterials = a.name == b.name && a.Z.size() == b.Z.size();
        for (const auto& [z, w] : a.Z) {
            if (!b.Z.contains(z)) {
                sameMaterials = false;
                break;
            }
            worstZ = std::max(worstZ, std::abs(b.Z.at(z) - w) / w);
        }
    }
    const bool sameAec = back->aecData().weights() == vol->aecData().weights() && back->aecData().start() == vol->aecData().start()
        && back->aecData().stop() == vol->aecData().stop();
    std::printf("{\"kind\": \"h5\", \"beams_saved\": %d, \"beams_loaded\": %zu, \"same_grid\": %s, \"same_materials\": %s, "
                "\"worst_composition_rel\": %.3g, \"same_aec\": %s}\n",
        nSaved, beams.size(), sameGrid ? "true" : "false", sameMaterials ? "true" : "false", worstZ, sameAec ? "true" : "false");
    dumpModel(model, "saved", saved.size());
    g_actors.clear();
    BeamSettingsModel model2;
    for (const auto& a : beams)
        model2.addBeam(a);
    dumpModel(model2, "loaded", beams.size());
    return 0;
}

// HDF5Wrapper::load() (the reference's loader, R:src/libopendxmc/hdf5wrapper.cpp:1070-1150) on a file written elsewhere,
// e.g. by dxb_save_dose
static int h5LoadMode(const std::string& file)
{
    HDF5Wrapper in(file, HDF5Wrapper::FileOpenMode::ReadOnly);
    auto d = in.load();
    if (!d)
        return 8;
    double dose = 0, count = 0, density = 0;
    for (double v : d->getDoseArray())
        dose += v;
    for (double v : d->getDoseEventCountArray())
        count += v;
    for (double v : d->getDensityArray())
        density += v;
    std::printf("{\"kind\": \"h5load\", \"dimensions\": [%zu, %zu, %zu], \"spacing\": [%.17g, %.17g, %.17g], \"materials\": %zu, "
                "\"dose_sum\": %.17g, \"count_sum\": %.17g, \"density_sum\": %.17g, \"variance_size\": %zu}\n",
        d->dimensions()[0], d->dimensions()[1], d->dimensions()[2], d->spacing()[0], d->spacing()[1], d->spacing()[2], d->getMaterials().size(),
        dose, count, density, d->getDoseVarianceArray().size());
    return 0;
}

template <typename T>
static std::vector<T> readRaw(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), {});
    std::vector<T> v(raw.size() / sizeof(T));
    std::copy(raw.begin(), raw.begin() + static_cast<std::ptrdiff_t>(v.size() * sizeof(T)), reinterpret_cast<char*>(v.data()));
    return v;
}

static int doseTableMode(char** a)
{
    const std::string prefix = a[2];
    auto d = std::make_shared<DataContainer>();
    d->setDimensions({ static_cast<std::size_t>(std::atoi(a[3])), static_cast<std::size_t>(std::atoi(a[4])), static_cast<std::size_t>(std::atoi(a[5])) });
    d->setSpacing({ std::atof(a[6]), std::atof(a[7]), std::atof(a[8]) });
    std::vector<std::string> names;
    for (int i = 0; i < std::atoi(a[9]); ++i)
        names.push_back("organ " + std::to_string(i));
    d->setOrganNames(names);
    if (!d->setImageArray(DataContainer::ImageType::Organ, readRaw<std::uint8_t>(prefix + ".organ.bin"))
        || !d->setImageArray(DataContainer::ImageType::Density, readRaw<double>(prefix + ".density.bin"))
        || !d->setImageArray(DataContainer::ImageType::Dose, readRaw<double>(prefix + ".dose.bin")))
        return 6;
    DoseTablePipeline table;
    table.updateImageData(d);
    std::printf("{\"kind\": \"dosetable\", \"header\": [");
    for (std::size_t i = 0; i < g_header.size(); ++i)
        std::printf("%s\"%s\"", i ? ", " : "", g_header[i].c_str());
    std::printf("], \"rows\": [");
    bool first = true;
    for (std::size_t i = 0; i < g_table.size(); ++i) {
        if (g_table[i].voxels < 0)
            continue; // organs without voxels get no row
        std::printf("%s{\"organ\": %zu, \"name\": \"%s\", \"voxels\": %d, \"volume\": %.17g, \"mass\": %.17g, \"dose\": %.17g}", first ? "" : ", ", i,
            g_table[i].name.c_str(), g_table[i].voxels, g_table[i].volume, g_table[i].mass, g_table[i].dose);
        first = false;
    }
    std::printf("]}\n");
    return 0;
}

template <typename T>
static void writeRaw(const std::string& path, const std::vector<T>& v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}

// the beam of `run`: one of six kinds with non-default settings (mirrored in tests/test_reference_sources_compile.py)
static std::shared_ptr<Beam> makeRunBeam(const std::string& kind, unsigned long long perExposure, double level, const std::shared_ptr<DataContainer>& vol,
    unsigned long long& exposures)
{
    std::shared_ptr<Beam> beam;
    if (kind == "sequential") {
        CTSequentialBeam b({ 0, 0, 0 }, { 0, 0, 1 }, { { 13, 9.0 } });
        b.setStepAngleDeg(10.0);
        b.setCTDIw(level);
        b.setNumberOfParticlesPerExposure(perExposure);
        exposures = b.numberOfExposures();
        beam = std::make_shared<Beam>(b);
    } else if (kind == "spiral") {
        CTSpiralBeam b({ 0, 0, -6 }, { 0, 0, 6 }, { { 13, 7.0 }, { 29, 0.05 } });
        b.setStepAngleDeg(15.0);
        b.setPitch(1.2);
        b.setCollimation(2.4);
        b.setScanFieldOfView(40.0);
        b.setStartAngleDeg(30.0);
        b.setTubeVoltage(100.0);
        b.setCTDIvol(level);
        b.setBowtieFilter(BowtieFilter({ { 0.0, 1.0 }, { 0.1, 0.8 }, { 0.2, 0.5 }, { 0.3, 0.25 }, { 0.39, 0.1 } }));
        b.setAECFilter(vol->calculateAECfilterFromWaterEquivalentDiameter(true));
        auto& o = b.organAECFilter();
        o.setUseFilter(true);
        o.setStartAngleDeg(300.0);
        o.setStopAngleDeg(60.0);
        o.setRampAngleDeg(25.0);
        o.setLowWeightFactor(0.4);
        b.setNumberOfParticlesPerExposure(perExposure);
        exposures = b.numberOfExposures();
        beam = std::make_shared<Beam>(b);
    } else if (kind == "dual") {
        CTSpiralDualEnergyBeam b({ 0, 0, -5 }, { 0, 0, 5 }, { { 13, 9.0 } });
        b.setStepAngleDeg(20.0);
        b.setPitch(2.0);
        b.setCollimation(3.0);
        b.setTubeAVoltage(80.0);
        b.setTubeBVoltage(140.0);
        b.addTubeBFiltrationMaterial(50, 0.4);
        b.setRelativeMasTubeB(0.6);
        b.setTubeBoffsetAngleDeg(95.0);
        b.setScanFieldOfViewB(33.0);
        b.setCTDIvol(level);
        b.setNumberOfParticlesPerExposure(perExposure);
        exposures = b.numberOfExposures();
        beam = std::make_shared<Beam>(b);
    } else if (kind == "dx") {
        DXBeam b({ { 13, 2.0 }, { 29, 0.1 } });
        b.setRotationCenter({ 0, 0, 1 });
        b.setSourcePatientDistance(60.0);
        b.setSourceDetectorDistance(110.0);
        b.setCollimation({ 20.0, 12.0 });
        b.setPrimaryAngleDeg(25.0);
        b.setSecondaryAngleDeg(-10.0);
        b.setTubeVoltage(80.0);
        b.setDAPvalue(level);
        b.setNumberOfExposures(12);
        b.setNumberOfParticlesPerExposure(perExposure);
        exposures = b.numberOfExposures();
        beam = std::make_shared<Beam>(b);
    } else if (kind == "cbct") {
        CBCTBeam b({ 0, 0, 0.5 }, { 0, 0, 1 }, { { 13, 2.5 } });
        b.setSourceDetectorDistance(90.0);
        b.setStartAngleDeg(10.0);
        b.setStopAngleDeg(250.0);
        b.setStepAngleDeg(12.0);
        b.setCollimationHalfAnglesDeg(8.0, 5.0);
        b.setDAPvalue(level);
        b.setNumberOfParticlesPerExposure(perExposure);
        exposures = b.numberOfExposures();
        beam = std::make_shared<Beam>(b);
    } else if (kind == "pencil") {
        PencilBeam b;
        b.setPosition({ 0.3, -30.0, 0.2 });
        b.setDirection({ 0, 1, 0 });
        b.setEnergy(70.0);
        b.setAirKerma(level);
        b.setNumberOfExposures(10);
        b.setNumberOfParticlesPerExposure(perExposure);
        exposures = b.numberOfExposures();
        beam = std::make_shared<Beam>(b);
    }
    return beam;
}

static int runMode(int mode, bool deleteAir, unsigned long long perExposure, const std::string& prefix, double ctdiw, const std::string& kind, bool repeat)
{
    auto vol = makeCylinder(0.5, 64, 32);
    unsigned long long nExposures = 0;
    auto beam = makeRunBeam(kind, perExposure, ctdiw, vol, nExposures);
    if (!beam) {
        std::fprintf(stderr, "ref_driver: unknown beam kind %s\n", kind.c_str());
        return 9;
    }
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
    // A second startSimulation() on the SAME pipeline object: the worker leaves the stop flag of the pipeline's single
    // TransportProgress raised (R:src/libopendxmc/simulationpipeline.cpp:234) and nothing in OpenDXMC clears it, so
    // dxmc::Transport must (progress->start()).  The second run builds a new world, hence draws the same Philox keys:
    // its result must arrive and equal the first one.
    int secondRun = -1;
    if (repeat) {
        const std::vector<double> firstDose = g_simulated->getDoseArray();
        g_simulated.reset();
        g_running = -1;
        sim.startSimulation();
        for (int i = 0; i < 60000 && !g_simulated; ++i) {
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
            sim.timerEvent(nullptr);
        }
        if (!g_simulated) {
            std::fprintf(stderr, "ref_driver: the second startSimulation() on the same pipeline produced no result\n");
            return 4;
        }
        // (the CPU double adds doubles from several threads: last bits move with the thread order)
        const std::vector<double>& second = g_simulated->getDoseArray();
        double sum = 0, worst = 0, top = 0;
        for (std::size_t i = 0; i < second.size() && i < firstDose.size(); ++i) {
            sum += second[i];
            top = std::max(top, std::fabs(firstDose[i]));
            worst = std::max(worst, std::fabs(second[i] - firstDose[i]));
        }
        secondRun = (second.size() == firstDose.size() && sum > 0 && worst <= 1e-9 * top) ? 1 : 0;
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
      << ", \"ctdiw\": " << ctdiw << ", \"per_exposure\": " << perExposure << ", \"exposures\": " << nExposures << ", \"beam\": \"" << kind << "\", \"second_run_identical\": " << secondRun << ", \"dose_units\": \""
      << d.units(DataContainer::ImageType::Dose) << "\"}\n";
    return 0;
}

int main(int argc, char** argv)
{
    const std::string what = argc > 1 ? argv[1] : "host";
    if (what == "host")
        return hostMode();
    if (what == "bowtie")
        return bowtieMode();
    if (what == "beammodel")
        return beamModelMode();
    if (what == "h5roundtrip")
        return h5RoundTripMode(argc > 2 ? argv[2] : "/tmp/opendxmc_ref_roundtrip.h5");
    if (what == "h5load" && argc >= 3)
        return h5LoadMode(argv[2]);
    if (what == "dosetable" && argc >= 10)
        return doseTableMode(argv);
    if (what == "icrp" && argc >= 9)
        return icrpMode(argv);
    if (what == "run" && argc >= 6)
        return runMode(std::atoi(argv[2]), std::atoi(argv[3]) != 0, std::strtoull(argv[4], nullptr, 10), argv[5], argc > 6 ? std::atof(argv[6]) : 1.0, argc > 7 ? argv[7] : "sequential", argc > 8 && std::atoi(argv[8]) != 0);
    std::fprintf(stderr, "usage: opendxmc_ref host | bowtie | beammodel | h5roundtrip [file] | h5load <file> | icrp <organ array> <organs.dat> <media.dat> <nx> <ny> <nz> <remove arms> | dosetable <prefix> <nx> <ny> <nz> <dx> <dy> <dz> <n organs> | run <mode> <delete_air> <histories per exposure> <out prefix> [level] [beam kind] [1: start twice]\n");
    return 1;
}
