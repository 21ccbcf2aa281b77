"""The C++ shim headers (include/dxmc/) compile stand-alone, each one, as C++20 — the include paths and names are the
ones OpenDXMC uses (R:src/libopendxmc/dxmc_specialization.hpp:21-30, simulationpipeline.cpp:23-25)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [
    "dxmc/transport.hpp", "dxmc/world/world.hpp", "dxmc/world/worlditems/aavoxelgrid.hpp", "dxmc/transportprogress.hpp",
    "dxmc/beams/beamtype.hpp", "dxmc/beams/cbctbeam.hpp", "dxmc/beams/ctsequentialbeam.hpp", "dxmc/beams/ctspiralbeam.hpp",
    "dxmc/beams/ctspiraldualenergybeam.hpp", "dxmc/beams/dxbeam.hpp", "dxmc/beams/pencilbeam.hpp", "dxmc/beams/tube/tube.hpp",
    "dxmc/material/material.hpp", "dxmc/material/nistmaterials.hpp", "dxmc/material/atomhandler.hpp", "dxmc/vectormath.hpp",
    "dxmc/constants.hpp",
]


@pytest.mark.parametrize("hdr", HEADERS)
def test_header_compiles_standalone(hdr, tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(f"#include <{hdr}>\nint main() {{ return 0; }}\n")
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT}/include", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_host_objects_work_without_gpu(tmp_path):
    """materials, tube, beams and exposures are host code: usable on a CPU-only box through the shim."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <dxmc/beams/ctspiralbeam.hpp>
#include <dxmc/beams/ctspiraldualenergybeam.hpp>
#include <dxmc/beams/dxbeam.hpp>
#include <dxmc/material/material.hpp>
#include <dxmc/material/atomhandler.hpp>
#include <cstdio>
#include <cmath>
int main() {
    auto w = dxmc::Material<5>::byNistName("Water, Liquid");
    if (!w) return 1;
    if (dxmc::Material<5>::byWeight({{120, 1.0}})) return 2;
    const double mu = w->attenuationValues(60.0).sum();
    if (std::abs(mu - 0.2059) > 0.01) return 3;
    dxmc::CTSpiralBeam<false> ct({0, 0, -15}, {0, 0, 15}, {{13, 9.0}});
    if (ct.numberOfExposures() != 563) return 4;               // ceil(30 / 3.84 * 360 / 5)
    const auto e = ct.exposure(0);
    if (std::abs(std::hypot(e.position()[0], e.position()[1]) - 59.5) > 1e-9) return 5;
    if (ct.tube().filtration(13) != 9.0 || ct.collimation() != 3.84) return 6;
    auto comp = dxmc::Material<5>::parseCompoundStr("H2O");
    if (comp[1] != 2 || comp[8] != 1) return 7;
    if (dxmc::AtomHandler::toSymbol(20) != "Ca") return 8;
    dxmc::CTSpiralDualEnergyBeam<false> d({0, 0, -15}, {0, 0, 15}, {{13, 9.0}});
    d.setRelativeMasTubeB(3.0);
    if (std::abs(d.tubeRelativeWeightA() + d.tubeRelativeWeightB() - 2.0) > 1e-9) return 9;
    std::printf("ok %.5f\n", mu);
    return 0;
}
''')
    exe = tmp_path / "t"
    r = subprocess.run(["g++", "-std=c++20", "-O1", f"-I{ROOT}/include", str(src), "-o", str(exe), f"-L{ROOT}/opendxmc_b200/lib", "-ldxmc_b200",
                        f"-Wl,-rpath,{ROOT}/opendxmc_b200/lib"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
