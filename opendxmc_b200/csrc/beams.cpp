// beams.cpp — host-side beam model: filters (bowtie, AEC, organ AEC) and the expansion of a
// dxb_beam_desc into exposures {position, direction cosines, half angles, weight}.
//
// Verified from the reference (authoritative):
//   photon direction from (cosines, half-angles)       R:src/libopendxmc/beamactorcontainer.cpp:44-69
//   dual-source interleave exposure(2i)=A, (2i+1)=B     R:src/libopendxmc/beamactorcontainer.cpp:134-146
//   CT defaults (SDD 119, collimation 3.84, 9 mm Al)    R:src/libopendxmc/beamsettingsmodel.cpp:1152-1171
//   bowtie data = unsorted (angle, weight) pairs        R:src/libopendxmc/bowtiefilterreader.cpp:74-93
//   AEC = (start, stop, weights)                        R:src/libopendxmc/datacontainer.cpp:37,59
// Recalled DXMClib design intent (unverified): gantry geometry, exposure
// count, weights.
#include "physics.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace dxb {

// ---------------------------------------------------------------- bowtie
BowtieTable makeBowtie(const dxb_bowtie& b)
{
    BowtieTable t;
    if (b.n < 2 || !b.angle_rad || !b.weight)
        return t;
    std::vector<std::pair<double, double>> d;
    for (uint32_t i = 0; i < b.n; ++i)
        d.emplace_back(std::fabs(b.angle_rad[i]), b.weight[i]);
    std::sort(d.begin(), d.end());
    // merge duplicate angles
    for (const auto& p : d) {
        if (!t.angle.empty() && p.first - t.angle.back() < 1e-12) {
            t.weight.back() = 0.5 * (t.weight.back() + p.second);
        } else {
            t.angle.push_back(p.first);
            t.weight.push_back(std::max(0.0, p.second));
        }
    }
    if (t.angle.size() < 2) {
        t.angle.clear();
        t.weight.clear();
        return t;
    }
    // normalise: mean of the piecewise-linear profile over [0, angle_max] is 1
    double area = t.weight.front() * t.angle.front(); // flat from 0 to the first knot
    for (size_t i = 1; i < t.angle.size(); ++i)
        area += 0.5 * (t.weight[i] + t.weight[i - 1]) * (t.angle[i] - t.angle[i - 1]);
    const double mean = area / t.angle.back();
    if (mean > 0)
        for (auto& w : t.weight)
            w /= mean;
    return t;
}

double BowtieTable::operator()(double a) const
{
    if (empty())
        return 1.0;
    a = std::fabs(a);
    if (a <= angle.front())
        return weight.front();
    if (a >= angle.back())
        return weight.back();
    size_t i = 1;
    while (angle[i] < a)
        ++i;
    const double f = (a - angle[i - 1]) / (angle[i] - angle[i - 1]);
    return weight[i - 1] + f * (weight[i] - weight[i - 1]);
}

// ---------------------------------------------------------------- CT AEC
AecTable makeAec(const dxb_aec& a)
{
    AecTable t;
    if (a.n < 2 || !a.weights)
        return t;
    double d[3] = { a.stop[0] - a.start[0], a.stop[1] - a.start[1], a.stop[2] - a.start[2] };
    const double len = vm::norm(d);
    if (!(len > 0))
        return t;
    t.length = len;
    for (int i = 0; i < 3; ++i) {
        t.start[i] = a.start[i];
        t.dir[i] = d[i] / len;
    }
    t.weights.assign(a.weights, a.weights + a.n);
    double mean = 0;
    for (double w : t.weights)
        mean += w;
    mean /= t.weights.size();
    if (mean > 0)
        for (auto& w : t.weights)
            w /= mean;
    return t;
}

double AecTable::operator()(const double pos[3]) const
{
    if (empty())
        return 1.0;
    const double rel[3] = { pos[0] - start[0], pos[1] - start[1], pos[2] - start[2] };
    double u = vm::dot(rel, dir) / length;
    u = std::clamp(u, 0.0, 1.0) * (weights.size() - 1);
    size_t i = static_cast<size_t>(u);
    if (i >= weights.size() - 1)
        return weights.back();
    const double f = u - i;
    return weights[i] + f * (weights[i + 1] - weights[i]);
}

// ---------------------------------------------------------------- organ AEC
namespace {
double wrap2pi(double a)
{
    const double twoPi = 2.0 * kPi;
    a = std::fmod(a, twoPi);
    if (a < 0)
        a += twoPi;
    return a;
}
}

double organAecMaxWeight(const dxb_organ_aec& o)
{
    if (!o.use_filter || !o.compensate_outside)
        return 1.0;
    const double twoPi = 2.0 * kPi;
    const double low = std::clamp(o.low_weight, 0.0, 1.0);
    double L = wrap2pi(o.stop_angle - o.start_angle);
    double r = std::max(0.0, o.ramp_angle);
    r = std::min(r, 0.5 * (twoPi - L));
    const double den = twoPi - L - r;
    if (!(den > 1e-9))
        return 1.0;
    return (twoPi - low * (L + r)) / den;
}

double organAecWeight(const dxb_organ_aec& o, double angle)
{
    if (!o.use_filter)
        return 1.0;
    const double twoPi = 2.0 * kPi;
    const double low = std::clamp(o.low_weight, 0.0, 1.0);
    const double high = organAecMaxWeight(o);
    const double L = wrap2pi(o.stop_angle - o.start_angle);
    double r = std::max(0.0, o.ramp_angle);
    r = std::min(r, 0.5 * (twoPi - L));
    // angle measured from the start of the low-weight sector
    const double a = wrap2pi(angle - o.start_angle);
    if (a <= L)
        return low;
    if (r > 0 && a < L + r) // ramp up after the sector
        return low + (high - low) * (a - L) / r;
    if (r > 0 && a > twoPi - r) // ramp down before the sector
        return low + (high - low) * (twoPi - a) / r;
    return high;
}

// ---------------------------------------------------------------- beams
namespace {

uint64_t ceilCount(double x)
{
    if (!(x > 0) || !std::isfinite(x))
        return 1;
    return std::max<uint64_t>(1, static_cast<uint64_t>(std::ceil(x - 1e-9)));
}

bool isDual(const dxb_beam_desc& b)
{
    return b.type == DXB_BEAM_CT_SPIRAL_DUAL || (b.type == DXB_BEAM_CTDI && b.spectrum[1].n > 0);
}

// gantry frame: fan direction, axis direction and source position for a rotation `angle` about
// `axis` through `center`
void gantry(const double center[3], const double axisIn[3], double angle, double sdd, dxb_exposure& e)
{
    double axis[3] = { axisIn[0], axisIn[1], axisIn[2] };
    vm::normalize(axis);
    double unit[3] = { 0, 0, 0 };
    unit[vm::argminAbs(axis)] = 1.0;
    double n0[3];
    vm::cross(unit, axis, n0);
    vm::normalize(n0);
    double fan[3];
    vm::rotate(n0, axis, angle, fan);
    for (int i = 0; i < 3; ++i) {
        e.cosines[0][i] = fan[i];
        e.cosines[1][i] = axis[i];
    }
    vm::cross(e.cosines[0], e.cosines[1], e.direction);
    for (int i = 0; i < 3; ++i)
        e.position[i] = center[i] - e.direction[i] * sdd * 0.5;
}

void tubeWeights(const dxb_beam_desc& b, double& wa, double& wb)
{
    auto total = [](const dxb_spectrum& s) {
        double t = 0;
        for (uint32_t i = 0; i < s.n; ++i)
            t += s.weight ? s.weight[i] : 0.0;
        return t;
    };
    const double ma = b.relative_mas_a > 0 ? b.relative_mas_a : 1.0;
    const double mb = b.relative_mas_b > 0 ? b.relative_mas_b : 1.0;
    double a = ma * total(b.spectrum[0]);
    double c = mb * total(b.spectrum[1]);
    if (!(a > 0) || !(c > 0)) {
        a = ma;
        c = mb;
    }
    const double mean = 0.5 * (a + c);
    wa = a / mean;
    wb = c / mean;
}

} // namespace

uint64_t beamNumberOfExposures(const dxb_beam_desc& b)
{
    const double step = std::fabs(b.step_angle) > 0 ? std::fabs(b.step_angle) : kPi / 180.0;
    switch (b.type) {
    case DXB_BEAM_DX:
    case DXB_BEAM_PENCIL:
        return std::max<uint64_t>(1, b.n_exposures);
    case DXB_BEAM_CT_SPIRAL:
    case DXB_BEAM_CT_SPIRAL_DUAL: {
        const double d[3] = { b.stop[0] - b.start[0], b.stop[1] - b.start[1], b.stop[2] - b.start[2] };
        const double feed = std::fabs(b.pitch * b.collimation);
        const double totalAngle = feed > 0 ? vm::norm(d) / feed * 2.0 * kPi : 2.0 * kPi;
        const uint64_t n = ceilCount(totalAngle / step);
        return b.type == DXB_BEAM_CT_SPIRAL_DUAL ? 2 * n : n;
    }
    case DXB_BEAM_CBCT:
        return ceilCount(std::fabs(b.stop_angle - b.start_angle) / step);
    case DXB_BEAM_CT_SEQUENTIAL:
        return std::max<uint64_t>(1, b.n_slices) * ceilCount(2.0 * kPi / step);
    case DXB_BEAM_CTDI:
        return ceilCount(2.0 * kPi / step) * (isDual(b) ? 2 : 1);
    default:
        return 0;
    }
}

int beamExposure(const dxb_beam_desc& b, uint64_t index, const AecTable& aec, dxb_exposure& e)
{
    std::memset(&e, 0, sizeof(e));
    const uint64_t N = beamNumberOfExposures(b);
    if (index >= N)
        return DXB_EINVAL;
    e.n_particles = b.particles_per_exposure;
    e.weight = 1.0;
    const double step = std::fabs(b.step_angle) > 0 ? std::fabs(b.step_angle) : kPi / 180.0;

    switch (b.type) {
    case DXB_BEAM_DX: {
        for (int i = 0; i < 3; ++i) {
            e.position[i] = b.position[i];
            e.cosines[0][i] = b.cosines[0][i];
            e.cosines[1][i] = b.cosines[1][i];
        }
        vm::normalize(e.cosines[0]);
        vm::normalize(e.cosines[1]);
        vm::cross(e.cosines[0], e.cosines[1], e.direction);
        e.half_angles[0] = std::fabs(b.half_angles[0]);
        e.half_angles[1] = std::fabs(b.half_angles[1]);
        return DXB_OK;
    }
    case DXB_BEAM_PENCIL: {
        double d[3] = { b.direction[0], b.direction[1], b.direction[2] };
        if (!(vm::norm(d) > 0))
            d[2] = 1.0;
        vm::normalize(d);
        double unit[3] = { 0, 0, 0 };
        unit[vm::argminAbs(d)] = 1.0;
        vm::cross(unit, d, e.cosines[0]);
        vm::normalize(e.cosines[0]);
        vm::cross(d, e.cosines[0], e.cosines[1]);
        vm::cross(e.cosines[0], e.cosines[1], e.direction);
        for (int i = 0; i < 3; ++i)
            e.position[i] = b.position[i];
        return DXB_OK;
    }
    case DXB_BEAM_CT_SPIRAL:
    case DXB_BEAM_CT_SPIRAL_DUAL: {
        const bool dual = b.type == DXB_BEAM_CT_SPIRAL_DUAL;
        const uint64_t k = dual ? index / 2 : index;
        const int tube = dual ? static_cast<int>(index & 1) : 0;
        double axis[3] = { b.stop[0] - b.start[0], b.stop[1] - b.start[1], b.stop[2] - b.start[2] };
        if (!(vm::norm(axis) > 0))
            axis[2] = 1.0;
        vm::normalize(axis);
        const double rot = static_cast<double>(k) * step;
        const double feed = b.pitch * b.collimation * rot / (2.0 * kPi);
        double center[3];
        for (int i = 0; i < 3; ++i)
            center[i] = b.start[i] + axis[i] * feed;
        const double angle = b.start_angle + rot + (tube ? b.tube_b_offset_angle : 0.0);
        gantry(center, axis, angle, b.sdd, e);
        const double fov = tube ? b.fov_b : b.fov;
        e.half_angles[0] = std::atan(fov / b.sdd);
        e.half_angles[1] = std::atan(b.collimation / b.sdd);
        e.tube = tube;
        e.weight = aec(center) * organAecWeight(b.organ_aec, angle);
        if (dual) {
            double wa, wb;
            tubeWeights(b, wa, wb);
            e.weight *= tube ? wb : wa;
        }
        return DXB_OK;
    }
    case DXB_BEAM_CT_SEQUENTIAL:
    case DXB_BEAM_CTDI: {
        const bool dual = isDual(b);
        const uint64_t perRot = ceilCount(2.0 * kPi / step);
        const uint64_t k = dual ? index / 2 : index;
        const int tube = dual ? static_cast<int>(index & 1) : 0;
        const uint64_t slice = k / perRot;
        const uint64_t a = k % perRot;
        double axis[3] = { b.direction[0], b.direction[1], b.direction[2] };
        if (!(vm::norm(axis) > 0))
            axis[2] = 1.0;
        vm::normalize(axis);
        double center[3];
        for (int i = 0; i < 3; ++i)
            center[i] = b.position[i] + axis[i] * b.slice_spacing * static_cast<double>(slice);
        const double angle = b.start_angle + static_cast<double>(a) * step + (tube ? b.tube_b_offset_angle : 0.0);
        gantry(center, axis, angle, b.sdd, e);
        const double fov = tube ? b.fov_b : b.fov;
        e.half_angles[0] = std::atan(fov / b.sdd);
        e.half_angles[1] = std::atan(b.collimation / b.sdd);
        e.tube = tube;
        e.weight = b.type == DXB_BEAM_CTDI ? 1.0 : organAecWeight(b.organ_aec, angle);
        if (dual) {
            double wa, wb;
            tubeWeights(b, wa, wb);
            e.weight *= tube ? wb : wa;
        }
        return DXB_OK;
    }
    case DXB_BEAM_CBCT: {
        const double sign = b.stop_angle >= b.start_angle ? 1.0 : -1.0;
        const double angle = b.start_angle + sign * static_cast<double>(index) * step;
        double axis[3] = { b.direction[0], b.direction[1], b.direction[2] };
        if (!(vm::norm(axis) > 0))
            axis[2] = 1.0;
        gantry(b.isocenter, axis, angle, b.sdd, e);
        e.half_angles[0] = std::fabs(b.half_angles[0]);
        e.half_angles[1] = std::fabs(b.half_angles[1]);
        return DXB_OK;
    }
    default:
        return DXB_EINVAL;
    }
}

double beamAnalyticCalibration(const dxb_beam_desc& b)
{
    // mGy per (keV/g) for the beams that are calibrated without a nested simulation.
    const double nTotal = static_cast<double>(beamNumberOfExposures(b)) * static_cast<double>(b.particles_per_exposure);
    if (!(nTotal > 0))
        return kKeVperGramToMilliGray;
    auto air = Material::byNistName("Air, Dry (near sea level)");
    if (b.type == DXB_BEAM_PENCIL) {
        const double e = std::clamp(b.energy, kEMin, kEMax);
        const double perHistory = e * air->massEnergyTransfer(e); // keV/g for one photon per cm2
        return b.air_kerma > 0 && perHistory > 0 ? b.air_kerma / (perHistory * nTotal) : kKeVperGramToMilliGray;
    }
    if (b.type == DXB_BEAM_DX || b.type == DXB_BEAM_CBCT) {
        const dxb_spectrum& s = b.spectrum[0];
        double sw = 0, k = 0;
        for (uint32_t i = 0; i < s.n; ++i) {
            const double e = std::clamp(s.energy_kev[i], kEMin, kEMax);
            sw += s.weight[i];
            k += s.weight[i] * s.energy_kev[i] * air->massEnergyTransfer(e);
        }
        if (!(sw > 0) || !(k > 0) || !(b.dap > 0))
            return kKeVperGramToMilliGray;
        return b.dap / (k / sw * nTotal);
    }
    return kKeVperGramToMilliGray; // CT beams: nested CTDI run (context.cu)
}

} // namespace dxb
