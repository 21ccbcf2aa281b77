// dxmc/beams/ctspiraldualenergybeam.hpp — dxmc::CTSpiralDualEnergyBeam<ENABLETRACKING>
// (R:src/libopendxmc/beamsettingsmodel.cpp:1413-1832). exposure(2i) = tube A, exposure(2i+1) = tube B
// (R:src/libopendxmc/beamactorcontainer.cpp:134-146).
#pragma once
#include "ctsequentialbeam.hpp"
namespace dxmc {
template <bool ENABLETRACKING = false>
class CTSpiralDualEnergyBeam : public detail::CTBeamBase {
public:
    CTSpiralDualEnergyBeam(const std::array<double, 3>& start = { 0, 0, 0 }, const std::array<double, 3>& stop = { 0, 0, 1 },
        const std::map<std::size_t, double>& filtrationMaterials = {})
        : detail::CTBeamBase(DXB_BEAM_CT_SPIRAL_DUAL)
    {
        m_nTubes = 2;
        setStartStopPosition(start, stop);
        m_tube[0].setFiltrationMaterials(filtrationMaterials);
        m_tube[1].setFiltrationMaterials(filtrationMaterials);
        m_d.tube_b_offset_angle = 90 * DEG_TO_RAD();
    }
    std::array<double, 3> startPosition() const { return get3(m_d.start); }
    std::array<double, 3> stopPosition() const { return get3(m_d.stop); }
    void setStartPosition(const std::array<double, 3>& p) { set3(m_d.start, p); }
    void setStopPosition(const std::array<double, 3>& p) { set3(m_d.stop, p); }
    void setStartStopPosition(const std::array<double, 3>& a, const std::array<double, 3>& b)
    {
        setStartPosition(a);
        setStopPosition(b);
    }
    double pitch() const { return m_d.pitch; }
    void setPitch(double p) { m_d.pitch = std::max(std::abs(p), 0.01); }
    double CTDIvol() const { return m_d.ctdi; }
    void setCTDIvol(double v) { m_d.ctdi = std::abs(v); }
    const CTAECFilter& AECFilter() const { return m_aec; }
    void setAECFilter(const CTAECFilter& f) { m_aec = f; }
    double scanFieldOfViewA() const { return m_d.fov; }
    double scanFieldOfViewB() const { return m_d.fov_b; }
    void setScanFieldOfViewA(double f) { m_d.fov = std::max(std::abs(f), 1.0); }
    void setScanFieldOfViewB(double f) { m_d.fov_b = std::max(std::abs(f), 1.0); }
    double tubeBoffsetAngle() const { return m_d.tube_b_offset_angle; }
    void setTubeBoffsetAngle(double a) { m_d.tube_b_offset_angle = a; }
    double tubeBoffsetAngleDeg() const { return m_d.tube_b_offset_angle * RAD_TO_DEG(); }
    void setTubeBoffsetAngleDeg(double a) { m_d.tube_b_offset_angle = a * DEG_TO_RAD(); }
    const Tube& tubeA() const { return m_tube[0]; }
    const Tube& tubeB() const { return m_tube[1]; }
    void setTubeA(const Tube& t) { m_tube[0] = t; }
    void setTubeB(const Tube& t) { m_tube[1] = t; }
    void setTubeAVoltage(double kv) { m_tube[0].setVoltage(kv); }
    void setTubeBVoltage(double kv) { m_tube[1].setVoltage(kv); }
    void setTubesAnodeAngle(double rad)
    {
        m_tube[0].setAnodeAngle(rad);
        m_tube[1].setAnodeAngle(rad);
    }
    void setTubesAnodeAngleDeg(double deg)
    {
        m_tube[0].setAnodeAngleDeg(deg);
        m_tube[1].setAnodeAngleDeg(deg);
    }
    void addTubeAFiltrationMaterial(std::size_t Z, double mm) { m_tube[0].addFiltrationMaterial(Z, mm); }
    void addTubeBFiltrationMaterial(std::size_t Z, double mm) { m_tube[1].addFiltrationMaterial(Z, mm); }
    double tubeAFiltration(std::size_t Z) const { return m_tube[0].filtration(Z); }
    double tubeBFiltration(std::size_t Z) const { return m_tube[1].filtration(Z); }
    void clearTubeAFiltrationMaterials() { m_tube[0].clearFiltrationMaterials(); }
    void clearTubeBFiltrationMaterials() { m_tube[1].clearFiltrationMaterials(); }
    double tubeAAlHalfValueLayer() const { return m_tube[0].mmAlHalfValueLayer(); }
    double tubeBAlHalfValueLayer() const { return m_tube[1].mmAlHalfValueLayer(); }
    double tubeAMeanSpecterEnergy() const { return m_tube[0].meanSpecterEnergy(); }
    double tubeBMeanSpecterEnergy() const { return m_tube[1].meanSpecterEnergy(); }
    double relativeMasTubeA() const { return m_d.relative_mas_a; }
    double relativeMasTubeB() const { return m_d.relative_mas_b; }
    void setRelativeMasTubeA(double v) { m_d.relative_mas_a = std::abs(v); }
    void setRelativeMasTubeB(double v) { m_d.relative_mas_b = std::abs(v); }
    // weights the exposures of the two tubes carry (mean 1), as the GUI displays them
    double tubeRelativeWeightA() const { return exposure(0).weight() / baseWeight(0); }
    double tubeRelativeWeightB() const { return exposure(1).weight() / baseWeight(1); }
    const BowtieFilter& bowtieFilterA() const { return m_bowtie[0]; }
    const BowtieFilter& bowtieFilterB() const { return m_bowtie[1]; }
    void setBowtieFilterA(const BowtieFilter& f) { m_bowtie[0] = f; }
    void setBowtieFilterB(const BowtieFilter& f) { m_bowtie[1] = f; }

private:
    double baseWeight(std::uint64_t i) const
    {
        // exposure weight = tube weight * AEC(z) * organ AEC(angle); divide the last two out
        dxb_beam_desc d = desc();
        d.relative_mas_a = d.relative_mas_b = 1.0;
        dxb_spectrum flat[2] = { d.spectrum[0], d.spectrum[0] };
        d.spectrum[1] = flat[0];
        dxb_exposure e {};
        dxb_beam_exposure(&d, i, &e);
        return e.weight > 0 ? e.weight : 1.0;
    }
};
}
