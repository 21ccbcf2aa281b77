// dxmc/world/worlditems/aavoxelgrid.hpp — AAVoxelGrid<NSHELLS, LOWENERGYCORRECTION, TRANSPARENTVOXELS>:
// setData(dim, density, materialIdx, materials), setSpacing, size, doseScored(i)
// (R:src/libopendxmc/simulationpipeline.cpp:127,145-150,174-219).  The grid keeps POINTERS to the caller's arrays until
// World::build() uploads them (the reference passes const& to DataContainer vectors that outlive the call, :147-149).
#pragma once
#include "../../../dxb.h"
#include "../../material/material.hpp"
#include "../dosescore.hpp"
#include <array>
#include <cstdint>
#include <vector>
namespace dxmc {
template <int NSHELLS = 5, int LOWENERGYCORRECTION = 1, std::uint8_t TRANSPARENTVOXELS = 255>
class AAVoxelGrid {
public:
    static constexpr int lowEnergyCorrection() { return LOWENERGYCORRECTION; }
    bool setData(const std::array<std::size_t, 3>& dim, const std::vector<double>& density, const std::vector<std::uint8_t>& materialIdx,
        const std::vector<Material<NSHELLS>>& materials)
    {
        const std::size_t n = dim[0] * dim[1] * dim[2];
        if (n == 0 || density.size() != n || materialIdx.size() != n || materials.empty() || materials.size() > 255)
            return false;
        m_dim = dim;
        m_density = density.data();
        m_material = materialIdx.data();
        m_materials = materials;
        m_doseValid = false;
        return true;
    }
    void setSpacing(const std::array<double, 3>& s) { m_spacing = s; }
    const std::array<double, 3>& spacing() const { return m_spacing; }
    const std::array<std::size_t, 3>& dimensions() const { return m_dim; }
    std::size_t size() const { return m_dim[0] * m_dim[1] * m_dim[2]; }
    void translate(const std::array<double, 3>& d)
    {
        for (int i = 0; i < 3; ++i)
            m_center[i] += d[i];
    }
    const std::array<double, 3>& center() const { return m_center; }

    DoseScore doseScored(std::size_t i) const
    {
        fetch();
        return DoseScore(m_dose[i], m_variance[i], m_events[i]);
    }
    // one D2H copy of all three arrays (what doseScored(i) indexes)
    const std::vector<double>& doseArray() const
    {
        fetch();
        return m_dose;
    }
    const std::vector<double>& varianceArray() const
    {
        fetch();
        return m_variance;
    }
    const std::vector<std::uint64_t>& eventArray() const
    {
        fetch();
        return m_events;
    }

    // ---- used by World / Transport
    int upload(dxb_ctx* ctx)
    {
        m_ctx = ctx;
        std::vector<const dxb_material*> h;
        for (const auto& m : m_materials)
            h.push_back(m.handle());
        int rc = dxb_set_materials(ctx, static_cast<uint32_t>(h.size()), h.data());
        if (rc != DXB_OK)
            return rc;
        dxb_set_grid_center(ctx, m_center.data());
        const uint64_t dim[3] = { m_dim[0], m_dim[1], m_dim[2] };
        rc = dxb_set_grid(ctx, dim, m_spacing.data(), m_density, m_material);
        m_doseValid = false;
        return rc;
    }
    void invalidateDose() { m_doseValid = false; }

private:
    void fetch() const
    {
        if (m_doseValid || !m_ctx)
            return;
        const std::size_t n = size();
        m_dose.assign(n, 0.0);
        m_variance.assign(n, 0.0);
        m_events.assign(n, 0);
        dxb_get_dose(m_ctx, m_dose.data(), m_variance.data(), m_events.data());
        m_doseValid = true;
    }
    std::array<std::size_t, 3> m_dim { 0, 0, 0 };
    std::array<double, 3> m_spacing { 1, 1, 1 }, m_center { 0, 0, 0 };
    const double* m_density = nullptr;
    const std::uint8_t* m_material = nullptr;
    std::vector<Material<NSHELLS>> m_materials;
    dxb_ctx* m_ctx = nullptr;
    mutable bool m_doseValid = false;
    mutable std::vector<double> m_dose, m_variance;
    mutable std::vector<std::uint64_t> m_events;
};
}
