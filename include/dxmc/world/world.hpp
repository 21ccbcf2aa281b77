// dxmc/world/world.hpp — dxmc::World<Item>: addItem<T>() -> T&, build() (R:src/libopendxmc/simulationpipeline.cpp:128-131,153).
// OpenDXMC's world holds exactly one AAVoxelGrid; the shim owns the dxb_ctx (one or more B200s).  GPUs are chosen with
// the environment variable DXMC_B200_DEVICES ("0,1,2,3"; default: the current device) so that the single-process GUI
// can use a whole box without source changes.
#pragma once
#include "../../dxb.h"
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
namespace dxmc {
template <typename Item>
class World {
public:
    World() = default;
    ~World()
    {
        if (m_ctx)
            dxb_destroy(m_ctx);
    }
    World(const World&) = delete;
    World& operator=(const World&) = delete;
    template <typename T>
    T& addItem(T item = T {})
    {
        static_assert(std::is_same_v<T, Item>, "this shim's World holds one item type");
        m_item = std::make_unique<Item>(std::move(item));
        return *m_item;
    }
    Item& item() { return *m_item; }
    const Item& item() const { return *m_item; }
    void build(double /*AABB padding, unused*/ = 0)
    {
        if (!m_item)
            throw std::runtime_error("dxmc::World::build: no item");
        if (!m_ctx) {
            std::vector<int> devs;
            if (const char* env = std::getenv("DXMC_B200_DEVICES")) {
                std::string s(env), tok;
                for (char ch : s + ",") {
                    if (ch == ',') {
                        if (!tok.empty())
                            devs.push_back(std::atoi(tok.c_str()));
                        tok.clear();
                    } else {
                        tok.push_back(ch);
                    }
                }
            }
            const int rc = dxb_create(&m_ctx, devs.empty() ? nullptr : devs.data(), static_cast<int>(devs.size()));
            if (rc != DXB_OK)
                throw std::runtime_error("dxmc::World::build: dxb_create failed (no CUDA device? there is no CPU fallback)");
            // DXMC_B200_OPTIONS="key=value,key=value": dxb_set_option for an application that cannot be recompiled, e.g.
            // "dense_box=0,local_majorant=0" = plain Woodcock tracking with one majorant (the reference's rule)
            if (const char* opts = std::getenv("DXMC_B200_OPTIONS")) {
                std::string tok;
                for (char ch : std::string(opts) + ",") {
                    if (ch != ',') {
                        tok.push_back(ch);
                        continue;
                    }
                    const auto eq = tok.find('=');
                    if (eq != std::string::npos && eq > 0) {
                        const std::string key = tok.substr(0, eq);
                        if (dxb_set_option(m_ctx, key.c_str(), std::atof(tok.c_str() + eq + 1)) != DXB_OK)
                            throw std::runtime_error("dxmc::World::build: DXMC_B200_OPTIONS: " + std::string(dxb_last_error(m_ctx)));
                    } else if (!tok.empty()) {
                        throw std::runtime_error("dxmc::World::build: DXMC_B200_OPTIONS: expected key=value, got '" + tok + "'");
                    }
                    tok.clear();
                }
            }
        }
        const int rc = m_item->upload(m_ctx);
        if (rc != DXB_OK)
            throw std::runtime_error(std::string("dxmc::World::build: ") + dxb_last_error(m_ctx));
    }
    void clearDoseScored()
    {
        if (m_ctx)
            dxb_clear_dose(m_ctx);
        if (m_item)
            m_item->invalidateDose();
    }
    dxb_ctx* ctx() const { return m_ctx; }

private:
    std::unique_ptr<Item> m_item;
    dxb_ctx* m_ctx = nullptr;
};
}
