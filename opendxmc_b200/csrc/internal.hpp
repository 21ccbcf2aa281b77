// internal.hpp — opaque handle definitions shared by capi.cpp and context.cu
#pragma once
#include <memory>
#include "physics.hpp"

struct dxb_material {
    std::shared_ptr<dxb::Material> m;
};
