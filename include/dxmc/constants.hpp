// dxmc/constants.hpp — angle constants used by OpenDXMC (R:src/libopendxmc/dxmc_specialization.cpp:64-75).
#pragma once
#include <numbers>
namespace dxmc {
template <typename T = double>
consteval T DEG_TO_RAD() { return std::numbers::pi_v<T> / T { 180 }; }
template <typename T = double>
consteval T RAD_TO_DEG() { return T { 180 } / std::numbers::pi_v<T>; }
template <typename T = double>
consteval T PI_VAL() { return std::numbers::pi_v<T>; }
constexpr double MIN_ENERGY() { return 1.0; }
constexpr double MAX_ENERGY() { return 150.0; }
}
