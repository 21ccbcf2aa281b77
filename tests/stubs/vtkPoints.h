// Functional stand-in for <vtkPoints.h>, tests only: keeps the points so that a test can read the geometry back.
#pragma once
#include <array>
#include <vector>
class vtkPoints {
public:
    long long InsertNextPoint(const double* p)
    {
        pts.push_back({ p[0], p[1], p[2] });
        return static_cast<long long>(pts.size()) - 1;
    }
    std::vector<std::array<double, 3>> pts;
};
