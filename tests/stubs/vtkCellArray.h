// Functional stand-in for <vtkCellArray.h>, tests only.
#pragma once
#include <vector>
class vtkCellArray {
public:
    long long InsertNextCell(long long n)
    {
        cells.emplace_back();
        cells.back().reserve(static_cast<std::size_t>(n));
        return static_cast<long long>(cells.size()) - 1;
    }
    void InsertCellPoint(long long id) { cells.back().push_back(id); }
    std::vector<std::vector<long long>> cells;
};
