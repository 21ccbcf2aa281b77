// Stand-in for <vtkTubeFilter.h>, tests only: remembers its input so that the test can reach the poly data.
#pragma once
#include "vtkPolyData.h"
class vtkTubeFilter;
struct vtkAlgorithmOutput {
    vtkTubeFilter* filter = nullptr;
};
class vtkTubeFilter {
public:
    void SetInputData(vtkPolyData* d) { input = d; }
    void SetRadius(double r) { radius = r; }
    void SetNumberOfSides(int n) { sides = n; }
    vtkAlgorithmOutput* GetOutputPort()
    {
        port.filter = this;
        return &port;
    }
    vtkPolyData* input = nullptr;
    double radius = 0;
    int sides = 0;
    vtkAlgorithmOutput port;
};
