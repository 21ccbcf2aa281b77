// Stand-in for <vtkPolyDataMapper.h>, tests only.
#pragma once
#include "vtkTubeFilter.h"
class vtkPolyDataMapper {
public:
    void SetInputConnection(vtkAlgorithmOutput* p) { input = p ? p->filter->input : nullptr; }
    void SetInputData(vtkPolyData* d) { input = d; }
    vtkPolyData* input = nullptr; // the filter object itself dies with createActor(); the poly data is owned by the container
};
