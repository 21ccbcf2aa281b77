// Functional stand-in for <vtkPolyData.h>, tests only.
#pragma once
#include "vtkCellArray.h"
#include "vtkPoints.h"
#include "vtkSmartPointer.h"
class vtkPolyData {
public:
    void SetPoints(vtkSmartPointer<vtkPoints> p) { points = p; }
    void SetLines(vtkSmartPointer<vtkCellArray> c) { lines = c; }
    vtkSmartPointer<vtkPoints> points;
    vtkSmartPointer<vtkCellArray> lines;
};
