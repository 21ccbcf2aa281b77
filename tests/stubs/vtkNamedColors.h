// Stand-in for <vtkNamedColors.h>, tests only.
#pragma once
struct vtkColor3d {
    double v[3] = { 0.5, 0.5, 0.5 };
    const double* GetData() const { return v; }
};
class vtkNamedColors {
public:
    vtkColor3d GetColor3d(const char*) const { return vtkColor3d(); }
};
