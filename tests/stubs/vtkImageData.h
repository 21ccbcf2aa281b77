// Stand-in for <vtkImageData.h>, tests only (see QObject in this directory): enough for datacontainer.cpp to compile
// and link; the tests drive DataContainer through its std::vector interface, never through VTK images.
#pragma once
#define VTK_UNSIGNED_CHAR 3
#define VTK_DOUBLE 11
class vtkImageData {
public:
    void GetDimensions(int* d) const { d[0] = dim[0], d[1] = dim[1], d[2] = dim[2]; }
    int GetNumberOfScalarComponents() const { return 1; }
    int GetScalarType() const { return scalarType; }
    void SetOrigin(const double* o) { origin[0] = o[0], origin[1] = o[1], origin[2] = o[2]; }
    int dim[3] = { 0, 0, 0 };
    int scalarType = VTK_DOUBLE;
    double origin[3] = { 0, 0, 0 };
    void* data = nullptr;
};
