// Stand-in for <vtkImageImport.h>, tests only.
#pragma once
#include "vtkImageData.h"
#include "vtkSmartPointer.h"
class vtkImageImport {
public:
    void ReleaseDataFlagOn() { }
    void SetDataScalarTypeToDouble() { out->scalarType = VTK_DOUBLE; }
    void SetDataScalarTypeToUnsignedChar() { out->scalarType = VTK_UNSIGNED_CHAR; }
    void SetNumberOfScalarComponents(int) { }
    void SetWholeExtent(const int* e) { out->dim[0] = e[1] - e[0] + 1, out->dim[1] = e[3] - e[2] + 1, out->dim[2] = e[5] - e[4] + 1; }
    void SetDataExtent(const int*) { }
    void SetDataExtentToWholeExtent() { }
    void SetDataSpacing(const double*) { }
    void SetImportVoidPointer(void* p) { out->data = p; }
    void Update() { }
    vtkSmartPointer<vtkImageData> GetOutput() { return out; }

private:
    vtkSmartPointer<vtkImageData> out = vtkSmartPointer<vtkImageData>::New();
};
