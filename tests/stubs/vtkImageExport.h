// Stand-in for <vtkImageExport.h>, tests only.
#pragma once
#include "vtkImageData.h"
class vtkImageExport {
public:
    void ReleaseDataFlagOn() { }
    void SetInputData(vtkImageData* d) { in = d; }
    void* GetPointerToData() { return in ? in->data : nullptr; }

private:
    vtkImageData* in = nullptr;
};
