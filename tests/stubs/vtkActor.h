// Stand-in for <vtkActor.h>, tests only.
#pragma once
#include "vtkPolyDataMapper.h"
#include "vtkProperty.h"
class vtkActor {
public:
    void SetMapper(vtkPolyDataMapper* m) { mapperInput = m ? m->input : nullptr; }
    vtkProperty* GetProperty() { return &property; }
    void SetDragable(bool d) { dragable = d; }
    vtkPolyData* mapperInput = nullptr;
    vtkProperty property;
    bool dragable = false;
};
