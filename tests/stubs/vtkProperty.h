// Stand-in for <vtkProperty.h>, tests only.
#pragma once
class vtkProperty {
public:
    void SetColor(const double* c) { color[0] = c[0], color[1] = c[1], color[2] = c[2]; }
    void SetOpacity(double o) { opacity = o; }
    void SetLineWidth(double) { }
    void RenderLinesAsTubesOn() { }
    double color[3] = { 0, 0, 0 };
    double opacity = 0;
};
