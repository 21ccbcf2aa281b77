// Minimal stand-in for <vtkSmartPointer.h>, tests only (see QObject in this directory): shared ownership, New().
#pragma once
#include <cstddef>
#include <memory>
template <typename T>
class vtkSmartPointer {
public:
    vtkSmartPointer() = default;
    vtkSmartPointer(std::nullptr_t) { }
    static vtkSmartPointer New()
    {
        vtkSmartPointer r;
        r.p = std::make_shared<T>();
        return r;
    }
    T* operator->() const { return p.get(); }
    T& operator*() const { return *p; }
    T* Get() const { return p.get(); }
    T* GetPointer() const { return p.get(); }
    operator T*() const { return p.get(); }

private:
    std::shared_ptr<T> p;
};
