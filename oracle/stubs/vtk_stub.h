// oracle/stubs/vtk*.h -- TEST INFRASTRUCTURE ONLY: just enough of the VTK reader API for
// SylinderSystem::setInitialFromVTKFile (SimToolbox/Sylinder/SylinderSystem.cpp:407-470) to COMPILE; VTK is absent
// from this image and reading a .pvtp restart through the oracle build aborts.
#pragma once
#include <cstdio>
#include <cstdlib>
struct vtkObjectStub {
    [[noreturn]] static void unavailable() {
        fprintf(stderr, "oracle build: VTK is not available (restart from .pvtp is not supported)\n");
        abort();
    }
};
template <class T>
class vtkSmartPointer {
  public:
    vtkSmartPointer() : p_(nullptr) {}
    vtkSmartPointer(T *p) : p_(p) {}
    template <class U>
    vtkSmartPointer(const vtkSmartPointer<U> &o) : p_(o.get()) {}
    static vtkSmartPointer New() { return vtkSmartPointer(new T()); }
    T *operator->() const { return p_; }
    T *get() const { return p_; }
    operator T *() const { return p_; }

  private:
    T *p_;
};
class vtkAbstractArray : public vtkObjectStub {
  public:
    virtual ~vtkAbstractArray() {}
};
class vtkDataArray : public vtkAbstractArray {
  public:
    double GetComponent(long long, int) { unavailable(); }
};
class vtkTypeInt32Array : public vtkDataArray {
  public:
    int GetTypedComponent(long long, int) { unavailable(); }
};
class vtkTypeUInt8Array : public vtkDataArray {
  public:
    unsigned char GetTypedComponent(long long, int) { unavailable(); }
};
template <class T>
T *vtkArrayDownCast(vtkAbstractArray *a) { return dynamic_cast<T *>(a); }
class vtkPoints : public vtkObjectStub {
  public:
    long long GetNumberOfPoints() { unavailable(); }
    void GetPoint(long long, double *) { unavailable(); }
};
class vtkCellData : public vtkObjectStub {
  public:
    vtkAbstractArray *GetAbstractArray(const char *) { unavailable(); }
    vtkDataArray *GetArray(const char *) { unavailable(); }
};
class vtkPolyData : public vtkObjectStub {
  public:
    vtkPoints *GetPoints() { unavailable(); }
    vtkCellData *GetCellData() { unavailable(); }
};
class vtkXMLPPolyDataReader : public vtkObjectStub {
  public:
    void SetFileName(const char *) {}
    void Update() { unavailable(); }
    vtkPolyData *GetOutput() { unavailable(); }
};
typedef vtkXMLPPolyDataReader vtkXMLPolyDataReader;
