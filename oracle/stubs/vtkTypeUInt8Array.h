#include "vtk_stub.h"
