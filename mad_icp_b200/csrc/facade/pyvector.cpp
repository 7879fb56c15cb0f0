// pyvector -- reference: mad_icp/src/pybind/pyvector.cpp (VectorEigen3d)
#include "py_common.hpp"
PYBIND11_MODULE(pyvector, m) { bind_vector_eigen3d(m); }
