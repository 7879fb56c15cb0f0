// py_common.hpp -- shared pieces of the pybind11 modules (pyvector, pymadtree, pymadicp): the opaque
// VectorEigen3d (reference: pybind/eigen_stl_bindings.h:44-97, pyvector.cpp) and numpy <-> pose helpers.
#pragma once
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "facade.hpp"

PYBIND11_MAKE_OPAQUE(madicp_b200::ContainerType);

namespace py = pybind11;
namespace mb = madicp_b200;

using NpArr = py::array_t<double, py::array::c_style | py::array::forcecast>;

// numpy N x 3 float64 -> vector (wrong shape raises cast_error, like the reference: eigen_stl_bindings.h:48-50)
inline mb::ContainerType vector_from_numpy(const NpArr& a) {
  if (a.ndim() != 2 || a.shape(1) != 3) throw py::cast_error();
  mb::ContainerType v(size_t(a.shape(0)));
  if (!v.empty()) std::memcpy(v[0].data(), a.data(), sizeof(double) * 3 * v.size());
  return v;
}
inline py::array_t<double> numpy3(const mb::Vector3d& p) {
  py::array_t<double> out(3);
  std::memcpy(out.mutable_data(), p.data(), 24);
  return out;
}
inline mb::Matrix4d pose_from_numpy(const NpArr& T) {
  if (T.ndim() != 2 || T.shape(0) != 4 || T.shape(1) != 4) throw py::cast_error();
  mb::Matrix4d M;
  std::memcpy(M.m, T.data(), sizeof(M.m));
  return M;
}
inline py::array_t<double> pose_to_numpy(const mb::Matrix4d& M) {
  py::array_t<double> out({4, 4});
  std::memcpy(out.mutable_data(), M.m, sizeof(M.m));
  return out;
}

// VectorEigen3d: registered once per interpreter (whichever module is imported first).
inline void bind_vector_eigen3d(py::module_& m) {
  if (auto* ti = py::detail::get_type_info(typeid(mb::ContainerType))) {  // another module got there first
    m.attr("VectorEigen3d") = py::reinterpret_borrow<py::object>(reinterpret_cast<PyObject*>(ti->type));
    return;
  }
  py::class_<mb::ContainerType>(m, "VectorEigen3d", py::buffer_protocol())
      .def(py::init<>())
      .def(py::init([](const NpArr& a) { return vector_from_numpy(a); }))
      .def_buffer([](mb::ContainerType& v) -> py::buffer_info {
        return py::buffer_info(v.empty() ? nullptr : v[0].data(), sizeof(double), py::format_descriptor<double>::format(),
                               2, {v.size(), size_t(3)}, {sizeof(mb::Vector3d), sizeof(double)});
      })
      .def("__len__", [](const mb::ContainerType& v) { return v.size(); })
      .def("__bool__", [](const mb::ContainerType& v) { return !v.empty(); })
      .def("__getitem__",
           [](const mb::ContainerType& v, long i) {
             if (i < 0) i += long(v.size());
             if (i < 0 || size_t(i) >= v.size()) throw py::index_error();
             return numpy3(v[size_t(i)]);
           })
      .def("__iter__",
           [](const mb::ContainerType& v) {
             py::list l;
             for (const auto& p : v) l.append(numpy3(p));
             return py::iter(l);
           })
      .def("append", [](mb::ContainerType& v, const NpArr& p) {
        if (p.size() != 3) throw py::cast_error();
        v.push_back({p.data()[0], p.data()[1], p.data()[2]});
      })
      .def("__copy__", [](const mb::ContainerType& v) { return mb::ContainerType(v); })
      .def("__deepcopy__", [](const mb::ContainerType& v, py::dict) { return mb::ContainerType(v); })
      .def("__repr__", [](const mb::ContainerType& v) {
        return std::string("std::vector<Eigen::Vector3d> with ") + std::to_string(v.size()) +
               " elements.\nUse numpy.asarray() to access data.";
      });
}

// arguments typed `VectorEigen3d` in the reference also accept a plain N x 3 numpy array here
inline mb::ContainerType cloud_arg(const py::object& o) {
  if (py::isinstance<mb::ContainerType>(o)) return o.cast<mb::ContainerType>();
  return vector_from_numpy(o.cast<NpArr>());
}
