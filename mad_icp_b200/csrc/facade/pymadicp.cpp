// pymadicp -- reference: mad_icp/src/pybind/tools/pymadicp.cpp:36-52 (MADicp: setQueryCloud /
// setReferenceCloud / compute), same names and defaults; the ICP loop runs in one persistent kernel.
#include "py_common.hpp"
PYBIND11_MODULE(pymadicp, m) {
  bind_vector_eigen3d(m);
  py::class_<mb::MADicpWrapper>(m, "MADicp")
      .def(py::init<int>(), py::arg("num_threads"))
      .def("setQueryCloud",
           [](mb::MADicpWrapper& w, const py::object& q, double b_max, double b_min) {
             w.setQueryCloud(cloud_arg(q), b_max, b_min);
           },
           py::arg("query"), py::arg("b_max") = 0.2, py::arg("b_min") = 0.1)
      .def("setReferenceCloud",
           [](mb::MADicpWrapper& w, const py::object& r, double b_max, double b_min) {
             w.setReferenceCloud(cloud_arg(r), b_max, b_min);
           },
           py::arg("reference"), py::arg("b_max") = 0.2, py::arg("b_min") = 0.1)
      .def("compute",
           [](mb::MADicpWrapper& w, const NpArr& T, size_t icp_iterations, double rho_ker, double b_ratio,
              bool print_stats) {
             return pose_to_numpy(w.compute(pose_from_numpy(T), icp_iterations, rho_ker, b_ratio, print_stats));
           },
           py::arg("T"), py::arg("icp_iterations") = 15, py::arg("rho_ker") = 0.1, py::arg("b_ratio") = 0.02,
           py::arg("print_stats") = false);
  py::register_exception<mb::Error>(m, "MadIcpError", PyExc_RuntimeError);
}
