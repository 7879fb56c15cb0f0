// pypeline -- reference: mad_icp/src/pybind/pypeline.cpp:52-75 (Pipeline + VectorEigen3d), same ctor
// arguments and method names; registration runs on the GPU.
#include "pipeline.hpp"
#include "py_common.hpp"
PYBIND11_MODULE(pypeline, m) {
  bind_vector_eigen3d(m);
  py::class_<mb::Pipeline>(m, "Pipeline")
      .def(py::init<double, bool, double, double, double, double, double, int, int, bool>(), py::arg("sensor_hz"),
           py::arg("deskew"), py::arg("b_max"), py::arg("rho_ker"), py::arg("p_th"), py::arg("b_min"), py::arg("b_ratio"),
           py::arg("num_keyframes"), py::arg("num_threads"), py::arg("realtime"))
      .def("currentPose", [](const mb::Pipeline& p) { return pose_to_numpy(p.currentPose()); })
      .def("trajectory",
           [](const mb::Pipeline& p) {
             py::list out;
             for (const auto& T : p.trajectory()) out.append(pose_to_numpy(T));
             return out;
           })
      .def("keyframePose", [](const mb::Pipeline& p) { return pose_to_numpy(p.keyframePose()); })
      .def("isInitialized", &mb::Pipeline::isInitialized)
      .def("isMapUpdated", &mb::Pipeline::isMapUpdated)
      .def("currentID", &mb::Pipeline::currentID)
      .def("keyframeID", &mb::Pipeline::keyframeID)
      .def("modelLeaves", &mb::Pipeline::modelLeaves)
      .def("currentLeaves", &mb::Pipeline::currentLeaves)
      .def("compute", [](mb::Pipeline& p, double stamp, const py::object& cloud) {
        // read the points where they are: a bound VectorEigen3d by reference, a numpy array through its buffer
        if (py::isinstance<mb::ContainerType>(cloud)) {
          const mb::ContainerType& v = cloud.cast<const mb::ContainerType&>();
          p.compute(stamp, v.empty() ? nullptr : v[0].data(), v.size());
        } else if (py::isinstance<py::array>(cloud) && py::array::ensure(cloud).dtype().is(py::dtype::of<float>())) {
          // float32 as the dataset readers deliver it: converted on the device (no host copy in float64)
          const auto a = cloud.cast<py::array_t<float, py::array::c_style | py::array::forcecast>>();
          if (a.ndim() != 2 || a.shape(1) != 3) throw py::cast_error();
          p.computeF32(stamp, a.shape(0) ? a.data() : nullptr, size_t(a.shape(0)));
        } else {
          const NpArr a = cloud.cast<NpArr>();  // a view when the array already is C-contiguous float64
          if (a.ndim() != 2 || a.shape(1) != 3) throw py::cast_error();
          p.compute(stamp, a.shape(0) ? a.data() : nullptr, size_t(a.shape(0)));
        }
      })
      // additions (not in the reference): diagnostics
      .def_static("_deskewOnly", [](const py::object& cloud, const NpArr& a, const NpArr& b, double sensor_hz, int num_threads) {
        return mb::Pipeline::deskewOnly(cloud_arg(cloud), pose_from_numpy(a), pose_from_numpy(b), sensor_hz, num_threads);
      }, py::arg("cloud"), py::arg("T_prev"), py::arg("T_now"), py::arg("sensor_hz"), py::arg("num_threads") = 1)
      .def("prefetch", [](mb::Pipeline& p, const py::object& cloud) {
        // the array is read in place when the batch is built (inside a later compute()): a reference keeps it alive
        // until then (it is dropped there, on the calling thread, with the GIL held)
        auto hold = [](const py::object& o) {
          py::object* ref = new py::object(o);
          return std::shared_ptr<void>(ref, [](void* q) { delete static_cast<py::object*>(q); });
        };
        if (py::isinstance<mb::ContainerType>(cloud)) {
          const mb::ContainerType& v = cloud.cast<const mb::ContainerType&>();
          return p.prefetch(v.empty() ? nullptr : v[0].data(), v.size(), false, hold(cloud));
        }
        if (py::isinstance<py::array>(cloud) && py::array::ensure(cloud).dtype().is(py::dtype::of<float>())) {
          const auto a = cloud.cast<py::array_t<float, py::array::c_style | py::array::forcecast>>();
          if (a.ndim() != 2 || a.shape(1) != 3) throw py::cast_error();
          return p.prefetch(a.data(), size_t(a.shape(0)), true, hold(a));
        }
        const NpArr a = cloud.cast<NpArr>();
        if (a.ndim() != 2 || a.shape(1) != 3) throw py::cast_error();
        return p.prefetch(a.data(), size_t(a.shape(0)), false, hold(a));
      }, py::arg("cloud"))
      .def("prefetched", &mb::Pipeline::prefetched)
      .def("lastIcpIterations", &mb::Pipeline::lastIcpIterations)
      .def("gpuBuild", &mb::Pipeline::gpuBuild)
      .def("inliersRatio", &mb::Pipeline::inliersRatio)
      .def("numKeyframes", &mb::Pipeline::numKeyframes);
  py::register_exception<mb::Error>(m, "MadIcpError", PyExc_RuntimeError);
}
