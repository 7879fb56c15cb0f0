"""pybind11 modules with the reference's names (mad_icp.src.pybind.{pyvector,pymadtree,pymadicp}):

    from mad_icp_b200.pybind.pyvector import VectorEigen3d
    from mad_icp_b200.pybind.pymadicp import MADicp
    from mad_icp_b200.pybind.pymadtree import MADtree

Built in-tree by `__graft_entry__.build()` (make -C mad_icp_b200/csrc/facade)."""
