// TEST INFRASTRUCTURE (oracle/_ref build only): empty stand-in so that the reference's distance
// headers, which include <pybind11/pybind11.h> only for `namespace py = pybind11;`, compile
// without Python.  Nothing from pybind11 is used by the functions we instantiate.
#pragma once
#include <Eigen/Core>
namespace pybind11 {}
