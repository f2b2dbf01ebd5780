// TEST INFRASTRUCTURE (oracle/_ref build only): see pybind11.h next to this file.
#pragma once
#include <pybind11/pybind11.h>
