// stand-in header: see shim_core.hpp (test infrastructure, not ROS/PCL/Eigen code)
#pragma once
#include "shim_core.hpp"
