// shim_check.cpp — compiles pin.cpp in the build container, where PCL does not exist: the stand-in headers of
// oracle/ref_shim take PCL's place and the two classes delegate to the oracle (oracle/ref_shim/pcl_delegates.hpp).
// This checks that pin.cpp compiles, runs and writes well-formed goldens; it pins NOTHING about PCL.
//   g++ -std=c++17 -I oracle -I oracle/ref_shim -I oracle/ref_shim/include tools/pcl_pin/shim_check.cpp \
//       oracle/cones_oracle.cpp -o /tmp/pcl_pin_shim
#include "cones_oracle.h"
#include "shim_core.hpp"
#include "pcl_delegates.hpp"

#include "pin.cpp"
