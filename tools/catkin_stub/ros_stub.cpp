// ros_stub.cpp — TEST INFRASTRUCTURE: the single definition of the stand-in ROS pump (oracle/ref_shim) for the
// catkin stand-in build of ros_shell/ (tools/catkin_stub/catkinConfig.cmake).
#include "pump_impl.hpp"
