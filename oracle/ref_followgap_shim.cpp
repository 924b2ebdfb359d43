// oracle/ref_followgap_shim.cpp -- extern "C" handle around the UNMODIFIED reference FollowGap class
// (followgap/followgap.hpp, header-only).  TEST INFRASTRUCTURE ONLY; compiled by oracle/Makefile with
// -I/root/reference/followgap into oracle/_ref/libfollowgap_ref.so.
#include <algorithm>
#include "followgap.hpp"

extern "C" __attribute__((visibility("default")))
float ref_followgap_eval(float* lidar, int size, int ws, float md, float ma, float inc) {
    FollowGap fg(ws, md, ma, inc);
    return fg.eval(lidar, size);
}
