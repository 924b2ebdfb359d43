// oracle/ref_shim.cpp -- extern "C" handle around the UNMODIFIED reference Car class.
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with
// /root/reference/racecar/src/racecar.cpp (read where it lies; never copied into the
// repo) into oracle/_ref/libracecar_ref.so.  The class and its methods are the
// reference's (racecar/include/racecar.hpp:26-118); this file only forwards calls.
#include "include/racecar.hpp"

extern "C" {
__attribute__((visibility("default")))
void* ref_car_create(const double* p /* 17 ctor args, reference order */) {
    return new Car(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11],
                   p[12], p[13], p[14], p[15], p[16]);
}
__attribute__((visibility("default"))) void ref_car_destroy(void* c) { delete static_cast<Car*>(c); }
__attribute__((visibility("default"))) void ref_car_control(void* c, double speed, double steer) {
    static_cast<Car*>(c)->control(speed, steer);
}
__attribute__((visibility("default"))) void ref_car_update(void* c, double dt) {
    static_cast<Car*>(c)->updatePosition(dt);
}
__attribute__((visibility("default"))) void ref_car_get_state(void* c, double* s) { static_cast<Car*>(c)->getState(s); }
__attribute__((visibility("default"))) void ref_car_set_state(void* c, double* s) { static_cast<Car*>(c)->setState(s); }
__attribute__((visibility("default"))) void ref_car_scan_pose(void* c, double d, double* pose) {
    static_cast<Car*>(c)->getScanPose(d, pose);
}
__attribute__((visibility("default"))) void ref_car_set_edges(void* c, int n, double amin, double inc, double d) {
    static_cast<Car*>(c)->setCarEdgeDistances(n, amin, inc, d);
}
__attribute__((visibility("default"))) int ref_car_is_crashed(void* c, float* rays, int num_rays, int poses) {
    return static_cast<Car*>(c)->isCrashed(rays, num_rays, poses);
}
}
