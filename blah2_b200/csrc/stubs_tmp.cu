// TEMPORARY stubs (replaced by wh.cu / det.cu)
#include "common.cuh"
using namespace b2;
struct b200dd_wh { int x; };
struct b200dd_det { int x; };
extern "C" {
int b200dd_wh_create(int32_t, int32_t, uint32_t, int32_t, b200dd_wh **) { return arg_fail("not implemented"); }
void b200dd_wh_destroy(b200dd_wh *) {}
int b200dd_wh_process_host(b200dd_wh *, const double *, double *) { return arg_fail("not implemented"); }
int b200dd_wh_process_device(b200dd_wh *, const void *, const void *, void *, void *) { return arg_fail("not implemented"); }
int b200dd_wh_last_status(b200dd_wh *) { return arg_fail("not implemented"); }
int b200dd_wh_debug_weights(b200dd_wh *, double *, double *, double *) { return arg_fail("not implemented"); }
uint32_t b200dd_wh_n_bins(const b200dd_wh *) { return 0; }
void *b200dd_wh_stream(b200dd_wh *) { return nullptr; }
int b200dd_det_create(const b200dd_det_params *, uint32_t, uint32_t, b200dd_det **) { return arg_fail("not implemented"); }
void b200dd_det_destroy(b200dd_det *) {}
int b200dd_det_set_metrics_device(b200dd_det *, const void *, uint32_t, uint32_t, double *, void *) { return arg_fail("not implemented"); }
int b200dd_det_process_device(b200dd_det *, int, const void *, uint32_t, uint32_t, const int32_t *, const double *, double, double *, double *, double *, uint32_t, uint32_t *, void *) { return arg_fail("not implemented"); }
int b200dd_det_process_host(b200dd_det *, int, const double *, uint32_t, uint32_t, const int32_t *, const double *, double, double *, double *, double *, uint32_t, uint32_t *) { return arg_fail("not implemented"); }
int b200dd_det_centroid_host(b200dd_det *, const double *, const double *, const double *, uint32_t, double *, double *, double *, uint32_t, uint32_t *) { return arg_fail("not implemented"); }
int b200dd_det_interpolate_host(b200dd_det *, const double *, const double *, const double *, uint32_t, const double *, uint32_t, uint32_t, const int32_t *, const double *, double, double *, double *, double *, uint32_t, uint32_t *) { return arg_fail("not implemented"); }
}
