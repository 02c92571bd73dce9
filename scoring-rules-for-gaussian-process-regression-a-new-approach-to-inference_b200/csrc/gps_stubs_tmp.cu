#include "gps_common.cuh"
extern "C" {
int gps_fitc_eval(gps_ctx* ctx, const double*, const double*, int, double, int, double*, double*, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_fitc_acc_len(int, int, int64_t*, int64_t*, int64_t*) { return GPS_ESTATE; }
int gps_fitc_begin(gps_ctx* ctx, const double*, const double*, int, double, int, int64_t) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_fitc_pass1(gps_ctx* ctx, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_fitc_pass2(gps_ctx* ctx, const double*, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_fitc_pass3(gps_ctx* ctx, const double*, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_fitc_finish(gps_ctx* ctx, const double*, const double*, double*, double*, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_fitc_loo(gps_ctx* ctx, double*, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_fitc_predict(gps_ctx* ctx, const double*, int64_t, double*, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
int gps_grid_eval(gps_ctx* ctx, const double*, const double*, int, const double*, const double*, int64_t, int, double*) { return gps_fail(ctx, GPS_ESTATE, "not implemented"); }
}
