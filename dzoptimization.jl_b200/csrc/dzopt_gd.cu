// dzopt_gd.cu -- placeholder translation unit (filled in by the GradientDescent / Riesz work)
#include "host_common.h"
using namespace dzo;
namespace dzo {
int riesz_dev_objective(int, int64_t, int, int64_t, int64_t, const double*, double*) { return fail(DZO_ERR_UNSUPPORTED, "riesz: not built yet"); }
int riesz_dev_gradient(int, int64_t, int, int64_t, int64_t, const double*, double*) { return fail(DZO_ERR_UNSUPPORTED, "riesz: not built yet"); }
int riesz_dev_line_search(int, int64_t, int, int64_t, const double*, const double*, double, double, double*, double*) { return fail(DZO_ERR_UNSUPPORTED, "riesz: not built yet"); }
}
struct dzo_gd { int dummy; };
#define STUB return fail(DZO_ERR_UNSUPPORTED, "dzo_gd_*: not built yet")
extern "C" {
int dzo_gd_create(dzo_gd**, int, int, int64_t, int64_t, int64_t, const double*, double, int, int) { STUB; }
int dzo_gd_set_stream(dzo_gd*, void*) { STUB; }
int dzo_gd_step(dzo_gd*, int) { STUB; }
int dzo_gd_step_async(dzo_gd*, int) { STUB; }
int dzo_gd_sync(dzo_gd*) { STUB; }
int dzo_gd_get_point(dzo_gd*, double*) { STUB; }
int dzo_gd_get_delta_point(dzo_gd*, double*) { STUB; }
int dzo_gd_get_gradient(dzo_gd*, double*) { STUB; }
int dzo_gd_get_delta_gradient(dzo_gd*, double*) { STUB; }
int dzo_gd_get_direction(dzo_gd*, double*) { STUB; }
int dzo_gd_get_objective(dzo_gd*, double*) { STUB; }
int dzo_gd_get_delta_objective(dzo_gd*, double*) { STUB; }
int dzo_gd_get_step_length(dzo_gd*, double*) { STUB; }
int dzo_gd_get_iteration_count(dzo_gd*, int64_t*) { STUB; }
int dzo_gd_get_terminated(dzo_gd*, uint8_t*) { STUB; }
int dzo_gd_info(dzo_gd*, int64_t*, int64_t*, int*) { STUB; }
void dzo_gd_destroy(dzo_gd*) {}
}
