// The reference's nubotracker element, compiled from its own source, plus C names for its static rectangle helpers
// (gstnubotracker.cpp:119-200).  TEST INFRASTRUCTURE ONLY.
#include "gstnubotracker.cpp"
#include "ref_wrap.h"
REF_REGISTER(gst_nubo_tracker_plugin_init)

REF_API float ref_trk_calc_dist(const int *a, const int *b) { return calc_dist(Point(a[0], a[1]), a[2], a[3], Point(b[0], b[1]), b[2], b[3]); }
REF_API void ref_trk_merge(const int *a, const int *b, int *out)
{
    Rect r = __merge(Rect(a[0], a[1], a[2], a[3]), Rect(b[0], b[1], b[2], b[3]));
    out[0] = r.x; out[1] = r.y; out[2] = r.width; out[3] = r.height;
}
// `tracker`: a nubotracker element from the harness (mh_element_new) whose set_min_area / set_max_area / set_distance
// properties hold the parameters
REF_API int ref_trk_join_objects(void *tracker, int *rects, int n, int cap)
{
    REF_TO_VEC(v, rects, n);
    __join_objects((GstNuboTracker *)tracker, v);
    return ref_from_vec(v, rects, cap);
}
