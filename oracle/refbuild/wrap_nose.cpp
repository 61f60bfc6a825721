// The reference's nubonosedetector element, compiled from its own source, plus a C name for its static merge helper
// (kmsnosedetect.cpp:745-790).  TEST INFRASTRUCTURE ONLY.
#include "kmsnosedetect.cpp"
#include "ref_wrap.h"
REF_REGISTER(kms_nose_detect_plugin_init)

REF_API int ref_nose_merge_consecutive(const int *cur, int ncur, const int *prev, int nprev, const int *face, int scale, int *out, int cap)
{
    REF_TO_VEC(cn, cur, ncur);
    REF_TO_VEC(nv, prev, nprev);
    Rect fc(face[0], face[1], face[2], face[3]);
    vector<Rect> *res = __merge_noses_consecutives_frames(&cn, &nv, fc, scale);
    int n = ref_from_vec(*res, out, cap);
    delete res;
    return n;
}
