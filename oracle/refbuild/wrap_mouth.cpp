// The reference's nubomouthdetector element, compiled from its own source, plus a C name for its static merge helper
// (kmsmouthdetect.cpp:750-796).  TEST INFRASTRUCTURE ONLY.
#include "kmsmouthdetect.cpp"
#include "ref_wrap.h"
REF_REGISTER(kms_mouth_detect_plugin_init)

REF_API int ref_mouth_merge_consecutive(const int *cur, int ncur, const int *prev, int nprev, const int *face, int scale, int *out, int cap)
{
    REF_TO_VEC(cm, cur, ncur);
    REF_TO_VEC(mv, prev, nprev);
    Rect fc(face[0], face[1], face[2], face[3]);
    vector<Rect> *res = __merge_mouths_consecutives_frames(&cm, &mv, fc, scale);
    int n = ref_from_vec(*res, out, cap);
    delete res;
    return n;
}
