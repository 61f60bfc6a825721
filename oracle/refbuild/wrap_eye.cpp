// The reference's nuboeyedetector element, compiled from its own source, plus C names for its static helpers
// (kmseyedetect.cpp:766-913).  TEST INFRASTRUCTURE ONLY.
#include "kmseyedetect.cpp"
#include "ref_wrap.h"
REF_REGISTER(kms_eye_detect_plugin_init)

REF_API int ref_eye_contain_bb(int px, int py, const int *r) { return __contain_bb(Point(px, py), Rect(r[0], r[1], r[2], r[3])) ? 1 : 0; }
// eye_r_same != 0: the call shape of the right eye (kmseyedetect.cpp:1016), where eye_r IS the list being merged
REF_API int ref_eye_merge_current_frame(const int *face_bb, const int *eye_r, int n_eye_r, int eye_r_same, int *eyes, int n_eyes, int scale,
                                        int eye_left, int cap)
{
    REF_TO_VEC(er, eye_r, n_eye_r);
    REF_TO_VEC(ev, eyes, n_eyes);
    __merge_eyes_current_frame(Rect(face_bb[0], face_bb[1], face_bb[2], face_bb[3]), eye_r_same ? &ev : &er, ev, scale, eye_left != 0);
    return ref_from_vec(ev, eyes, cap);
}
REF_API int ref_eye_merge_consecutive(const int *cur, int ncur, const int *prev, int nprev, const int *face, int scale, int eye_left, int *out, int cap)
{
    REF_TO_VEC(ce, cur, ncur);
    REF_TO_VEC(ev, prev, nprev);
    Rect fc(face[0], face[1], face[2], face[3]);
    vector<Rect> *res = __merge_eyes_consecutives_frames(&ce, &ev, fc, scale, eye_left != 0);
    int n = ref_from_vec(*res, out, cap);
    delete res;
    return n;
}
REF_API int ref_eye_to_global(int *eyes, int n, const int *face, int scale)
{
    REF_TO_VEC(ev, eyes, n);
    transform_2_global_coordinates(&ev, Rect(face[0], face[1], face[2], face[3]), scale);
    return ref_from_vec(ev, eyes, n);
}
