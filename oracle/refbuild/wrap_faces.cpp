// Faces / BaseFace (Faces.cpp, BaseFace.cpp are compiled beside this file from /root/reference, unmodified): a C driver
// around Faces::track_faces (Faces.cpp:78-153) for the glue fuzz tests.  TEST INFRASTRUCTURE ONLY.
#include "Faces.hpp"
#include "ref_wrap.h"

REF_API void *ref_faces_new(void) { return new Faces(); }
REF_API void ref_faces_free(void *f) { delete (Faces *)f; }
REF_API void ref_faces_clear(void *f) { ((Faces *)f)->clear(); }
// kmsfacedetect.cpp:813-816: Faces cf(*current_faces); faces->track_faces(&cf, track, euclidean, area, num_iter)
REF_API void ref_faces_track(void *f, const int *cur, int ncur, int track_threshold, int pos_threshold, int area_threshold, int n_iter)
{
    REF_TO_VEC(v, cur, ncur);
    Faces cf(v);
    ((Faces *)f)->track_faces(&cf, track_threshold, pos_threshold, area_threshold, n_iter);
}
REF_API int ref_faces_get(void *f, int *rects, int *ids, int cap)
{
    vector<BaseFace> dummy, *bf = &dummy;
    ((Faces *)f)->get_faces(&bf);
    int n = 0;
    for (BaseFace &b : *bf) {
        if (n < cap) { Rect r = b.get_face(); rects[4 * n] = r.x; rects[4 * n + 1] = r.y; rects[4 * n + 2] = r.width; rects[4 * n + 3] = r.height; ids[n] = b.get_id(); }
        n++;
    }
    return n;
}
