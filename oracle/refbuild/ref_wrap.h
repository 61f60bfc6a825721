// ref_wrap.h — shared by the wrapper translation units of oracle/refbuild: each one #includes ONE reference element
// source as it lies under /root/reference (found through -I, never copied), registers the element with the mock
// registry at load time and exports the file's static helper functions under C names for the glue fuzz tests.
// TEST INFRASTRUCTURE ONLY.
#ifndef REF_WRAP_H
#define REF_WRAP_H
#define REF_API extern "C" __attribute__((visibility("default")))
#define REF_REGISTER(init_fn) namespace { struct RefRegistrar { RefRegistrar() { init_fn(NULL); } } ref_registrar_instance; }

// flat rectangle lists <-> vector<Rect>
#define REF_TO_VEC(vec, ptr, n) std::vector<cv::Rect> vec; for (int i_ = 0; i_ < (n); i_++) vec.push_back(cv::Rect((ptr)[4 * i_], (ptr)[4 * i_ + 1], (ptr)[4 * i_ + 2], (ptr)[4 * i_ + 3]))
static inline int ref_from_vec(const std::vector<cv::Rect> &v, int *out, int cap)
{
    int n = 0;
    for (const cv::Rect &r : v) { if (n < cap) { out[4 * n] = r.x; out[4 * n + 1] = r.y; out[4 * n + 2] = r.width; out[4 * n + 3] = r.height; } n++; }
    return n;
}
#endif
