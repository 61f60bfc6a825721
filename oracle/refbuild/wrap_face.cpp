// The reference's nubofacedetector element, compiled from its own source (test infrastructure, see ref_wrap.h).
#include "kmsfacedetect.cpp"
#include "ref_wrap.h"
REF_REGISTER(kms_face_detect_plugin_init)
