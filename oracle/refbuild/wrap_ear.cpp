// The reference's nuboeardetector element, compiled from its own source (test infrastructure, see ref_wrap.h).
#include "kmseardetect.cpp"
#include "ref_wrap.h"
REF_REGISTER(kms_ear_detect_plugin_init)
