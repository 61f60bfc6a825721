/* stand-in: everything lives in minigst.h (test infrastructure, see tests/mock_gst/minigst.h) */
#include "minigst.h"
