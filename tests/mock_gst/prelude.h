// prelude.h — forced include (-include) of the mock builds: routes clock() and gettimeofday() to the harness so that the
// tests inject time (the reference stamps its motion history with clock(), gstnubotracker.cpp:349, and rate-limits its
// signals with gettimeofday(), kmsfacedetect.cpp:228-236).  The standard headers that mention these names are pulled in
// first, while the names still mean the C library's.  TEST INFRASTRUCTURE ONLY.
#ifndef MOCK_PRELUDE_H
#define MOCK_PRELUDE_H
#include <sys/time.h>
#include <time.h>
#ifdef __cplusplus
#include <chrono>
#include <ctime>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
extern "C" {
#endif
clock_t mh_hook_clock(void);
int mh_hook_gettimeofday(struct timeval *tv, void *tz);
#ifdef __cplusplus
}
#endif
#define clock mh_hook_clock
#define gettimeofday mh_hook_gettimeofday
#endif
