/* stand-in for the CMake-generated config.h of the reference plugins */
#ifndef VERSION
#define VERSION "0.0.0-mock"
#endif
#ifndef PACKAGE
#define PACKAGE "nubo-mock"
#endif
