// Library identification for libogc_b200.so (see include/ogc_b200.h).
#include "common.cuh"

extern "C" const char *ogc_version(void) { return "ogc_b200 0.1.0 sm_100a"; }
