#include "../../include/b200lc.h"
#ifndef B200LC_GIT
#define B200LC_GIT "dev"
#endif
extern "C" const char *b200lc_version(void) { return "b200lc " B200LC_GIT " sm_100a"; }
