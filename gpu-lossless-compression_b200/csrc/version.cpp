#include "../../include/b200lc.h"
#ifndef B200LC_GIT
#define B200LC_GIT "dev"
#endif
extern "C" const char *b200lc_version(void) { return "b200lc " B200LC_GIT " sm_100a"; }

namespace b200lc {
unsigned &context_epoch()
{
    static unsigned epoch = 1;
    return epoch;
}
}  // namespace b200lc
