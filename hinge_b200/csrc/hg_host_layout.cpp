// placeholder until the maximal / layout stages land
#include <stdio.h>
#include "../../include/hinge_b200.h"
extern "C" int hg_main_maximal(int, char**) { fprintf(stderr, "hinge maximal: not built yet\n"); return 1; }
extern "C" int hg_main_layout(int, char**) { fprintf(stderr, "hinge layout: not built yet\n"); return 1; }
