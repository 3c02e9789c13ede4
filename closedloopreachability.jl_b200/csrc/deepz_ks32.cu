#define DZ_KS 32
#define DZ_NAME launch_deepz_ks32
#include "deepz_ks.inc"
