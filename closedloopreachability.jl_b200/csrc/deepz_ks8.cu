#define DZ_KS 8
#define DZ_NAME launch_deepz_ks8
#include "deepz_ks.inc"
