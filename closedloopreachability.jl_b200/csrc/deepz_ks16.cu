#define DZ_KS 16
#define DZ_NAME launch_deepz_ks16
#include "deepz_ks.inc"
