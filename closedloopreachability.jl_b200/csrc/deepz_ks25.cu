#define DZ_KS 25
#define DZ_NAME launch_deepz_ks25
#include "deepz_ks.inc"
