#define VSM_T float
#define VSM_TM 3
#include "vsm_inst.inc"
