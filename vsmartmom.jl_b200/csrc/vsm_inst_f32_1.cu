#define VSM_T float
#define VSM_TM 1
#include "vsm_inst.inc"
