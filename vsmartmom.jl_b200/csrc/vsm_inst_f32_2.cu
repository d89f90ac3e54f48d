#define VSM_T float
#define VSM_TM 2
#include "vsm_inst.inc"
