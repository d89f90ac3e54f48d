#define VSM_T float
#define VSM_TM 4
#include "vsm_inst.inc"
