#define VSM_T double
#define VSM_TM 1
#include "vsm_inst.inc"
