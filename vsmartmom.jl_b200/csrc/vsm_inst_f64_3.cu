#define VSM_T double
#define VSM_TM 3
#include "vsm_inst.inc"
