#define VSM_T double
#define VSM_TM 2
#include "vsm_inst.inc"
