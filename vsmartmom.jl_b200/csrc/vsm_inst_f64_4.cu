#define VSM_T double
#define VSM_TM 4
#include "vsm_inst.inc"
