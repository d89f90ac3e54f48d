#define VSM_T float
#define VSM_TM 1
#define VSM_NN 0
#define VSM_MINB 1
#include "vsm_inst.inc"
