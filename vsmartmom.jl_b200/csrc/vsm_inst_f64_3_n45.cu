#define VSM_T double
#define VSM_TM 3
#define VSM_NN 45
#define VSM_MINB 3
#include "vsm_inst.inc"
