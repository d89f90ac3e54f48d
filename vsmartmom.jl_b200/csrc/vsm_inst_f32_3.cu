#define VSM_T float
#define VSM_TM 3
#define VSM_NN 0
#define VSM_MINB 1
#define VSM_LTHR 256
#define VSM_ATHR 256
#include "vsm_inst.inc"
