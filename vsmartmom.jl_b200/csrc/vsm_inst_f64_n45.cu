#define VSM_T double
#define VSM_TM 5
#define VSM_ATM 5
#define VSM_NN 45
#define VSM_MINB 3
#define VSM_LTHR 128
#define VSM_ATHR 384
#define VSM_ADD2 1
#define VSM_LTMI 5
#define VSM_LKS 1
#include "vsm_inst.inc"
