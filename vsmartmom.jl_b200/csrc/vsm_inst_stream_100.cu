#define VSM_SNN 100
#define VSM_SWR 4
#define VSM_SWC 4
#define VSM_SKP 16
#define VSM_SNST 3
#define VSM_SMINB 1
#include "vsm_stream_inst.inc"
