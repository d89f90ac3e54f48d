#define VSM_SNN 75
#define VSM_SWR 2
#define VSM_SWC 5
#define VSM_SKP 16
#define VSM_SNST 3
#define VSM_SMINB 2
#include "vsm_stream_inst.inc"
