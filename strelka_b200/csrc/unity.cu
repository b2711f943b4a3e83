// The whole backend as ONE translation unit: the __constant__ Sobol table and the SB_HD device
// functions are shared by the kernels and the C ABI without relocatable device code.
#include "kernels.cu"
#include "sb_api.cu"
