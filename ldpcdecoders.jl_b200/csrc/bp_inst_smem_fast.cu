#define BP_VARIANT 2
#include "bp_smem_inst.cuh"
