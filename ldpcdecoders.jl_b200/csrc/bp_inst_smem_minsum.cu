#define BP_VARIANT 1
#include "bp_smem_inst.cuh"
