#define BP_VARIANT 0
#include "bp_smem_inst.cuh"
