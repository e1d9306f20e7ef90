#define BP_INST_MODE 2
#define BP_INST_BIG 1
#define BP_VARIANT 0
#include "bp_launch_inst.cuh"
