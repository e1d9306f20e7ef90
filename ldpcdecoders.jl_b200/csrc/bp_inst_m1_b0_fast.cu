#define BP_INST_MODE 1
#define BP_INST_BIG 0
#define BP_VARIANT 2
#include "bp_launch_inst.cuh"
