// fnft_b200 -- translation unit that owns the upper-level spectrum-carry tree kernels (tree_up.cuh)
#define FNFTB_TU_UP
#include "launch.cuh"
#include "tree_up.cuh"
