// fnft_b200 -- translation unit that owns the low-level spectrum-carry tree kernels
// (tree_low2.cuh: NSE first-row-only mode, tree_low2g.cuh: general 2x2 mode)
#define FNFTB_TU_LOW2
#include "launch.cuh"
#include "tree_low2.cuh"
#include "tree_low2g.cuh"
