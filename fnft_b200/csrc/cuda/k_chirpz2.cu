// fnft_b200 -- translation unit that owns the chirp-z kernels (chirpz_kernels.cuh, chirpz2.cuh)
#define FNFTB_TU_CZ2
#include "chirpz_driver.cuh"
#include "chirpz2.cuh"
