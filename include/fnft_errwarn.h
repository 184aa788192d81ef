/* Forwarder: programs written against FNFT's `fnft_errwarn.h` compile unchanged against
 * the fnft_b200 drop-in library -- every declaration lives in fnft_b200.h. */
#ifndef FNFT_B200_SHIM_FNFT_ERRWARN_H
#define FNFT_B200_SHIM_FNFT_ERRWARN_H
#include "fnft_b200.h"
#endif
