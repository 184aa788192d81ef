// fnft_b200 -- what inverse_api.cu (its own translation unit) borrows from the context of device_api.cu
#pragma once
#include "common.cuh"
#include "fnftb_device.h"
cudaStream_t fnftb__stream(fnftb_ctx *c);
int fnftb__activate(fnftb_ctx *c);
void **fnftb__inv_slot(fnftb_ctx *c, void (***dtor)(void *));
int fnftb__fail(int code, const char *what, const char *file, int line);
int fnftb__pair2x2_prepare(fnftb_ctx *c, size_t B, size_t d, cplx **lev0);
int fnftb__pair2x2_run(fnftb_ctx *c, size_t B, size_t d, const cplx **res);
int fnftb__dft(fnftb_ctx *c, size_t B, size_t n, const cplx *in_rev, cplx *out, int sign);
