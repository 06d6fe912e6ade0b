// Cross-translation-unit access to the handle (struct rcb_ctx is private to b200chan.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/b200chan.h"

int rcb_internal_device(rcb_t* h);
cudaStream_t rcb_internal_stream(rcb_t* h);
int rcb_internal_fail(rcb_t* h, cudaError_t e, const char* what);  // records the message, returns RCB_ECUDA
void rcb_internal_count(rcb_t* h, int launches, size_t h2d_bytes, size_t d2h_bytes);
void** rcb_internal_post_slot(rcb_t* h);   // K6 state owned by postdemod_api.cu
void rcb_post_free_all(rcb_t* h);          // called from rcb_close
