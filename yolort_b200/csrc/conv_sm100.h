// Host-side handle of one prepared convolution launch (tensor maps + kernel parameters).
#pragma once
#include <cuda_runtime.h>

#include "../../include/yolort_b200.h"

namespace yb {
struct ConvOp;
int conv_op_create(const yb_op_desc& d, ConvOp** out);
int conv_op_launch(const ConvOp* op, cudaStream_t stream);
void conv_op_destroy(ConvOp* op);

// HBM-bound helpers of the neck (pool_upsample.cu)
int spp_pool_launch(const yb_op_desc& d, cudaStream_t stream);
int upsample2x_launch(const yb_op_desc& d, cudaStream_t stream);
int validate_pool_or_upsample(const yb_op_desc& d);
}  // namespace yb
