// stand-in for <cuda_runtime.h> in the host-only sanitizer build of physecs_b200/csrc/batch.cpp (tests/abi_recorder/sanitize.sh):
// the batch driver's only CUDA call is cudaSetDevice in each shard thread
#pragma once
inline int cudaSetDevice(int) { return 0; }
