// Same header name and class name as the reference (frechet/include/NNFrechet.hpp): code written
// against NNFrechet::NNFrechet compiles against the GPU-backed planner (plus one setWorld call).
#ifndef NN_FRECHET_HPP_
#define NN_FRECHET_HPP_
#include "pr_gpu/NNFrechet.h"
namespace NNFrechet {
typedef pr_gpu::NNFrechet NNFrechet;
}
#endif
