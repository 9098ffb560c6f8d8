// TEST INFRASTRUCTURE (oracle/_ref CPU build only).  pipeline.cpp references
// MemoryPool's destructor and reset() unconditionally, but they are defined only
// in the CUDA-only memory_pool.cu.  The CPU build never creates a pool.
#include "pcr/engine/memory_pool.h"

namespace pcr {
struct MemoryPool::Impl {};
MemoryPool::~MemoryPool() {}
void MemoryPool::reset() {}
}  // namespace pcr
