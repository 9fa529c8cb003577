// Transport internals shared by exchange.cpp (communicator, peer-memory heap) and exchange_ops.cpp
// (B200PartitionedOutput): the two phases of one exchange.
#pragma once
#include <memory>
#include <vector>

#include "../../../include/velox_b200.h"
#include "device.h"

namespace velox_b200 {

bool exchangeUsesPeerMemory(vb2_comm* c);
size_t exchangeMaxMetadataBytes(vb2_comm* c);

// Phase 1 — metadata: every rank's block (blockBytes, a multiple of 16; its first world * 8 bytes are
// replaced by countsDev, the device-side per-destination row counts) reaches every rank. Returns the
// world blocks in rank order in pinned host memory, after the exchange's ONE host synchronisation.
// Peer memory: put + flag + wait kernels; otherwise ncclAllGather. `after`: the stream that produced countsDev.
// patches: device-resident pieces copied into the block at the given offsets before it is sent.
struct ExchangePatch {
  size_t offset;
  const void* src;
  size_t bytes;
};
std::shared_ptr<void> exchangeMetadata(vb2_comm* c, const uint8_t* myBlockHost, size_t blockBytes, const int64_t* countsDev,
                                       const std::vector<ExchangePatch>& patches, cudaStream_t after);

// Phase 2 — payload: rows order[j] (NULL = identity) of the source columns, grouped by destination
// (matrix = world x world row counts from phase 1, identical on every rank), arrive in `outs` (one
// contiguous buffer per column, rows ordered by source rank). Peer memory when every block fits a
// heap segment (the gather is fused into the transfer kernel), else gather + grouped NCCL send/recv.
// Returns an event recorded when `outs` are complete. No host synchronisation.
std::shared_ptr<void> exchangePayload(vb2_comm* c, const int32_t* order, const int64_t* countsDev, const int64_t* matrix, int64_t n,
                                      const std::vector<const void*>& cols, const std::vector<int32_t>& widths, const std::vector<void*>& outs,
                                      bool broadcast, cudaStream_t after, bool* usedPeerMemory);

}  // namespace velox_b200
