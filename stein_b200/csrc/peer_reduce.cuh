// peer_reduce.cuh -- mailbox layout and entry points of the peer-memory all-reduce (peer_reduce.cu)
#pragma once
#include "common.cuh"

namespace stein {

constexpr int MB_RANKS = MAX_PEERS + 1;
constexpr int MB_BLOCKS = 16;                        // slices (CTAs, flags) per all-reduce
constexpr int MB_CAP = HIST_MAX_BINS + 8;            // 8-byte words per rank slot
constexpr size_t MB_FLAGS = (size_t)MB_RANKS * MB_BLOCKS;                   // flag words per parity
constexpr size_t MB_PARITY_WORDS = MB_FLAGS + (size_t)MB_RANKS * MB_CAP;    // flags, then the rank slots
// bytes an engine appends to its particle buffer (zero-initialised, exported with it through IPC)
constexpr size_t MBOX_BYTES = 2 * MB_PARITY_WORDS * 8;

struct PeerReduce;
// mailboxes[r]: rank r's mailbox as mapped on this GPU (own one included)
// epoch0: all-reduces already issued through these mailboxes (0 for fresh, zeroed mailboxes): the flags
// hold epoch numbers, so a re-opened connection must keep counting where the last one stopped
int peer_reduce_create(stein_ctx *ctx, int rank, int world, void *const *mailboxes, unsigned long long epoch0,
                       PeerReduce **out);
unsigned long long peer_reduce_epoch(const PeerReduce *pr);
void peer_reduce_destroy(PeerReduce *pr);

}  // namespace stein
