// Multi-GPU exchange for one packed batch spread over G GPUs (SURVEY §8(e) "partitioning A").
//
// The independent units of the encoder layer are column ciphertexts: the 768 bootstrappings of every stage
// (M/test/test_full_scheme.hpp:654-660), the 12 heads (:530-533), the 3072 GELUs (:884-888) and the output columns of
// the ct-pt matmuls (Ct_pt_matrix_mul.hpp:20).  Rank r computes its share of them and the shares are exchanged with
// an all-gather of raw uint64 limbs over NVLink (NCCL); sums are modular, so there is no reduction collective.
// One process per GPU; the communicator belongs to the Context.  NCCL is resolved at run time from the copy the
// process has already loaded (torch's), so libmoai_b200.so has no link-time dependency on it.
#pragma once
#include "context.hpp"
#include <utility>
#include <vector>

namespace moai
{
    struct Comm
    {
        void *nccl = nullptr; // ncclComm_t
        int rank = 0, world = 1;
        unsigned long long gathers = 0, gathered_bytes = 0;
    };

    // balanced contiguous share of `count` units for `rank` of `world`: [first, last)
    inline std::pair<long long, long long> shard_range(long long count, int world, int rank)
    {
        const long long base = count / world, extra = count % world;
        const long long b0 = rank * base + (rank < extra ? rank : extra);
        return { b0, b0 + base + (rank < extra ? 1 : 0) };
    }

    void comm_unique_id(unsigned char out[128]);
    void comm_init(Context *c, const unsigned char id[128], int rank, int world);
    void comm_destroy(Context *c);

    // `buf` holds `items` of `item_words` uint64 each on every rank; rank r has produced the items in owned[r]
    // (a list of [first, last) ranges).  Afterwards every rank holds every item.  Stream-ordered on c->stream.
    void comm_all_gather_items(Context *c, u64 *buf, size_t item_words,
                               const std::vector<std::vector<std::pair<long long, long long>>> &owned);
} // namespace moai
