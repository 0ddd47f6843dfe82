#include "comm.hpp"
#include <dlfcn.h>
#include <nccl.h>

namespace moai
{
    namespace
    {
        struct NcclApi
        {
            void *lib = nullptr;
            ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
            ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
            ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
            ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
            ncclResult_t (*GroupStart)() = nullptr;
            ncclResult_t (*GroupEnd)() = nullptr;
            const char *(*GetErrorString)(ncclResult_t) = nullptr;
        };

        NcclApi &api()
        {
            static NcclApi a = [] {
                NcclApi r;
                // RTLD_NOLOAD first: reuse the NCCL the process already has (torch's bundled copy)
                for (const char *name : { "libnccl.so.2", "libnccl.so" })
                {
                    r.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
                    if (!r.lib)
                    {
                        r.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
                    }
                    if (r.lib)
                    {
                        break;
                    }
                }
                if (!r.lib)
                {
                    return r;
                }
                r.GetUniqueId = (decltype(r.GetUniqueId))dlsym(r.lib, "ncclGetUniqueId");
                r.CommInitRank = (decltype(r.CommInitRank))dlsym(r.lib, "ncclCommInitRank");
                r.CommDestroy = (decltype(r.CommDestroy))dlsym(r.lib, "ncclCommDestroy");
                r.Broadcast = (decltype(r.Broadcast))dlsym(r.lib, "ncclBroadcast");
                r.GroupStart = (decltype(r.GroupStart))dlsym(r.lib, "ncclGroupStart");
                r.GroupEnd = (decltype(r.GroupEnd))dlsym(r.lib, "ncclGroupEnd");
                r.GetErrorString = (decltype(r.GetErrorString))dlsym(r.lib, "ncclGetErrorString");
                return r;
            }();
            if (!a.lib || !a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Broadcast || !a.GroupStart ||
                !a.GroupEnd)
            {
                throw StatusError{ LOGIC_ERROR, "NCCL (libnccl.so.2) is not available in this process" };
            }
            return a;
        }

        void nccl_check(ncclResult_t r, const char *what)
        {
            if (r != ncclSuccess)
            {
                const char *msg = api().GetErrorString ? api().GetErrorString(r) : "error";
                throw StatusError{ CUDA_ERROR, std::string("NCCL ") + what + ": " + msg };
            }
        }
    } // namespace

    void comm_unique_id(unsigned char out[128])
    {
        ncclUniqueId id;
        nccl_check(api().GetUniqueId(&id), "ncclGetUniqueId");
        static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
        memcpy(out, &id, 128);
    }

    void comm_init(Context *c, const unsigned char idb[128], int rank, int world)
    {
        MOAI_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world size");
        comm_destroy(c);
        if (world == 1)
        {
            return;
        }
        ncclUniqueId id;
        memcpy(&id, idb, 128);
        ncclComm_t comm;
        nccl_check(api().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
        c->comm = new Comm{ comm, rank, world };
    }

    void comm_destroy(Context *c)
    {
        if (c->comm)
        {
            if (c->comm->nccl)
            {
                api().CommDestroy((ncclComm_t)c->comm->nccl);
            }
            delete c->comm;
            c->comm = nullptr;
        }
    }

    void comm_all_gather_items(Context *c, u64 *buf, size_t item_words,
                               const std::vector<std::vector<std::pair<long long, long long>>> &owned)
    {
        Comm *cm = c->comm;
        if (!cm || cm->world == 1)
        {
            return;
        }
        MOAI_REQUIRE((int)owned.size() == cm->world, "one range list per rank expected");
        KernelTimer kt(c, "k_nccl_all_gather", 1);
        // a ragged all-gather = one broadcast per (rank, range), fused into one NCCL group; in place
        nccl_check(api().GroupStart(), "ncclGroupStart");
        for (int r = 0; r < cm->world; r++)
        {
            for (const auto &rg : owned[r])
            {
                if (rg.second <= rg.first)
                {
                    continue;
                }
                u64 *p = buf + (size_t)rg.first * item_words;
                const size_t cnt = (size_t)(rg.second - rg.first) * item_words;
                nccl_check(api().Broadcast(p, p, cnt, ncclUint64, r, (ncclComm_t)cm->nccl, c->stream), "ncclBroadcast");
                if (r != cm->rank)
                {
                    cm->gathered_bytes += cnt * sizeof(u64);
                }
            }
        }
        nccl_check(api().GroupEnd(), "ncclGroupEnd");
        cm->gathers += 1;
        c->launches += 1;
    }
} // namespace moai
