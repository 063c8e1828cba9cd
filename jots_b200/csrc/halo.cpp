// Interface exchange over NCCL (NVLink 5 / NVSwitch): grouped ncclSend/ncclRecv of the packed shared-DOF
// partial sums with every neighbour, canonical-order reduction on the device, and the scalar
// all-reduce behind every dot product.
#include "halo.hpp"

#include <nccl.h>

#include <stdexcept>
#include <string>

#include "ctx.hpp"
#include "partition.hpp"

namespace jots {

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};

static void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + ncclGetErrorString(r));
}

void comm_unique_id(char out[128]) {
    ncclUniqueId id;
    nccl_check(ncclGetUniqueId(&id), "ncclGetUniqueId");
    static_assert(sizeof(ncclUniqueId) <= 128, "unique id larger than 128 bytes");
    for (size_t i = 0; i < 128; ++i) out[i] = i < sizeof(id) ? id.internal[i] : 0;
}

Comm* comm_create(const char id_bytes[128], int rank, int nranks, int device) {
    ncclUniqueId id;
    for (size_t i = 0; i < sizeof(id); ++i) id.internal[i] = id_bytes[i];
    JOTS_CUDA(cudaSetDevice(device));
    Comm* c = new Comm;
    c->rank = rank; c->nranks = nranks;
    ncclResult_t r = ncclCommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete c; nccl_check(r, "ncclCommInitRank"); }
    return c;
}

void comm_destroy(Comm* c) {
    if (!c) return;
    if (c->comm) ncclCommDestroy(c->comm);
    delete c;
}
int comm_rank(const Comm* c) { return c->rank; }
int comm_size(const Comm* c) { return c->nranks; }

struct Halo {
    Comm* comm = nullptr;
    std::vector<int> nbr_rank, nbr_offset;
    int n_send = 0, n_red = 0;
    DArr<int> d_send_idx, d_red_dof, d_red_ptr, d_red_src;
    DArr<double> d_send, d_recv;
    DArr<int64_t> d_gid;
};

Halo* halo_create(const Partition& P, Comm* comm) {
    Halo* h = new Halo;
    h->comm = comm;
    h->nbr_rank = P.nbr_rank; h->nbr_offset = P.nbr_offset;
    h->n_send = (int)P.send_idx.size(); h->n_red = (int)P.red_dof.size();
    h->d_send_idx.upload(P.send_idx);
    h->d_red_dof.upload(P.red_dof); h->d_red_ptr.upload(P.red_ptr); h->d_red_src.upload(P.red_src);
    h->d_send.alloc(h->n_send > 0 ? h->n_send : 1);
    h->d_recv.alloc(h->n_send > 0 ? h->n_send : 1);
    h->d_gid.upload(P.global_id);
    return h;
}
void halo_destroy(Halo* h) {
    if (!h) return;
    delete h;
}

void halo_sum_exchange(Ctx& c, double* y) {
    Halo* h = c.halo;
    if (!h || h->nbr_rank.empty()) return;
    v_pack(c.stream, h->n_send, h->d_send_idx.p, y, h->d_send.p);
    nccl_check(ncclGroupStart(), "ncclGroupStart");
    for (size_t k = 0; k < h->nbr_rank.size(); ++k) {
        const int off = h->nbr_offset[k], cnt = h->nbr_offset[k + 1] - off;
        nccl_check(ncclSend(h->d_send.p + off, cnt, ncclDouble, h->nbr_rank[k], h->comm->comm, c.stream), "ncclSend");
        nccl_check(ncclRecv(h->d_recv.p + off, cnt, ncclDouble, h->nbr_rank[k], h->comm->comm, c.stream), "ncclRecv");
    }
    nccl_check(ncclGroupEnd(), "ncclGroupEnd");
    v_reduce_shared(c.stream, h->n_red, h->d_red_dof.p, h->d_red_ptr.p, h->d_red_src.p, h->d_recv.p, y);
    c.stats.kernel_launches += 2;
    ++c.stats.halo_exchanges;
}

void halo_allreduce(Ctx& c, double* d_vals, int nv) {
    Halo* h = c.halo;
    if (!h || h->comm->nranks == 1) return;
    nccl_check(ncclAllReduce(d_vals, d_vals, nv, ncclDouble, ncclSum, h->comm->comm, c.stream), "ncclAllReduce");
}

void halo_allreduce_max(Ctx& c, double* d_vals, int nv) {
    Halo* h = c.halo;
    if (!h || h->comm->nranks == 1) return;
    nccl_check(ncclAllReduce(d_vals, d_vals, nv, ncclDouble, ncclMax, h->comm->comm, c.stream), "ncclAllReduce(max)");
}
void halo_gather_owned(Ctx& c, const double* v, std::vector<int64_t>& gid, std::vector<double>& val) {
    Halo* h = c.halo;
    gid.clear(); val.clear();
    if (!h) return;
    const int nr = h->comm->nranks, me = h->comm->rank;
    // owned counts of every rank
    DArr<long long> d_cnt;
    d_cnt.alloc(nr);
    const long long mine = c.n_owned;
    JOTS_CUDA(cudaMemcpyAsync(d_cnt.p + me, &mine, sizeof(long long), cudaMemcpyHostToDevice, c.stream));
    nccl_check(ncclAllGather(d_cnt.p + me, d_cnt.p, 1, ncclInt64, h->comm->comm, c.stream), "ncclAllGather");
    std::vector<long long> cnt(nr);
    JOTS_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt.p, nr * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    JOTS_CUDA(cudaStreamSynchronize(c.stream));
    if (me == 0) {
        long long total = 0;
        for (long long k : cnt) total += k;
        DArr<long long> d_gid;
        DArr<double> d_val;
        d_gid.alloc(total); d_val.alloc(total);
        JOTS_CUDA(cudaMemcpyAsync(d_gid.p, h->d_gid.p, (size_t)cnt[0] * sizeof(long long), cudaMemcpyDeviceToDevice, c.stream));
        JOTS_CUDA(cudaMemcpyAsync(d_val.p, v, (size_t)cnt[0] * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
        nccl_check(ncclGroupStart(), "ncclGroupStart");
        long long off = cnt[0];
        for (int r = 1; r < nr; ++r) {
            nccl_check(ncclRecv(d_gid.p + off, cnt[r], ncclInt64, r, h->comm->comm, c.stream), "ncclRecv");
            nccl_check(ncclRecv(d_val.p + off, cnt[r], ncclDouble, r, h->comm->comm, c.stream), "ncclRecv");
            off += cnt[r];
        }
        nccl_check(ncclGroupEnd(), "ncclGroupEnd");
        gid.resize((size_t)total); val.resize((size_t)total);
        static_assert(sizeof(long long) == sizeof(int64_t), "64-bit ids");
        JOTS_CUDA(cudaMemcpyAsync(gid.data(), d_gid.p, (size_t)total * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
        JOTS_CUDA(cudaMemcpyAsync(val.data(), d_val.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        JOTS_CUDA(cudaStreamSynchronize(c.stream));
    } else {
        nccl_check(ncclGroupStart(), "ncclGroupStart");
        nccl_check(ncclSend(h->d_gid.p, mine, ncclInt64, 0, h->comm->comm, c.stream), "ncclSend");
        nccl_check(ncclSend(v, mine, ncclDouble, 0, h->comm->comm, c.stream), "ncclSend");
        nccl_check(ncclGroupEnd(), "ncclGroupEnd");
        JOTS_CUDA(cudaStreamSynchronize(c.stream));
    }
}
int halo_rank(const Ctx& c) { return c.halo ? c.halo->comm->rank : 0; }

const int64_t* halo_global_index(Ctx& c) { return c.halo ? c.halo->d_gid.p : nullptr; }

}  // namespace jots
