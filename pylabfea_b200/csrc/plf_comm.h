// plf_comm.h -- the exchange step of the strip-partitioned solve: one NCCL communicator per context,
// one halo node column per neighbour and CG iteration, scalar all-reduces for the dot products.  NCCL is loaded with
// dlopen on first use so that the library itself has no link-time dependency on it (single-GPU users, CPU build check).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace plf {

struct Comm {
    void *nccl = nullptr;   // ncclComm_t
    int rank = 0, world = 1;
    // owned node columns [own_first, own_end) of the local mesh; a ghost column is the one column outside on either side
    int own_first = 0, own_end = 0;
    bool strip_set = false;
    bool mg_all = false;   // every rank's strip has a multigrid hierarchy (decided together in plf_comm_set_strip)
    // position of the strip in the global mesh (plf_comm_set_partition): needed by the load-increment controller for the
    // boundary conditions (edges, force shares, node sets) and the global averages
    bool global_set = false;
    int first_col_global = 0, NX_global = 0;
    double x_offset = 0., lenx_global = 0.;
    // peer mailboxes (comm_p2p_setup): every rank exposes one device buffer through CUDA IPC; halo columns and the scalars of
    // the all-reduces are then written straight into the other ranks' HBM over NVLink, one kernel per exchange
    bool p2p = false;
    unsigned char *box = nullptr;        // this rank's mailbox
    unsigned char *peer_box[64] = {};    // every rank's mailbox as mapped into this process ([rank] == box)
    int box_cols = 0;                    // doubles one halo slot holds
    unsigned long long halo_epoch = 0, red_epoch = 0;
    bool active() const { return nccl != nullptr; }
    bool ghost_left() const { return own_first > 0; }
    bool ghost_right(int NX) const { return own_end < NX + 1; }
};

// all return 0 on success; the message of the last failure is in comm_error()
int comm_load();                                            // dlopen libnccl.so.2 and resolve the entry points
int comm_unique_id(unsigned char *id128);                   // ncclGetUniqueId (128 bytes)
int comm_init(Comm &c, const unsigned char *id128, int rank, int world);
void comm_destroy(Comm &c);
int comm_allreduce_sum(Comm &c, double *dbuf, int count, cudaStream_t s);   // in place, device pointer
int comm_allreduce_op(Comm &c, double *dbuf, int count, int op /* 0 sum, 1 max, 2 min */, cudaStream_t s);
// vec: nodal vector of the local (NX+1) x (NY+1) node grid, 2 dofs per node.  Sends the first / last owned node column to
// the left / right neighbour and receives their last / first owned column into the ghost columns.
int comm_halo_exchange(Comm &c, double *vec, int NX, int NY, cudaStream_t s);
// the same exchange on a GLOBAL nodal vector (every rank holds all node columns 0 .. NX, owns [a, b)): column a goes to the
// left neighbour and column a - 1 comes back, column b - 1 goes to the right neighbour and column b comes back
int comm_halo_exchange_cols(Comm &c, double *vec, int NY, int a, int b, int NX, cudaStream_t s);
// Maps the mailboxes of all ranks (collective; needs NCCL for the handle exchange).  Leaves c.p2p false, and the exchanges on
// NCCL, when CUDA IPC is not available between the ranks or PLF_P2P=0 is set.  `cols`: doubles per node column.
int comm_p2p_setup(Comm &c, int cols, cudaStream_t s);
int comm_gather_ranges(Comm &c, const double *mine, double *global, const int64_t *goff, const int64_t *cnt, cudaStream_t s);
const char *comm_error();

}   // namespace plf
