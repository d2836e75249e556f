// dist.h -- row-partitioned (DMDA) distribution: communicator, layout, halo exchange.
// Replaces what PETSc does inside MatMult_MPIAIJ / VecScatter / MPI_Allreduce for the reference's
// DMDACreate2d(PETSC_COMM_WORLD, ..., PETSC_DECIDE, PETSC_DECIDE, ...) partition (src/Discretization.c:17).
#pragma once
#include <condition_variable>
#include <mutex>
#include "core.h"

namespace b200sp {

// ---------------------------------------------------------------- communicator
// Two transports behind one interface:
//   NcclComm  one process per GPU (torchrun): ncclAllReduce / grouped ncclSend+ncclRecv on the compute streams
//   LocalComm all ranks are threads of ONE process (one or several GPUs): host barrier + cudaMemcpyPeerAsync.
//             Used by the tests (the whole distributed algorithm runs on a 1-GPU box, ranks never wait on each
//             other inside a kernel) and usable as a single-process multi-GPU mode.
constexpr int HALO_MAX_DOF = 3; // values per node a halo exchange can carry (2-D velocity 2, 3-D velocity 3, pressure 1)
struct HaloMsg { int peer; int64_t send_off, send_cnt, recv_off, recv_cnt; }; // in doubles

struct Comm {
  virtual ~Comm() {}
  virtual int rank() const = 0;
  virtual int size() const = 0;
  // in-place global sum of k doubles on the device, ordered on `s`
  virtual void allreduce_sum(double *d, int k, cudaStream_t s) = 0;
  // neighbour exchange: sendbuf/recvbuf are device arrays of this rank; message list is symmetric across ranks
  virtual void exchange(const double *sendbuf, double *recvbuf, const std::vector<HaloMsg> &msgs, cudaStream_t s) = 0;
  // gather `cnt` doubles from every rank into out[rank*cnt...] on every rank (redundant coarse multigrid levels)
  virtual void allgather(const double *in, double *out, int64_t cnt, cudaStream_t s) = 0;
  virtual void barrier() = 0;
  virtual bool capturable() const = 0; // may its calls be recorded into a CUDA graph (no host-side waits)?
  // one process per GPU on distinct GPUs: halos may be pushed straight into the neighbours' memory over NVLink
  // (CUDA IPC peer mappings + device-side flags) instead of going through send/recv
  virtual bool p2p_capable() const = 0;
};

// small-message collectives at a fixed use site (the Krylov reductions; the multigrid coarse-level bridge).  With a
// peer-to-peer capable communicator they are single kernels over CUDA-IPC mapped buffers (push to every peer, raise a
// flag, poll own flags, combine in RANK ORDER -> deterministic); otherwise they forward to the communicator.
struct Collective {
  virtual ~Collective() {}
  virtual void allreduce_sum(double *d, int k, cudaStream_t s) = 0;                              // k <= 32
  virtual void allgather(const double *in, double *out, int64_t cnt, cudaStream_t s) = 0;       // cnt <= capacity
};
Collective *make_collective(Ctx *c, int64_t capacity_doubles);

struct LocalGroup { // shared by the rank-threads of one process
  int size;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  std::vector<const double *> ptr_a;   // per rank: published device pointer (send buffer / gather input)
  std::vector<double *> ptr_b;         // per rank: published device pointer (reduction operand)
  std::vector<int> device;
  std::vector<double> host_scratch;    // [size][N_SCALARS]
  explicit LocalGroup(int n) : size(n), ptr_a((size_t)n), ptr_b((size_t)n), device((size_t)n, 0), host_scratch((size_t)n * N_SCALARS) {}
  void barrier();
};

Comm *make_local_comm(std::shared_ptr<LocalGroup> g, int rank, int device);
Comm *make_nccl_comm(ncclComm_t c, int rank, int size);

// ---------------------------------------------------------------- layout (host index arithmetic)
struct Layout {
  int M = 0, N = 0, size = 1, m = 1, n = 1;
  std::vector<int> lx, ly, xoff, yoff, rstart; // rstart[r] = first PETSc global node id of rank r
  Layout() {}
  Layout(int M_, int N_, int size_);                                                    // PETSC_DECIDE ownership
  Layout(int M_, int N_, int m_, int n_, const std::vector<int> &lx_, const std::vector<int> &ly_); // explicit ownership
  void finish();
  int owner_x(int i) const;
  int owner_y(int j) const;
  int owner(int i, int j) const { return owner_y(j) * m + owner_x(i); }
  int gnode(int i, int j) const;
  void box(int rank, int *xs, int *ys, int *xm, int *ym) const;
  // the layout of the next coarser DMDA: coarse node ic belongs to the owner of fine node 2*ic (DMCoarsen keeps the process grid)
  Layout coarsen() const;
};

// ---------------------------------------------------------------- halo plan (host index arithmetic only)
// Everything a rank needs to know about its box-stencil halo on a layout.  One implementation serves the device halo
// (make_halo), the host-only C ABI (b200sp_dmda_halo_plan / b200sp_dmda_halo_push_table) and therefore the CPU tests.
struct HaloPlan {
  int xs = 0, ys = 0, xm = 0, ym = 0, n_owned = 0;
  std::vector<int> ghost_gnode, ghost_owner, ghost_i, ghost_j; // ascending PETSc global node id (MPIAIJ garray order)
  std::vector<HaloMsg> msgs;      // one per neighbour, ascending rank; offsets/counts in NODES; receives are
                                  // contiguous ranges of the sorted ghost list
  std::vector<int> send_lnode;    // owned local node ids to send, grouped by message, in the receiver's ghost order
  // node-keyed view of the send lists (pushes fused into producing kernels, PushOut in core.h)
  std::vector<unsigned char> push_grp; // per 64 owned nodes: any of them sent?
  std::vector<int> push_node_ent;      // per owned node: (first entry << 3) | count (<= 7); 0 = not sent
  std::vector<int> push_ent_msg, push_ent_pos; // entry -> message index, position inside that message; entry 0 unused
  bool push_valid = true;              // false when some node has more than 3 destinations (boxes thinner than 2 nodes)
};
HaloPlan plan_halo(const Layout &L, int rank);

// ---------------------------------------------------------------- halo of one rank on one layout
struct Halo {
  Ctx *ctx = nullptr;
  int xs = 0, ys = 0, xm = 0, ym = 0, M = 0, N = 0;
  int n_owned = 0, n_ghost = 0;                   // nodes
  std::vector<int> ghost_gnode, ghost_i, ghost_j; // sorted by PETSc global node id (MPIAIJ garray order)
  std::vector<HaloMsg> node_msgs;                 // per neighbour, counts in NODES
  DevBuf<int> d_send_lnode;                       // owned local node ids to pack, grouped by neighbour
  int n_send = 0;
  DevBuf<int> d_ring2ghost;                       // ring position -> ghost index (-1 outside the domain)
  DevBuf<double> sendbuf, ghost;                  // sized for dof <= HALO_MAX_DOF (ghost holds two parities in peer-to-peer mode)
  cudaEvent_t ev_packed = nullptr, ev_arrived = nullptr;
  // ---- peer-to-peer mode (NVLink): the "pack" kernel stores every outgoing value DIRECTLY into the neighbour's ghost
  // buffer (CUDA IPC mapping) and then raises a sequence flag in the neighbour's memory; the receiver polls its own
  // flags in a one-warp kernel before the SpMV.  Ghost buffers are double-buffered by exchange parity: a sender can
  // reach exchange k+2 only after the receiver's push of exchange k+1, which is stream-ordered after the receiver's
  // SpMV of exchange k, so parity k is free again.  No NCCL call, no second stream, ~2 tiny kernels per exchange.
  bool p2p = false;
  int64_t ghost_stride = 0;                       // doubles per parity
  DevBuf<unsigned long long> seq, flags;          // exchanges started (device counter); one flag per incoming message
  DevBuf<unsigned> ticket;
  using P2PMsg = HaloP2PMsg;
  DevBuf<P2PMsg> d_p2p;                           // per outgoing message
  std::vector<void *> ipc_opened;
  // fused push (PushOut in core.h): which owned nodes go where, keyed by node
  std::vector<unsigned char> h_push_grp;          // host copies of the plan's push tables, uploaded by the peer-to-peer setup
  std::vector<int> h_push_node_ent, h_push_ent_msg, h_push_ent_pos;
  DevBuf<unsigned char> d_push_grp;
  DevBuf<int> d_push_node_ent;
  DevBuf<int2> d_push_ents;
  const double *pushed_vec = nullptr;             // a producer already pushed this vector for the next exchange
  PushOut push_out(int dof) const;                // descriptor for a producing kernel
  int n_msgs = 0;
  const double *ghost_now();                      // host: synchronise and return the parity that holds the latest halo
  ~Halo();
  // pack x (dof interleaved) and start the exchange on the halo stream; end() makes the compute stream wait for it
  void begin(const double *x, int dof);
  void end();
};
std::shared_ptr<Halo> make_halo(Ctx *c, const Layout &L, int rank);
std::shared_ptr<Halo> make_halo_general(Ctx *c, const Layout &L, int rank, const std::vector<int> &ghost_gnode); // collective

// device-side view of a column space (owned box + ghost ring) used by the assembly kernels to classify columns
struct ColSpace {
  int xs, ys, xm, ym;
  const int *ring2ghost; // null on one rank
  __host__ __device__ int ring_id(int i, int j) const {
    if (j == ys - 1) return i - (xs - 1);
    if (j == ys + ym) return (xm + 2) + i - (xs - 1);
    if (i == xs - 1) return 2 * (xm + 2) + (j - ys);
    return 2 * (xm + 2) + ym + (j - ys);
  }
};

} // namespace b200sp
