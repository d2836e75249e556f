// dist.cu -- communicators (NCCL / in-process thread group), DMDA layout arithmetic, halo plan and exchange.
#include "dist.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "dev.cuh"

namespace b200sp {

// ------------------------------------------------------------------ LocalGroup / LocalComm
void LocalGroup::barrier() {
  std::unique_lock<std::mutex> lk(mu);
  const uint64_t gen = generation;
  if (++arrived == size) {
    arrived = 0;
    ++generation;
    cv.notify_all();
  } else {
    cv.wait(lk, [&] { return generation != gen; });
  }
}

namespace {

struct LocalComm : Comm {
  std::shared_ptr<LocalGroup> g;
  int r, dev;
  LocalComm(std::shared_ptr<LocalGroup> g_, int rank, int device) : g(g_), r(rank), dev(device) { g->device[(size_t)rank] = device; }
  int rank() const override { return r; }
  int size() const override { return g->size; }
  void barrier() override { g->barrier(); }
  bool capturable() const override { return false; }
  bool p2p_capable() const override { return false; } // rank-threads may share one GPU: kernels must not wait on each other
  void allreduce_sum(double *d, int k, cudaStream_t s) override {
    B2_REQUIRE(k <= N_SCALARS, "allreduce: too many scalars");
    double *mine = g->host_scratch.data() + (size_t)r * N_SCALARS;
    B2_CUDA(cudaMemcpyAsync(mine, d, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    g->barrier();
    double tot[N_SCALARS];
    for (int j = 0; j < k; ++j) { // fixed rank order: deterministic
      double t = 0.0;
      for (int q = 0; q < g->size; ++q) t += g->host_scratch[(size_t)q * N_SCALARS + j];
      tot[j] = t;
    }
    g->barrier(); // everyone has read before anyone overwrites host_scratch again
    B2_CUDA(cudaMemcpyAsync(d, tot, sizeof(double) * (size_t)k, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaStreamSynchronize(s)); // tot lives on this stack frame
  }
  void exchange(const double *sendbuf, double *recvbuf, const std::vector<HaloMsg> &msgs, cudaStream_t s) override {
    B2_CUDA(cudaStreamSynchronize(s)); // my packed data is complete
    g->ptr_a[(size_t)r] = sendbuf;
    g->ptr_b[(size_t)r] = reinterpret_cast<double *>(const_cast<std::vector<HaloMsg> *>(&msgs));
    g->barrier();
    for (const HaloMsg &m : msgs) {
      if (m.recv_cnt == 0) continue;
      // the peer's send offset for me is in the peer's message list
      const auto *pm = reinterpret_cast<const std::vector<HaloMsg> *>(g->ptr_b[(size_t)m.peer]);
      int64_t off = -1;
      for (const HaloMsg &q : *pm) if (q.peer == r) { off = q.send_off; B2_REQUIRE(q.send_cnt == m.recv_cnt, "halo: asymmetric message sizes"); }
      B2_REQUIRE(off >= 0, "halo: peer has no message for this rank");
      B2_CUDA(cudaMemcpyPeerAsync(recvbuf + m.recv_off, dev, g->ptr_a[(size_t)m.peer] + off, g->device[(size_t)m.peer], sizeof(double) * (size_t)m.recv_cnt, s));
    }
    B2_CUDA(cudaStreamSynchronize(s));
    g->barrier(); // peers may now reuse their send buffers
  }
  void allgather(const double *in, double *out, int64_t cnt, cudaStream_t s) override {
    B2_CUDA(cudaStreamSynchronize(s));
    g->ptr_a[(size_t)r] = in;
    g->barrier();
    for (int q = 0; q < g->size; ++q)
      B2_CUDA(cudaMemcpyPeerAsync(out + (size_t)q * cnt, dev, g->ptr_a[(size_t)q], g->device[(size_t)q], sizeof(double) * (size_t)cnt, s));
    B2_CUDA(cudaStreamSynchronize(s));
    g->barrier();
  }
};

struct NcclComm : Comm {
  ncclComm_t c;
  int r, n;
  NcclComm(ncclComm_t c_, int rank, int size) : c(c_), r(rank), n(size) {}
  int rank() const override { return r; }
  int size() const override { return n; }
  void barrier() override {}
  bool capturable() const override { return true; }
  bool p2p_capable() const override { static const bool off = getenv("B200SP_NO_P2P") && atoi(getenv("B200SP_NO_P2P")); return !off; }
  void allreduce_sum(double *d, int k, cudaStream_t s) override { B2_NCCL(nccl().AllReduce(d, d, (size_t)k, ncclDouble, ncclSum, c, s)); }
  void exchange(const double *sendbuf, double *recvbuf, const std::vector<HaloMsg> &msgs, cudaStream_t s) override {
    B2_NCCL(nccl().GroupStart());
    for (const HaloMsg &m : msgs) {
      if (m.send_cnt) B2_NCCL(nccl().Send(sendbuf + m.send_off, (size_t)m.send_cnt, ncclDouble, m.peer, c, s));
      if (m.recv_cnt) B2_NCCL(nccl().Recv(recvbuf + m.recv_off, (size_t)m.recv_cnt, ncclDouble, m.peer, c, s));
    }
    B2_NCCL(nccl().GroupEnd());
  }
  void allgather(const double *in, double *out, int64_t cnt, cudaStream_t s) override {
    // grouped send/recv (ncclAllGather is not in the dlopen table; this is the same wire pattern)
    B2_NCCL(nccl().GroupStart());
    for (int q = 0; q < n; ++q) {
      B2_NCCL(nccl().Send(in, (size_t)cnt, ncclDouble, q, c, s));
      B2_NCCL(nccl().Recv(out + (size_t)q * cnt, (size_t)cnt, ncclDouble, q, c, s));
    }
    B2_NCCL(nccl().GroupEnd());
  }
};

} // namespace

Comm *make_local_comm(std::shared_ptr<LocalGroup> g, int rank, int device) { return new LocalComm(g, rank, device); }
Comm *make_nccl_comm(ncclComm_t c, int rank, int size) { return new NcclComm(c, rank, size); }

// ------------------------------------------------------------------ peer-to-peer small collectives
namespace {
struct SymPeer { double *rb; unsigned long long *flags; };

// all-reduce of k <= 32 doubles in ONE single-block kernel: push my values into every rank's receive slot (parity of
// this call), fence, raise my flag on every rank, poll my own flags, sum the slots in rank order.
__global__ void __launch_bounds__(256) k_sym_allreduce(double *d, int k, int size, int rank, int64_t cap, const SymPeer *__restrict__ peers,
                                                       double *rb, unsigned long long *flags, unsigned long long *seq, int *err) {
  __shared__ unsigned long long s_k;
  if (threadIdx.x == 0) s_k = *seq;
  __syncthreads();
  const unsigned long long kk = s_k, par = kk & 1ull;
  for (int t = threadIdx.x; t < size * k; t += blockDim.x) {
    const int q = t / k, j = t % k;
    peers[q].rb[((int64_t)par * size + rank) * cap + j] = d[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < size) {
    volatile unsigned long long *f = peers[threadIdx.x].flags + rank;
    *f = kk + 1ull;
    volatile const unsigned long long *mine = flags + threadIdx.x;
    for (long long spin = 0; *mine < kk + 1ull; ++spin) {
      __nanosleep(100);
      if (spin > 20000000LL) { *err = 200 + (int)threadIdx.x; __threadfence_system(); break; }
    }
    __threadfence_system();
  }
  __syncthreads();
  if ((int)threadIdx.x < k) {
    double sum = 0.0;
    for (int r = 0; r < size; ++r) sum += ((volatile double *)rb)[((int64_t)par * size + r) * cap + threadIdx.x];
    d[threadIdx.x] = sum;
  }
  if (threadIdx.x == 0) *seq = kk + 1ull;
}
__global__ void __launch_bounds__(256) k_sym_gather_push(const double *__restrict__ in, int64_t cnt, int size, int rank, int64_t cap,
                                                         const SymPeer *__restrict__ peers, unsigned long long *seq, unsigned *ticket) {
  __shared__ bool s_last;
  const unsigned long long kk = *seq, par = kk & 1ull;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)size * cnt; t += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(t / cnt);
    const int64_t i = t % cnt;
    peers[q].rb[((int64_t)par * size + rank) * cap + i] = in[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence_system(); s_last = atomicAdd(ticket, 1u) == gridDim.x - 1; }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < size) { volatile unsigned long long *f = peers[threadIdx.x].flags + rank; *f = kk + 1ull; }
  if (threadIdx.x == 0) { *ticket = 0u; *seq = kk + 1ull; }
}
__global__ void __launch_bounds__(256) k_sym_gather_wait_copy(double *out, int64_t cnt, int size, int64_t cap, const double *rb,
                                                              const unsigned long long *flags, const unsigned long long *seq, int *err) {
  const unsigned long long want = *seq, par = (want - 1ull) & 1ull;
  if ((int)threadIdx.x < size) {
    volatile const unsigned long long *f = flags + threadIdx.x;
    for (long long spin = 0; *f < want; ++spin) {
      __nanosleep(100);
      if (spin > 20000000LL) { *err = 300 + (int)threadIdx.x; __threadfence_system(); break; }
    }
    __threadfence_system();
  }
  __syncthreads();
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)size * cnt; t += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(t / cnt);
    const int64_t i = t % cnt;
    out[t] = ((volatile const double *)rb)[((int64_t)par * size + q) * cap + i];
  }
}

struct FallbackCollective : Collective {
  Ctx *c;
  explicit FallbackCollective(Ctx *ctx) : c(ctx) {}
  void allreduce_sum(double *d, int k, cudaStream_t s) override { c->dcomm->allreduce_sum(d, k, s); }
  void allgather(const double *in, double *out, int64_t cnt, cudaStream_t s) override { c->dcomm->allgather(in, out, cnt, s); }
};

struct SymCollective : Collective {
  Ctx *c;
  int size, rank;
  int64_t cap;
  DevBuf<double> rb;
  DevBuf<unsigned long long> flags, seq;
  DevBuf<unsigned> ticket;
  DevBuf<SymPeer> peers;
  std::vector<void *> opened;
  SymCollective(Ctx *ctx, int64_t cap_) : c(ctx), size(ctx->size), rank(ctx->rank), cap((cap_ + 15) & ~15LL) {
    B2_REQUIRE(size <= 64, "p2p collective: at most 64 ranks");
    rb.alloc((size_t)2 * size * cap);
    rb.zero(c->stream);
    flags.alloc((size_t)size); flags.zero(c->stream);
    seq.alloc(1); seq.zero(c->stream);
    ticket.alloc(1); ticket.zero(c->stream);
    struct Rec { cudaIpcMemHandle_t hb, hf; };
    constexpr int REC = 16; // doubles
    static_assert(sizeof(Rec) <= REC * sizeof(double), "record too large");
    Rec mine;
    B2_CUDA(cudaIpcGetMemHandle(&mine.hb, rb.p));
    B2_CUDA(cudaIpcGetMemHandle(&mine.hf, flags.p));
    DevBuf<double> d_in(REC), d_all((size_t)REC * size);
    std::vector<double> h_in(REC, 0.0), h_all((size_t)REC * size);
    std::memcpy(h_in.data(), &mine, sizeof(mine));
    B2_CUDA(cudaMemcpyAsync(d_in.p, h_in.data(), sizeof(double) * REC, cudaMemcpyHostToDevice, c->stream));
    c->dcomm->allgather(d_in.p, d_all.p, REC, c->stream);
    B2_CUDA(cudaMemcpyAsync(h_all.data(), d_all.p, sizeof(double) * h_all.size(), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    std::vector<SymPeer> pp((size_t)size);
    for (int q = 0; q < size; ++q) {
      if (q == rank) { pp[(size_t)q] = SymPeer{rb.p, flags.p}; continue; }
      Rec pr;
      std::memcpy(&pr, h_all.data() + (size_t)REC * q, sizeof(pr));
      void *pb = nullptr, *pf = nullptr;
      B2_CUDA(cudaIpcOpenMemHandle(&pb, pr.hb, cudaIpcMemLazyEnablePeerAccess));
      B2_CUDA(cudaIpcOpenMemHandle(&pf, pr.hf, cudaIpcMemLazyEnablePeerAccess));
      opened.push_back(pb); opened.push_back(pf);
      pp[(size_t)q] = SymPeer{(double *)pb, (unsigned long long *)pf};
    }
    peers.alloc((size_t)size);
    B2_CUDA(cudaMemcpyAsync(peers.p, pp.data(), sizeof(SymPeer) * pp.size(), cudaMemcpyHostToDevice, c->stream));
    c->sync();
  }
  ~SymCollective() override { for (void *p : opened) cudaIpcCloseMemHandle(p); }
  void allreduce_sum(double *d, int k, cudaStream_t s) override {
    for (int j0 = 0; j0 < k; j0 += 32) { // 32 values per kernel
      const int kk = k - j0 < 32 ? k - j0 : 32;
      c->launches++;
      k_sym_allreduce<<<1, 256, 0, s>>>(d + j0, kk, size, rank, cap, peers.p, rb.p, flags.p, seq.p, c->d_err);
      check_launch("k_sym_allreduce");
    }
  }
  void allgather(const double *in, double *out, int64_t cnt, cudaStream_t s) override {
    B2_REQUIRE(cnt <= cap, "p2p allgather: message larger than the registered capacity");
    int grid = (int)std::min<int64_t>(32, ((int64_t)size * cnt + 2047) / 2048);
    if (grid < 1) grid = 1;
    c->launches += 2;
    k_sym_gather_push<<<grid, 256, 0, s>>>(in, cnt, size, rank, cap, peers.p, seq.p, ticket.p);
    check_launch("k_sym_gather_push");
    k_sym_gather_wait_copy<<<grid, 256, 0, s>>>(out, cnt, size, cap, rb.p, flags.p, seq.p, c->d_err);
    check_launch("k_sym_gather_wait_copy");
  }
};
} // namespace

Collective *make_collective(Ctx *c, int64_t capacity_doubles) {
  B2_REQUIRE(c->dcomm, "make_collective: context has no communicator");
  if (c->dcomm->p2p_capable()) return new SymCollective(c, capacity_doubles < 32 ? 32 : capacity_doubles);
  return new FallbackCollective(c);
}

// ------------------------------------------------------------------ Layout
Layout::Layout(int M_, int N_, int size_) : M(M_), N(N_), size(size_) {
  B2_REQUIRE(M >= 2 && N >= 2 && size >= 1, "dmda: need M,N >= 2 and size >= 1");
  dmda_proc_grid(M, N, size, &m, &n);
  B2_REQUIRE(m * n == size, "dmda: size does not factor into a process grid");
  B2_REQUIRE(m <= M && n <= N, "dmda: more ranks than nodes in a direction");
  lx.resize((size_t)m);
  ly.resize((size_t)n);
  dmda_ownership(M, m, lx.data());
  dmda_ownership(N, n, ly.data());
  finish();
}
Layout::Layout(int M_, int N_, int m_, int n_, const std::vector<int> &lx_, const std::vector<int> &ly_)
    : M(M_), N(N_), size(m_ * n_), m(m_), n(n_), lx(lx_), ly(ly_) {
  finish();
}
void Layout::finish() {
  xoff.assign((size_t)m + 1, 0);
  yoff.assign((size_t)n + 1, 0);
  for (int i = 0; i < m; ++i) xoff[(size_t)i + 1] = xoff[(size_t)i] + lx[(size_t)i];
  for (int j = 0; j < n; ++j) yoff[(size_t)j + 1] = yoff[(size_t)j] + ly[(size_t)j];
  B2_REQUIRE(xoff[(size_t)m] == M && yoff[(size_t)n] == N, "dmda: ownership ranges do not sum to the grid size");
  rstart.assign((size_t)size + 1, 0);
  for (int r = 0; r < size; ++r) rstart[(size_t)r + 1] = rstart[(size_t)r] + lx[(size_t)(r % m)] * ly[(size_t)(r / m)];
}
int Layout::owner_x(int i) const { return (int)(std::upper_bound(xoff.begin(), xoff.end(), i) - xoff.begin()) - 1; }
int Layout::owner_y(int j) const { return (int)(std::upper_bound(yoff.begin(), yoff.end(), j) - yoff.begin()) - 1; }
int Layout::gnode(int i, int j) const {
  const int pi = owner_x(i), pj = owner_y(j), r = pj * m + pi;
  return rstart[(size_t)r] + (j - yoff[(size_t)pj]) * lx[(size_t)pi] + (i - xoff[(size_t)pi]);
}
void Layout::box(int rank, int *xs, int *ys, int *xm, int *ym) const {
  const int pi = rank % m, pj = rank / m;
  *xs = xoff[(size_t)pi]; *ys = yoff[(size_t)pj]; *xm = lx[(size_t)pi]; *ym = ly[(size_t)pj];
}
Layout Layout::coarsen() const {
  B2_REQUIRE((M - 1) % 2 == 0 && (N - 1) % 2 == 0, "dmda: grid not coarsenable");
  const int Mc = (M - 1) / 2 + 1, Nc = (N - 1) / 2 + 1;
  std::vector<int> clx((size_t)m), cly((size_t)n);
  for (int i = 0; i < m; ++i) { // coarse ic owned by the owner of fine 2*ic: ic in [ceil(xs/2), floor((xe-1)/2)]
    const int lo = (xoff[(size_t)i] + 1) / 2, hi = (xoff[(size_t)i + 1] - 1) / 2;
    clx[(size_t)i] = hi - lo + 1;
  }
  for (int j = 0; j < n; ++j) {
    const int lo = (yoff[(size_t)j] + 1) / 2, hi = (yoff[(size_t)j + 1] - 1) / 2;
    cly[(size_t)j] = hi - lo + 1;
  }
  for (int v : clx) B2_REQUIRE(v >= 2, "dmda: a rank would own fewer than 2 coarse nodes in x; use fewer distributed levels");
  for (int v : cly) B2_REQUIRE(v >= 2, "dmda: a rank would own fewer than 2 coarse nodes in y; use fewer distributed levels");
  return Layout(Mc, Nc, m, n, clx, cly);
}

// ------------------------------------------------------------------ Halo
namespace {
__global__ void __launch_bounds__(256) k_pack(int n, int dof, const int *__restrict__ lnode, const double *__restrict__ x, double *buf) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * dof; t += gridDim.x * blockDim.x) {
    const int k = t / dof, c = t % dof;
    buf[t] = x[(size_t)lnode[k] * dof + c];
  }
}
} // namespace

// push: every outgoing value goes straight into the neighbour's ghost buffer (parity of this exchange); the last
// block to finish (after a system-scope fence) raises the sequence flag in every neighbour's memory.
__global__ void __launch_bounds__(256) k_halo_push(int n_send, int dof, int nmsg, const int *__restrict__ lnode, const double *__restrict__ x,
                                                   const Halo::P2PMsg *__restrict__ msgs, unsigned long long *seq, unsigned *ticket) {
  __shared__ bool s_last;
  const unsigned long long k = *seq; // exchanges completed before this one
  const unsigned long long par = k & 1ull;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_send * dof; t += gridDim.x * blockDim.x) {
    const int node = t / dof, c = t % dof;
    int m = 0;
    while (m + 1 < nmsg && node >= msgs[m].send_off + msgs[m].send_cnt) ++m;
    const Halo::P2PMsg &g = msgs[m];
    g.peer_ghost[par * g.peer_stride + (long long)(g.peer_recv_off + node - g.send_off) * dof + c] = x[(size_t)lnode[node] * dof + c];
  }
  // one system-scope fence per block (fences are cumulative: the barrier orders the block's stores before thread 0's
  // fence), not one per thread -- per-thread fences made this kernel ~16 us
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < nmsg) {
    volatile unsigned long long *f = msgs[threadIdx.x].peer_flag;
    *f = k + 1ull;
  }
  if (threadIdx.x == 0) { *ticket = 0u; *seq = k + 1ull; }
}
// wait: one lane per incoming message polls this rank's own flag word until the neighbour's push of the current
// exchange has landed.  Bounded: after ~2 s of polling it reports through the context's mapped error word.
__global__ void k_halo_wait(int nmsg, const unsigned long long *seq, const unsigned long long *flags, int *err) {
  if ((int)threadIdx.x >= nmsg) return;
  const unsigned long long want = *seq;
  volatile const unsigned long long *f = flags + threadIdx.x;
  for (long long spin = 0; *f < want; ++spin) {
    __nanosleep(200);
    if (spin > 10000000LL) { *err = 100 + (int)threadIdx.x; __threadfence_system(); return; }
  }
}

Halo::~Halo() {
  for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
  if (ev_packed) cudaEventDestroy(ev_packed);
  if (ev_arrived) cudaEventDestroy(ev_arrived);
}

PushOut Halo::push_out(int dof) const {
  PushOut o;
  o.grp = d_push_grp.p; o.node_ent = d_push_node_ent.p; o.ents = d_push_ents.p; o.msgs = d_p2p.p;
  o.seq = seq.p; o.ticket = ticket.p; o.nmsg = n_msgs; o.dof = dof;
  return o;
}

const double *Halo::ghost_now() {
  if (!p2p) return ghost.p;
  unsigned long long k = 0;
  B2_CUDA(cudaMemcpyAsync(&k, seq.p, sizeof(k), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->sync();
  return ghost.p + ((k - 1ull) & 1ull) * ghost_stride;
}

// Peer-to-peer setup: every rank publishes {IPC handle of its ghost buffer, IPC handle of its flag words, stride, and
// for each incoming message the sender rank and the receive offset}; neighbours open the handles and remember where
// to write.  The records travel through the communicator's all-gather (512 bytes per rank), no host side channel.
static void setup_p2p(Halo &h, const Layout &L, int rank) {
  Ctx *c = h.ctx;
  constexpr int REC = 64; // doubles per record
  struct Rec { cudaIpcMemHandle_t hg, hf; long long stride; int nmsg; int peer[8]; int recv_off[8]; };
  static_assert(sizeof(Rec) <= REC * sizeof(double), "record too large");
  h.n_msgs = (int)h.node_msgs.size();
  B2_REQUIRE(h.n_msgs <= 8, "halo: more than 8 neighbours");
  h.ghost_stride = ((int64_t)h.n_ghost * HALO_MAX_DOF + 2 + 15) & ~15LL;
  h.ghost.alloc((size_t)h.ghost_stride * 2);
  h.ghost.zero(c->stream);
  h.seq.alloc(1); h.seq.zero(c->stream);
  h.flags.alloc(8); h.flags.zero(c->stream);
  h.ticket.alloc(1); h.ticket.zero(c->stream);
  Rec mine;
  std::memset(&mine, 0, sizeof(mine));
  B2_CUDA(cudaIpcGetMemHandle(&mine.hg, h.ghost.p));
  B2_CUDA(cudaIpcGetMemHandle(&mine.hf, h.flags.p));
  mine.stride = h.ghost_stride;
  mine.nmsg = h.n_msgs;
  for (int m = 0; m < h.n_msgs; ++m) { mine.peer[m] = h.node_msgs[(size_t)m].peer; mine.recv_off[m] = (int)h.node_msgs[(size_t)m].recv_off; }
  DevBuf<double> d_in(REC), d_all((size_t)REC * L.size);
  std::vector<double> h_in(REC, 0.0), h_all((size_t)REC * L.size);
  std::memcpy(h_in.data(), &mine, sizeof(mine));
  B2_CUDA(cudaMemcpyAsync(d_in.p, h_in.data(), sizeof(double) * REC, cudaMemcpyHostToDevice, c->stream));
  c->dcomm->allgather(d_in.p, d_all.p, REC, c->stream);
  B2_CUDA(cudaMemcpyAsync(h_all.data(), d_all.p, sizeof(double) * h_all.size(), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  std::vector<Halo::P2PMsg> pm((size_t)h.n_msgs);
  for (int m = 0; m < h.n_msgs; ++m) {
    const HaloMsg &hm = h.node_msgs[(size_t)m];
    Rec pr;
    std::memcpy(&pr, h_all.data() + (size_t)REC * hm.peer, sizeof(pr));
    void *pg = nullptr, *pf = nullptr;
    B2_CUDA(cudaIpcOpenMemHandle(&pg, pr.hg, cudaIpcMemLazyEnablePeerAccess));
    B2_CUDA(cudaIpcOpenMemHandle(&pf, pr.hf, cudaIpcMemLazyEnablePeerAccess));
    h.ipc_opened.push_back(pg);
    h.ipc_opened.push_back(pf);
    int slot = -1;
    for (int j = 0; j < pr.nmsg; ++j) if (pr.peer[j] == rank) slot = j;
    B2_REQUIRE(slot >= 0, "halo p2p: neighbour has no message slot for this rank");
    Halo::P2PMsg &q = pm[(size_t)m];
    q.peer_ghost = (double *)pg;
    q.peer_flag = (unsigned long long *)pf + slot;
    q.peer_stride = pr.stride;
    q.send_off = (int)hm.send_off;
    q.send_cnt = (int)hm.send_cnt;
    q.peer_recv_off = pr.recv_off[slot];
    q.pad = 0;
  }
  { // node-keyed send tables for pushes fused into producing kernels (built by plan_halo)
    std::vector<int2> ents(h.h_push_ent_msg.size());
    for (size_t e = 0; e < ents.size(); ++e) ents[e] = make_int2(h.h_push_ent_msg[e], h.h_push_ent_pos[e]);
    h.d_push_grp.alloc(h.h_push_grp.size() + 1);
    h.d_push_node_ent.alloc(h.h_push_node_ent.size() + 1);
    h.d_push_ents.alloc(ents.size() + 1);
    B2_CUDA(cudaMemcpyAsync(h.d_push_grp.p, h.h_push_grp.data(), h.h_push_grp.size(), cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(h.d_push_node_ent.p, h.h_push_node_ent.data(), sizeof(int) * h.h_push_node_ent.size(), cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaMemcpyAsync(h.d_push_ents.p, ents.data(), sizeof(int2) * ents.size(), cudaMemcpyHostToDevice, c->stream));
    c->sync();
  }
  h.d_p2p.alloc((size_t)h.n_msgs + 1);
  if (h.n_msgs) B2_CUDA(cudaMemcpyAsync(h.d_p2p.p, pm.data(), sizeof(Halo::P2PMsg) * pm.size(), cudaMemcpyHostToDevice, c->stream));
  c->sync();
  h.p2p = true;
}

// node-keyed view of the send lists (pushes fused into producing kernels, PushOut in core.h)
static void build_push_tables(HaloPlan &P) {
  const int n = P.n_owned;
  P.push_grp.assign((size_t)(n + 63) / 64 + 1, 0);
  P.push_node_ent.assign((size_t)n + 1, 0);
  std::vector<int> cnt((size_t)n + 1, 0), first((size_t)n + 1, 0), fill((size_t)n + 1, 0);
  for (const HaloMsg &m : P.msgs)
    for (int64_t k = 0; k < m.send_cnt; ++k) cnt[(size_t)P.send_lnode[(size_t)(m.send_off + k)]]++;
  int total = 1; // entry 0 is never used so that "0" can mean "none"
  for (int v = 0; v < n; ++v) {
    if (cnt[(size_t)v] > 7) P.push_valid = false; // more destinations than the 3-bit count holds (boxes thinner than 2 nodes): no fused push, no peer-to-peer
    first[(size_t)v] = total;
    total += cnt[(size_t)v];
  }
  P.push_ent_msg.assign((size_t)total, 0);
  P.push_ent_pos.assign((size_t)total, 0);
  for (size_t mi = 0; mi < P.msgs.size(); ++mi)
    for (int64_t k = 0; k < P.msgs[mi].send_cnt; ++k) {
      const int v = P.send_lnode[(size_t)(P.msgs[mi].send_off + k)];
      const int e = first[(size_t)v] + fill[(size_t)v]++;
      P.push_ent_msg[(size_t)e] = (int)mi;
      P.push_ent_pos[(size_t)e] = (int)k;
    }
  for (int v = 0; v < n; ++v)
    if (cnt[(size_t)v] && P.push_valid) { P.push_node_ent[(size_t)v] = (first[(size_t)v] << 3) | cnt[(size_t)v]; P.push_grp[(size_t)(v >> 6)] = 1; }
}

HaloPlan plan_halo(const Layout &L, int rank) {
  HaloPlan P;
  L.box(rank, &P.xs, &P.ys, &P.xm, &P.ym);
  const int xs = P.xs, ys = P.ys, xm = P.xm, ym = P.ym;
  P.n_owned = xm * ym;
  // ghosts: the ring of width 1 around the owned box, clipped to the domain (box stencil -> corners included)
  struct G { int g, owner, i, j; };
  std::vector<G> gh;
  for (int j = std::max(ys - 1, 0); j <= std::min(ys + ym, L.N - 1); ++j)
    for (int i = std::max(xs - 1, 0); i <= std::min(xs + xm, L.M - 1); ++i)
      if (i < xs || i >= xs + xm || j < ys || j >= ys + ym) gh.push_back({L.gnode(i, j), L.owner(i, j), i, j});
  std::sort(gh.begin(), gh.end(), [](const G &a, const G &b) { return a.g < b.g; });
  for (const G &g : gh) { P.ghost_gnode.push_back(g.g); P.ghost_owner.push_back(g.owner); P.ghost_i.push_back(g.i); P.ghost_j.push_back(g.j); }
  // messages: neighbours in ascending rank.  Sends: the owned nodes inside q's ghost ring in OUR global order, which is
  // the order in which they appear in q's sorted ghost list restricted to owner == rank; receives: that range of ours.
  for (int q = 0; q < L.size; ++q) {
    if (q == rank) continue;
    int qxs, qys, qxm, qym;
    L.box(q, &qxs, &qys, &qxm, &qym);
    const int i0 = std::max(std::max(qxs - 1, 0), xs), i1 = std::min(std::min(qxs + qxm, L.M - 1), xs + xm - 1);
    const int j0 = std::max(std::max(qys - 1, 0), ys), j1 = std::min(std::min(qys + qym, L.N - 1), ys + ym - 1);
    HaloMsg msg{q, (int64_t)P.send_lnode.size(), 0, 0, 0};
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) { P.send_lnode.push_back((j - ys) * xm + (i - xs)); msg.send_cnt++; }
    int first = -1, cnt = 0;
    for (size_t t = 0; t < gh.size(); ++t)
      if (gh[t].owner == q) { if (first < 0) first = (int)t; cnt++; }
    msg.recv_off = first < 0 ? 0 : first;
    msg.recv_cnt = cnt;
    if (msg.send_cnt || msg.recv_cnt) P.msgs.push_back(msg);
  }
  build_push_tables(P);
  return P;
}

// device side of a halo from its plan; `ring` (DMDA halos only): the width-1 ring classification used by the assembly
static std::shared_ptr<Halo> halo_from_plan(Ctx *c, const Layout &L, int rank, const HaloPlan &P, bool ring_of_box) {
  auto h = std::make_shared<Halo>();
  h->ctx = c;
  h->M = L.M; h->N = L.N;
  h->xs = P.xs; h->ys = P.ys; h->xm = P.xm; h->ym = P.ym;
  const int xs = h->xs, ys = h->ys, xm = h->xm, ym = h->ym;
  h->n_owned = P.n_owned;
  h->n_ghost = (int)P.ghost_gnode.size();
  h->ghost_gnode = P.ghost_gnode; h->ghost_i = P.ghost_i; h->ghost_j = P.ghost_j;
  std::vector<int> ring((size_t)(2 * (xm + 2) + 2 * ym), -1);
  if (ring_of_box) {
    ColSpace cs{xs, ys, xm, ym, nullptr};
    for (int t = 0; t < h->n_ghost; ++t) ring[(size_t)cs.ring_id(P.ghost_i[(size_t)t], P.ghost_j[(size_t)t])] = t;
  }
  h->node_msgs = P.msgs;
  const std::vector<int> &send_lnode = P.send_lnode;
  h->h_push_grp = P.push_grp; h->h_push_node_ent = P.push_node_ent; h->h_push_ent_msg = P.push_ent_msg; h->h_push_ent_pos = P.push_ent_pos;
  h->n_send = (int)send_lnode.size();
  h->d_send_lnode.alloc((size_t)h->n_send + 1);
  h->d_ring2ghost.alloc(ring.size() + 1);
  if (h->n_send) B2_CUDA(cudaMemcpyAsync(h->d_send_lnode.p, send_lnode.data(), sizeof(int) * send_lnode.size(), cudaMemcpyHostToDevice, c->stream));
  B2_CUDA(cudaMemcpyAsync(h->d_ring2ghost.p, ring.data(), sizeof(int) * ring.size(), cudaMemcpyHostToDevice, c->stream));
  h->sendbuf.alloc((size_t)h->n_send * HALO_MAX_DOF + 2);
  h->ghost.alloc((size_t)h->n_ghost * HALO_MAX_DOF + 2);
  h->ghost.zero(c->stream);
  B2_CUDA(cudaEventCreateWithFlags(&h->ev_packed, cudaEventDisableTiming));
  B2_CUDA(cudaEventCreateWithFlags(&h->ev_arrived, cudaEventDisableTiming));
  c->sync();
  // every message must carry data in both directions for the flag protocol (true for a box-stencil ring)
  bool symmetric = (int)h->node_msgs.size() <= 8;
  for (const HaloMsg &m : h->node_msgs) symmetric = symmetric && m.send_cnt > 0 && m.recv_cnt > 0;
  symmetric = symmetric && P.push_valid;
  if (c->dcomm && c->dcomm->p2p_capable() && L.size > 1) {
    // the peer-to-peer setup is collective: every rank must take the same decision
    DevBuf<double> d_in(2), d_all((size_t)L.size + 1);
    const double mine = symmetric ? 1.0 : 0.0;
    std::vector<double> all((size_t)L.size);
    B2_CUDA(cudaMemcpyAsync(d_in.p, &mine, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->dcomm->allgather(d_in.p, d_all.p, 1, c->stream);
    B2_CUDA(cudaMemcpyAsync(all.data(), d_all.p, sizeof(double) * (size_t)L.size, cudaMemcpyDeviceToHost, c->stream));
    c->sync();
    for (double v : all) symmetric = symmetric && v != 0.0;
    if (symmetric) setup_p2p(*h, L, rank);
  }
  return h;
}

std::shared_ptr<Halo> make_halo(Ctx *c, const Layout &L, int rank) {
  HaloPlan P = plan_halo(L, rank);
  B2_REQUIRE(L.size == 1 || (P.xm >= 2 && P.ym >= 2), "dmda: every rank must own at least 2 x 2 nodes");
  return halo_from_plan(c, L, rank, P, true);
}

// Halo for an ARBITRARY ghost set (sorted, unique global node ids in PETSc numbering, none owned by this rank): what
// MatSetUpMultiply_MPIAIJ builds from the off-diagonal columns of an assembled matrix.  Used for matrices produced by
// the distributed SpGEMM (A10 A01 has a two-node-wide stencil).  Collective: every rank publishes its ghost list through
// the communicator's all-gather and finds the nodes it has to send in the requesters' order.
std::shared_ptr<Halo> make_halo_general(Ctx *c, const Layout &L, int rank, const std::vector<int> &ghost_gnode) {
  B2_REQUIRE(c->dcomm, "make_halo_general: context has no communicator");
  const int size = L.size;
  DevBuf<double> d_in(2), d_cnt((size_t)size + 1);
  const double mycnt = (double)ghost_gnode.size();
  B2_CUDA(cudaMemcpyAsync(d_in.p, &mycnt, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->dcomm->allgather(d_in.p, d_cnt.p, 1, c->stream);
  std::vector<double> h_cnt((size_t)size);
  B2_CUDA(cudaMemcpyAsync(h_cnt.data(), d_cnt.p, sizeof(double) * (size_t)size, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  int64_t maxc = 1;
  for (double v : h_cnt) maxc = std::max<int64_t>(maxc, (int64_t)v);
  std::vector<double> mine((size_t)maxc, -1.0), all((size_t)maxc * size);
  for (size_t t = 0; t < ghost_gnode.size(); ++t) mine[t] = (double)ghost_gnode[t];
  DevBuf<double> d_mine((size_t)maxc), d_all((size_t)maxc * size);
  B2_CUDA(cudaMemcpyAsync(d_mine.p, mine.data(), sizeof(double) * (size_t)maxc, cudaMemcpyHostToDevice, c->stream));
  c->dcomm->allgather(d_mine.p, d_all.p, maxc, c->stream);
  B2_CUDA(cudaMemcpyAsync(all.data(), d_all.p, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  HaloPlan P;
  P.n_owned = L.rstart[(size_t)rank + 1] - L.rstart[(size_t)rank];
  const int g0 = L.rstart[(size_t)rank], g1 = L.rstart[(size_t)rank + 1];
  P.ghost_gnode = ghost_gnode;
  for (int g : ghost_gnode) {
    B2_REQUIRE(g >= 0 && g < L.rstart[(size_t)size] && (g < g0 || g >= g1), "make_halo_general: ghost id owned by this rank or out of range");
    const int owner = (int)(std::upper_bound(L.rstart.begin(), L.rstart.end(), g) - L.rstart.begin()) - 1;
    P.ghost_owner.push_back(owner);
  }
  for (int q = 0; q < size; ++q) {
    if (q == rank) continue;
    HaloMsg msg{q, (int64_t)P.send_lnode.size(), 0, 0, 0};
    const int qn = (int)h_cnt[(size_t)q];
    for (int t = 0; t < qn; ++t) { // q's ghost list is sorted: the nodes I own appear in ascending order = q's receive order
      const int g = (int)all[(size_t)q * maxc + t];
      if (g >= g0 && g < g1) { P.send_lnode.push_back(g - g0); msg.send_cnt++; }
    }
    int first = -1, cnt = 0;
    for (size_t t = 0; t < P.ghost_owner.size(); ++t)
      if (P.ghost_owner[t] == q) { if (first < 0) first = (int)t; cnt++; }
    msg.recv_off = first < 0 ? 0 : first;
    msg.recv_cnt = cnt;
    if (msg.send_cnt || msg.recv_cnt) P.msgs.push_back(msg);
  }
  build_push_tables(P);
  return halo_from_plan(c, L, rank, P, false);
}

void Halo::begin(const double *x, int dof) {
  B2_REQUIRE(dof >= 1 && dof <= HALO_MAX_DOF, "halo: dof must be 1, 2 or 3");
  Ctx *c = ctx;
  if (!c->dcomm || (n_send == 0 && n_ghost == 0)) return;
  if (p2p && pushed_vec) { // the kernel that produced x already pushed it (csr_spmv_epi push_to)
    B2_REQUIRE(pushed_vec == x, "halo: the vector pushed ahead is not the one being multiplied");
    pushed_vec = nullptr;
    return;
  }
  if (p2p) { // push over NVLink + flag; the matching wait is in end()
    LaunchScope ls(c, "halo:p2p_push");
    int grid = (n_send * dof + 1023) / 1024; // a few elements per thread: fewer blocks, fewer fences and tickets
    if (grid > 16) grid = 16;
    if (grid < 1) grid = 1;
    k_halo_push<<<grid, 256, 0, c->stream>>>(n_send, dof, n_msgs, d_send_lnode.p, x, d_p2p.p, seq.p, ticket.p);
    check_launch("k_halo_push");
    return;
  }
  if (n_send) {
    LaunchScope ls(c, "halo");
    int grid = (n_send * dof + 255) / 256;
    k_pack<<<grid, 256, 0, c->stream>>>(n_send, dof, d_send_lnode.p, x, sendbuf.p);
    check_launch("k_pack");
  }
  B2_CUDA(cudaEventRecord(ev_packed, c->stream));
  B2_CUDA(cudaStreamWaitEvent(c->stream2, ev_packed, 0));
  std::vector<HaloMsg> msgs = node_msgs;
  for (HaloMsg &m : msgs) { m.send_off *= dof; m.send_cnt *= dof; m.recv_off *= dof; m.recv_cnt *= dof; }
  if (c->profile) { // measurement pass: time the exchange itself on the halo stream
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, c->stream2);
    c->dcomm->exchange(sendbuf.p, ghost.p, msgs, c->stream2);
    cudaEventRecord(e1, c->stream2);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    auto &pe = c->prof["comm:halo_exchange"];
    pe.ms += ms; pe.n++;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  } else {
    c->dcomm->exchange(sendbuf.p, ghost.p, msgs, c->stream2);
  }
  B2_CUDA(cudaEventRecord(ev_arrived, c->stream2));
}
void Halo::end() {
  Ctx *c = ctx;
  if (!c->dcomm || (n_send == 0 && n_ghost == 0)) return;
  if (p2p) {
    LaunchScope ls(c, "halo:p2p_wait");
    k_halo_wait<<<1, 32, 0, c->stream>>>(n_msgs, seq.p, flags.p, c->d_err);
    check_launch("k_halo_wait");
    return;
  }
  B2_CUDA(cudaStreamWaitEvent(c->stream, ev_arrived, 0));
}

} // namespace b200sp
