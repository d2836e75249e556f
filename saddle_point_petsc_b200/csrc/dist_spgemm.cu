// dist_spgemm.cu -- C = A * B for row-partitioned (MPIAIJ-like) operands: MatMatMult_MPIAIJ_MPIAIJ.
//
// PETSc forms the Schur preconditioning matrices of PCFIELDSPLIT (selfp: A11 - A10 diag(A00)^-1 A01) and PCLSC
// (L = A10 [diag(A00)^-1] A01) with MatMatMult on the distributed blocks.  Row i of the product needs the rows of B
// that A's row i references, including the rows of the GHOST columns of A, which live on the neighbours:
//   1. every rank sends the rows of B that belong to the nodes on its halo send list (lengths first, then the
//      (global column, value) pairs) through the communicator's neighbour exchange -- the same messages as one halo
//      exchange of A's column space, setup only;
//   2. the local rows of B and the received ghost rows are stacked in A's local column numbering (owned rows, then
//      ghost rows in ghost order) with their columns renumbered into a working column space: owned columns first,
//      then the sorted set of every other global column that occurs (a superset of B's own ghost columns: the product
//      stencil is wider than either factor's);
//   3. the local SpGEMM kernels run on that stack (kernels_setup.cu, spgemm_raw);
//   4. the result gets a halo of its own, built from the extended ghost set (make_halo_general).
// The products of one entry are added in the order in which k appears in A's LOCAL row (owned columns, then ghosts),
// so the values agree with the single-rank product to rounding, not bit for bit (PETSc has the same property).
#include <algorithm>
#include <cstring>
#include "dist.h"

namespace b200sp {

std::shared_ptr<Csr> csr_matmat_dist(const Csr &A, const Csr &B) {
  Ctx *c = A.ctx;
  B2_REQUIRE(A.halo && B.halo && A.layout && B.layout && c->dcomm, "csr_matmat (row-partitioned): both operands must be distributed DMDA matrices");
  B2_REQUIRE(A.ncols == B.nrows && A.halo_dof == B.dof_r && B.dof_r >= 1, "csr_matmat (row-partitioned): inner dimensions / node dof differ");
  const Layout &L = *A.layout;
  const int rank = c->rank, dr = A.halo_dof, dc = B.halo_dof;
  Halo &hA = *A.halo;
  // ---- B on the host, columns as GLOBAL scalar ids (PETSc numbering)
  std::vector<int> rp((size_t)B.nrows + 1), cl((size_t)B.nnz + 1);
  std::vector<double> va((size_t)B.nnz + 1);
  B2_CUDA(cudaMemcpyAsync(rp.data(), B.rowptr.p, sizeof(int) * ((size_t)B.nrows + 1), cudaMemcpyDeviceToHost, c->stream));
  if (B.nnz) {
    B2_CUDA(cudaMemcpyAsync(cl.data(), B.col.p, sizeof(int) * (size_t)B.nnz, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaMemcpyAsync(va.data(), B.val.p, sizeof(double) * (size_t)B.nnz, cudaMemcpyDeviceToHost, c->stream));
  }
  std::vector<int> send_lnode((size_t)hA.n_send + 1);
  if (hA.n_send) B2_CUDA(cudaMemcpyAsync(send_lnode.data(), hA.d_send_lnode.p, sizeof(int) * (size_t)hA.n_send, cudaMemcpyDeviceToHost, c->stream));
  c->sync();
  const int64_t cg0 = B.col_gstart, cg1 = B.col_gstart + B.ncols;
  auto gcol = [&](int k) -> int64_t {
    const int q = cl[(size_t)k];
    if (q < B.ncols) return cg0 + q;
    return (int64_t)B.halo->ghost_gnode[(size_t)((q - B.ncols) / dc)] * dc + (q - B.ncols) % dc;
  };
  // ---- 1a. row lengths of the rows on the send list -> lengths of my ghost rows
  const int n_send = hA.n_send, n_ghost = hA.n_ghost;
  std::vector<double> s_len((size_t)n_send * dr + 1, 0.0), g_len((size_t)n_ghost * dr + 1, 0.0);
  for (int t = 0; t < n_send; ++t)
    for (int d = 0; d < dr; ++d) {
      const int r = send_lnode[(size_t)t] * dr + d;
      s_len[(size_t)t * dr + d] = (double)(rp[(size_t)r + 1] - rp[(size_t)r]);
    }
  std::vector<HaloMsg> m1 = hA.node_msgs;
  for (HaloMsg &m : m1) { m.send_off *= dr; m.send_cnt *= dr; m.recv_off *= dr; m.recv_cnt *= dr; }
  {
    DevBuf<double> d_s(s_len.size()), d_g(g_len.size());
    B2_CUDA(cudaMemcpyAsync(d_s.p, s_len.data(), sizeof(double) * s_len.size(), cudaMemcpyHostToDevice, c->stream));
    c->dcomm->exchange(d_s.p, d_g.p, m1, c->stream);
    B2_CUDA(cudaMemcpyAsync(g_len.data(), d_g.p, sizeof(double) * g_len.size(), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
  }
  // ---- 1b. the rows themselves: (global column, value) pairs, message by message
  std::vector<int> grp((size_t)n_ghost * dr + 1, 0);
  for (int t = 0; t < n_ghost * dr; ++t) grp[(size_t)t + 1] = grp[(size_t)t] + (int)g_len[(size_t)t];
  std::vector<int64_t> s_off((size_t)n_send * dr + 1, 0);
  for (int t = 0; t < n_send * dr; ++t) s_off[(size_t)t + 1] = s_off[(size_t)t] + (int64_t)s_len[(size_t)t];
  const int64_t s_tot = s_off[(size_t)n_send * dr], g_tot = grp[(size_t)n_ghost * dr];
  std::vector<double> s_pay((size_t)2 * s_tot + 2), g_pay((size_t)2 * g_tot + 2);
  for (int t = 0; t < n_send; ++t)
    for (int d = 0; d < dr; ++d) {
      const int r = send_lnode[(size_t)t] * dr + d;
      int64_t o = 2 * s_off[(size_t)t * dr + d];
      for (int k = rp[(size_t)r]; k < rp[(size_t)r + 1]; ++k) { s_pay[(size_t)o++] = (double)gcol(k); s_pay[(size_t)o++] = va[(size_t)k]; }
    }
  std::vector<HaloMsg> m2 = m1;
  for (size_t i = 0; i < m2.size(); ++i) {
    const HaloMsg &a = m1[i];
    m2[i].send_off = 2 * s_off[(size_t)a.send_off];
    m2[i].send_cnt = 2 * (s_off[(size_t)(a.send_off + a.send_cnt)] - s_off[(size_t)a.send_off]);
    m2[i].recv_off = 2 * (int64_t)grp[(size_t)a.recv_off];
    m2[i].recv_cnt = 2 * (int64_t)(grp[(size_t)(a.recv_off + a.recv_cnt)] - grp[(size_t)a.recv_off]);
  }
  {
    DevBuf<double> d_s(s_pay.size()), d_g(g_pay.size());
    B2_CUDA(cudaMemcpyAsync(d_s.p, s_pay.data(), sizeof(double) * s_pay.size(), cudaMemcpyHostToDevice, c->stream));
    c->dcomm->exchange(d_s.p, d_g.p, m2, c->stream);
    B2_CUDA(cudaMemcpyAsync(g_pay.data(), d_g.p, sizeof(double) * g_pay.size(), cudaMemcpyDeviceToHost, c->stream));
    c->sync();
  }
  // ---- 2. extended ghost set of the column space, working column ids, the stacked B
  std::vector<int> ext;
  for (int64_t k = 0; k < B.nnz; ++k)
    if (cl[(size_t)k] >= B.ncols) ext.push_back((int)(gcol((int)k) / dc));
  for (int64_t t = 0; t < g_tot; ++t) {
    const int64_t g = (int64_t)g_pay[(size_t)(2 * t)];
    if (g < cg0 || g >= cg1) ext.push_back((int)(g / dc));
  }
  std::sort(ext.begin(), ext.end());
  ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
  auto work = [&](int64_t g) -> int {
    if (g >= cg0 && g < cg1) return (int)(g - cg0);
    const int node = (int)(g / dc);
    const int idx = (int)(std::lower_bound(ext.begin(), ext.end(), node) - ext.begin());
    return B.ncols + idx * dc + (int)(g % dc);
  };
  const int nrows_ext = B.nrows + n_ghost * dr;
  const int64_t nnz_ext = B.nnz + g_tot;
  B2_REQUIRE(nnz_ext < (int64_t)2147483647, "csr_matmat (row-partitioned): stacked operand too large");
  std::vector<int> e_rp((size_t)nrows_ext + 1), e_col((size_t)nnz_ext + 1);
  std::vector<double> e_val((size_t)nnz_ext + 1);
  for (int r = 0; r <= B.nrows; ++r) e_rp[(size_t)r] = rp[(size_t)r];
  for (int64_t k = 0; k < B.nnz; ++k) { e_col[(size_t)k] = work(gcol((int)k)); e_val[(size_t)k] = va[(size_t)k]; }
  for (int t = 0; t < n_ghost * dr; ++t) e_rp[(size_t)B.nrows + t + 1] = (int)(B.nnz + grp[(size_t)t + 1]);
  for (int64_t t = 0; t < g_tot; ++t) { e_col[(size_t)(B.nnz + t)] = work((int64_t)g_pay[(size_t)(2 * t)]); e_val[(size_t)(B.nnz + t)] = g_pay[(size_t)(2 * t + 1)]; }
  DevBuf<int> d_rp(e_rp.size()), d_col(e_col.size());
  DevBuf<double> d_val(e_val.size());
  B2_CUDA(cudaMemcpyAsync(d_rp.p, e_rp.data(), sizeof(int) * e_rp.size(), cudaMemcpyHostToDevice, c->stream));
  B2_CUDA(cudaMemcpyAsync(d_col.p, e_col.data(), sizeof(int) * e_col.size(), cudaMemcpyHostToDevice, c->stream));
  B2_CUDA(cudaMemcpyAsync(d_val.p, e_val.data(), sizeof(double) * e_val.size(), cudaMemcpyHostToDevice, c->stream));
  // ---- 3. local SpGEMM on the stack
  auto C = spgemm_raw(c, A.nrows, A.rowptr.p, A.col.p, A.val.p, d_rp.p, d_col.p, d_val.p, B.ncols + (int)ext.size() * dc);
  c->sync();
  // ---- 4. the product's own column space: owned columns of B + the extended ghost set
  C->ncols = B.ncols;
  C->halo = make_halo_general(c, L, rank, ext);
  C->halo_dof = dc;
  C->layout = A.layout;
  C->row_gstart = A.row_gstart;
  C->col_gstart = B.col_gstart;
  C->grid_M = A.grid_M; C->grid_N = A.grid_N; C->dof_r = A.dof_r; C->dof_c = dc;
  C->plan();
  return C;
}

} // namespace b200sp
