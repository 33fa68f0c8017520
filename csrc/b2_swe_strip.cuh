// mpi4jax_b200 -- shallow water, the BULK kernel of the communication-avoiding step: one whole model
// step (fluxes -> tendencies -> friction-u -> friction-v) in ONE pass over memory.
//
// The reference's step makes XLA stream ~37 full arrays through HBM, the stand-alone kernels of
// b2_swe.cu 32.  A step only HAS to read h, u, v, dh, du, dv and write them back: 12 passes.  This
// kernel does exactly that.  A CTA owns a strip of TXW = NT - 7 columns and marches down RY rows;
// one thread per column.  Everything between the six input and six output streams lives in small
// shared-memory rings that roll along with the row index r:
//
//   L  load   h, u, v            row r      -> H, U, V      (4-row rings)
//   F  fluxes fe, fn, q, ke      row r - 1  -> FE, FN, Q, KE
//   T  tendencies + AB2 update   row r - 2  -> h', dh, du, dv to memory;  u', v' -> UP, VP
//   U  friction-u                row r - 3  -> UPP (u'')
//   V  friction-v                row r - 3  -> u'', v'' to memory
//
// Each quantity is computed ONCE per cell (the register-only fused kernels it replaces evaluated
// 9.5 flux quantities and 2.25 friction stencils per cell and were issue-bound); the redundancy is
// the halo of the strip: 7 of NT columns, 7 rows per RY.  The dependency cone of an output cell is
// four cells wide and reaches only the step's INPUT arrays, so the kernel needs nothing from the
// step's halo exchange: it runs concurrently with the frame pipeline of b2_swe_ca.cu, which owns
// the cells within three of the block edge.
//
// Outputs go to the ping-pong partners of the inputs (a neighbouring CTA still reads the old value
// of a cell this CTA has finished).  The arithmetic is the shared explicit-rounding helpers of
// b2_swe_body.cuh on the same operands as the stand-alone kernels: same bits.
#pragma once

#include "b2_swe_body.cuh"

#define STRIP_NT 256                 // threads per CTA = columns incl. halo (4 west, 3 east)
#define STRIP_TXW (STRIP_NT - 7)     // output columns per CTA
#define STRIP_RY 128                 // output rows per CTA

struct StripArgs {
  B2SweParams p;
  int cb1;                           // bulk columns [4, cb1), bulk rows [4, ny - 4)
  const float *h, *u, *v, *dh, *du, *dv;             // step inputs
  float *h_o, *u_o, *v_o, *dh_o, *du_o, *dv_o;       // step outputs (ping-pong partners)
};

// the rings of one CTA
struct StripSmem {
  float H[4][STRIP_NT], U[4][STRIP_NT], V[4][STRIP_NT];
  float FE[4][STRIP_NT], FN[4][STRIP_NT], Q[4][STRIP_NT], KE[4][STRIP_NT];
  float UP[4][STRIP_NT], VP[4][STRIP_NT];
  float UPP[2][STRIP_NT];
};

__host__ __device__ inline int strip_nstrips(const B2SweParams& p, int cb1) { return (cb1 - 4 + STRIP_TXW - 1) / STRIP_TXW; }
__host__ __device__ inline int strip_nchunks(const B2SweParams& p) { return (p.ny - 8 + STRIP_RY - 1) / STRIP_RY; }

// geometry of CTA `b`: first output column / row range, and this thread's column
struct StripGeo {
  int i0, j0, j1;
};
__host__ __device__ inline StripGeo strip_geo(const B2SweParams& p, int cb1, int b) {
  const int ns = strip_nstrips(p, cb1);
  StripGeo g;
  g.i0 = 4 + (b % ns) * STRIP_TXW;
  g.j0 = 4 + (b / ns) * STRIP_RY;
  g.j1 = g.j0 + STRIP_RY < p.ny - 4 ? g.j0 + STRIP_RY : p.ny - 4;
  return g;
}
__host__ __device__ __forceinline__ int strip_col(const B2SweParams& p, const StripGeo& g, int tid) {
  const int i = g.i0 - 4 + tid;                     // columns beyond the block are clamped: their values
  return i > p.nx - 1 ? p.nx - 1 : i;               // feed only cells that are not written
}

// what a thread knows about itself and its CTA, computed once (the row loop then spends its integer
// instructions on one offset increment and a handful of row compares)
struct StripThr {
  int tid;
  bool f_ok, t_ok, u_ok, o_ok;       // this column takes part in the flux / tendency / friction-u phase; is an output column
  int r_f, r_t, r_u;                 // first row index r at which those phases have valid operands
  int r_t0, r_t1, r_o0, r_o1;        // row indices whose tendency / friction results are written to memory
  int r_last;                        // last input row
  size_t col, pitch;
};
__host__ __device__ inline StripThr strip_thread(const StripArgs& a, const StripGeo& g, int tid) {
  StripThr t;
  t.tid = tid;
  const int i = g.i0 - 4 + tid;
  t.f_ok = tid >= 1 && tid <= STRIP_NT - 2;
  t.t_ok = tid >= 2 && tid <= STRIP_NT - 3;
  t.u_ok = tid >= 3 && tid <= STRIP_NT - 4;
  t.o_ok = tid >= 4 && tid <= STRIP_NT - 4 && i < a.cb1;
  t.r_f = (g.j0 - 3 > 1 ? g.j0 - 3 : 1) + 1;          // jf = r - 1 >= max(1, j0 - 3)
  t.r_t = (g.j0 - 2 > 2 ? g.j0 - 2 : 2) + 2;          // jt = r - 2 >= max(2, j0 - 2)
  t.r_u = (g.j0 - 1 > 3 ? g.j0 - 1 : 3) + 3;          // jr = r - 3 >= max(3, j0 - 1)
  t.r_t0 = g.j0 + 2; t.r_t1 = g.j1 + 2;
  t.r_o0 = g.j0 + 3; t.r_o1 = g.j1 + 3;
  t.r_last = a.p.ny - 1;
  t.col = (size_t)strip_col(a.p, g, tid);
  t.pitch = (size_t)a.p.pitch;
  return t;
}

// ---- the five phases of row index r (a sync separates consecutive phases) ------------------------
// Q = r & 3 is a template parameter: the row loop is unrolled by four, so every ring slot is a
// compile-time constant and a shared-memory access costs one instruction (the first version spent
// more instructions on ring-index arithmetic than on floating point).  `off` = r * pitch + column.
template <int Q>
__device__ __forceinline__ void strip_load(const StripArgs& a, StripSmem& s, const StripThr& t, int r, size_t off) {
  if (r > t.r_last) return;
  const int tid = t.tid;
  s.H[Q][tid] = a.h[off];
  s.U[Q][tid] = a.u[off];
  s.V[Q][tid] = a.v[off];
}
// fluxes of row jf = r - 1 (rows 1 .. ny-3: no wall rule applies)
template <int Q>
__device__ __forceinline__ void strip_flux(const StripArgs& a, StripSmem& s, const StripThr& t, int r) {
  if (!t.f_ok || r < t.r_f) return;
  const int tid = t.tid;
  constexpr int c = (Q + 3) & 3, n = Q, m = (Q + 2) & 3;
  const float h_c = s.H[c][tid], h_e = s.H[c][tid + 1], h_n = s.H[n][tid], h_ne = s.H[n][tid + 1];
  const float u_c = s.U[c][tid], v_c = s.V[c][tid];
  s.FE[c][tid] = swe_fe(h_c, h_e, u_c);
  s.FN[c][tid] = swe_fn(h_c, h_n, v_c);
  s.Q[c][tid] = swe_q(a.p, a.p.coriolis[r - 1], s.V[c][tid + 1], v_c, s.U[n][tid], u_c, h_c, h_e, h_n, h_ne);
  s.KE[c][tid] = swe_ke(u_c, s.U[c][tid - 1], v_c, s.V[m][tid]);
}
// tendencies + update of row jt = r - 2
template <int Q>
__device__ __forceinline__ void strip_tend(const StripArgs& a, StripSmem& s, const StripThr& t, int r, size_t off_r) {
  if (!t.t_ok || r < t.r_t) return;
  const int tid = t.tid;
  const B2SweParams& p = a.p;
  constexpr int c = (Q + 2) & 3, n = (Q + 3) & 3, m = (Q + 1) & 3;
  const size_t off = off_r - 2 * t.pitch;
  SweK2In in;
  in.fe_c = s.FE[c][tid]; in.fe_w = s.FE[c][tid - 1]; in.fen_c = s.FE[n][tid]; in.fen_w = s.FE[n][tid - 1];
  in.fn_c = s.FN[c][tid]; in.fn_e = s.FN[c][tid + 1]; in.fns_c = s.FN[m][tid]; in.fns_e = s.FN[m][tid + 1];
  in.q_c = s.Q[c][tid]; in.q_w = s.Q[c][tid - 1]; in.qs_c = s.Q[m][tid];
  in.ke_c = s.KE[c][tid]; in.ke_e = s.KE[c][tid + 1]; in.ken_c = s.KE[n][tid];
  in.h_c = s.H[c][tid]; in.h_e = s.H[c][tid + 1]; in.h_n = s.H[n][tid];
  in.u_o = s.U[c][tid]; in.v_o = s.V[c][tid];
  in.dh_o = in.du_o = in.dv_o = 0.f;
  if (!p.first_step) { in.dh_o = a.dh[off]; in.du_o = a.du[off]; in.dv_o = a.dv[off]; }
  const SweK2Out o = swe_k2_cell(p, in);
  s.UP[c][tid] = o.u;
  s.VP[c][tid] = o.v;
  if (t.o_ok && r >= t.r_t0 && r < t.r_t1) {
    a.h_o[off] = o.h; a.dh_o[off] = o.dh; a.du_o[off] = o.du; a.dv_o[off] = o.dv;
  }
}
// friction-u of row jr = r - 3
template <int Q>
__device__ __forceinline__ void strip_fric_u(const StripArgs& a, StripSmem& s, const StripThr& t, int r) {
  if (!t.u_ok || r < t.r_u) return;
  const int tid = t.tid;
  constexpr int c = (Q + 1) & 3, n = (Q + 2) & 3, m = Q;          // rows jr, jr + 1, jr - 1 (= r - 4)
  s.UPP[(Q + 1) & 1][tid] = swe_friction_u(a.p, s.UP[c][tid], s.UP[c][tid + 1], s.UP[c][tid - 1], s.UP[n][tid],
                                           s.UP[m][tid], false, false);
}
// friction-v of row jr = r - 3, and the step's u'', v'' to memory
template <int Q>
__device__ __forceinline__ void strip_fric_v(const StripArgs& a, StripSmem& s, const StripThr& t, int r, size_t off_r) {
  if (!t.o_ok || r < t.r_o0 || r >= t.r_o1) return;
  const int tid = t.tid;
  const B2SweParams& p = a.p;
  constexpr int c = (Q + 1) & 3, n = (Q + 2) & 3, k = (Q + 1) & 1;
  const float upp = s.UPP[k][tid], v_c = s.VP[c][tid];
  const float fe2_c = swe_visc_flux(p.c_nux, s.VP[c][tid + 1], upp);
  const float fe2_w = swe_visc_flux(p.c_nux, v_c, s.UPP[k][tid - 1]);
  const float fn2_c = swe_visc_flux(p.c_nuy, s.VP[n][tid], upp);
  const float fn2_s = swe_visc_flux(p.c_nuy, v_c, s.UPP[k ^ 1][tid]);
  const size_t off = off_r - 3 * t.pitch;
  a.u_o[off] = upp;
  a.v_o[off] = swe_apply_div(p, v_c, fe2_c, fe2_w, fn2_c, fn2_s);
}

// One row index (Q = r & 3, r >= 0).  `each(phase)` runs `phase(thread)` for the CTA's threads and then
// synchronises them: on the device it is the calling thread + __syncthreads, the host emulation
// loops over all threads.  `thr(tid)` returns the thread's StripThr.
template <int Q, class Each>
__device__ __forceinline__ void strip_row(const StripArgs& a, StripSmem& s, int r, Each&& each) {
  each([&](const StripThr& t) { strip_load<Q>(a, s, t, r, (size_t)r * t.pitch + t.col); });   // (slot of row r - 4: no longer read)
  each([&](const StripThr& t) { strip_flux<Q>(a, s, t, r); });
  each([&](const StripThr& t) { strip_tend<Q>(a, s, t, r, (size_t)r * t.pitch + t.col); });
  each([&](const StripThr& t) { strip_fric_u<Q>(a, s, t, r); });
  each([&](const StripThr& t) { strip_fric_v<Q>(a, s, t, r, (size_t)r * t.pitch + t.col); });
}
// all rows of one CTA: r runs from the multiple of four at or below j0 - 4 to j1 + 2
template <class Each>
__device__ __forceinline__ void strip_cta(const StripArgs& a, StripSmem& s, const StripGeo& g, Each&& each) {
  for (int rb = (g.j0 - 4) & ~3; rb <= g.j1 + 2; rb += 4) {
    strip_row<0>(a, s, rb, each);
    strip_row<1>(a, s, rb + 1, each);
    strip_row<2>(a, s, rb + 2, each);
    strip_row<3>(a, s, rb + 3, each);
  }
}
