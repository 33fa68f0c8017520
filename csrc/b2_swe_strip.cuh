// mpi4jax_b200 -- shallow water, the BULK kernel of the communication-avoiding step: one whole model
// step (fluxes -> tendencies -> friction-u -> friction-v) in ONE pass over memory.
//
// The reference's step makes XLA stream ~37 full arrays through HBM, the stand-alone kernels of
// b2_swe.cu 32.  A step only HAS to read h, u, v, dh, du, dv and write them back: 12 passes.  This
// kernel does exactly that.  A CTA owns a strip of TXW = NT - 7 columns and marches down RY rows;
// one thread per column.  Everything between the six input and six output streams lives in small
// shared-memory rings that roll along with the row index r.  The five phases of an iteration work on
// FIVE DIFFERENT rows, each consuming only what the previous iterations produced, so one barrier
// per row is enough and the phases' instructions interleave freely:
//
//   L  input rows              row r + 3  cp.async global -> H, U, V rings (three rows in flight per array;
//                                         old tendencies of row r - 1 likewise, used three iterations on)
//   F  fluxes fe, fn, q, ke    row r - 2  -> FE, FN, Q, KE
//   T  tendencies + AB2 update row r - 4  -> h', dh, du, dv to memory;  u', v' -> UP, VP
//   U  friction-u              row r - 6  -> UPP (u'')
//   V  friction-v              row r - 7  -> u'', v'' to memory
//
// (The first version had the phases on consecutive rows with a barrier and a global load between
// them: 1.8 us per row; the second prefetched one row ahead into registers: still waiting on
// memory every iteration.  A CTA's row loop is serial, so its loads must be issued microseconds,
// not one iteration, before they are needed.)
//
// Each quantity is computed ONCE per cell (the register-only fused kernels it replaces evaluated
// 9.5 flux quantities and 2.25 friction stencils per cell and were issue-bound); the redundancy is
// the halo of the strip: 7 of NT columns, 11 rows per RY.  The dependency cone of an output cell is
// four cells wide and reaches only the step's INPUT arrays, so the kernel needs nothing from the
// step's halo exchange: it runs concurrently with the frame pipeline of b2_swe_ca.cu, which owns
// the cells within three of the block edge.
//
// Outputs go to the ping-pong partners of the inputs (a neighbouring CTA still reads the old value
// of a cell this CTA has finished).  The arithmetic is the shared explicit-rounding helpers of
// b2_swe_body.cuh on the same operands as the stand-alone kernels: same bits.
#pragma once

#include "b2_swe_body.cuh"

// CTA shape: NT threads = NT columns incl. the halo (4 west, 3 east), RY output rows.  Both are chosen
// per block size on the host (strip_shape): wide strips / long chunks amortise the halo on a 4096^2
// block, a 2048 x 1024 block (8 GPUs) needs narrower, shorter CTAs to fill 148 SMs and to keep the
// serial row loop of a CTA -- the kernel's latency -- short.
struct StripArgs {
  B2SweParams p;
  int cb1;                           // bulk columns [4, cb1), bulk rows [4, ny - 4)
  int nt, ry;                        // threads per CTA (128 or 256), output rows per CTA
  const float *h, *u, *v, *dh, *du, *dv;             // step inputs
  float *h_o, *u_o, *v_o, *dh_o, *du_o, *dv_o;       // step outputs (ping-pong partners)
};

// the rings of one CTA
template <int NT>
struct StripSmem {
  float H[8][NT], U[8][NT], V[8][NT];
  float FE[4][NT], FN[4][NT], Q[4][NT], KE[4][NT];
  float UP[4][NT], VP[4][NT];
  float UPP[4][NT];
  float DH[4][NT], DU[4][NT], DV[4][NT];      // old tendencies, own column only (cp.async landing zone)
};

__host__ __device__ inline int strip_nstrips(const StripArgs& a) { return (a.cb1 - 4 + (a.nt - 7) - 1) / (a.nt - 7); }
__host__ __device__ inline int strip_nchunks(const StripArgs& a) { return (a.p.ny - 8 + a.ry - 1) / a.ry; }
// (nt, ry) for a block: the cheapest in thread-rows among the shapes that give every SM about three CTAs
inline void strip_shape(StripArgs& a, int sm_count) {
  double best = 1e300;
  const int nts[2] = {256, 128}, rys[5] = {128, 64, 32, 16, 8};
  for (int x = 0; x < 2; ++x)
    for (int y = 0; y < 5; ++y) {
      StripArgs t = a;
      t.nt = nts[x]; t.ry = rys[y];
      const double ctas = (double)strip_nstrips(t) * strip_nchunks(t);
      double cost = ctas * t.nt * (t.ry + 11);                                   // thread-rows executed
      if (ctas < 3.0 * sm_count) cost *= 3.0 * sm_count / ctas;                  // idle SMs / long serial loops
      if (cost < best) { best = cost; a.nt = t.nt; a.ry = t.ry; }
    }
}

// geometry of CTA `b`: first output column / row range, and this thread's column
struct StripGeo {
  int i0, j0, j1;
};
__host__ __device__ inline StripGeo strip_geo(const StripArgs& a, int b) {
  const int ns = strip_nstrips(a);
  StripGeo g;
  g.i0 = 4 + (b % ns) * (a.nt - 7);
  g.j0 = 4 + (b / ns) * a.ry;
  g.j1 = g.j0 + a.ry < a.p.ny - 4 ? g.j0 + a.ry : a.p.ny - 4;
  return g;
}
__host__ __device__ __forceinline__ int strip_col(const B2SweParams& p, const StripGeo& g, int tid) {
  const int i = g.i0 - 4 + tid;                     // columns beyond the block are clamped: their values
  return i > p.nx - 1 ? p.nx - 1 : i;               // feed only cells that are not written
}

// what a thread knows about itself and its CTA, computed once (the row loop then spends its integer
// instructions on a handful of row compares)
struct StripThr {
  int tid;
  bool f_ok, t_ok, u_ok, o_ok;       // this column takes part in the flux / tendency / friction-u phase; is an output column
  int r_f, r_t, r_u;                 // first row index r at which those phases have valid operands
  int r_t0, r_t1, r_o0, r_o1;        // row indices whose tendency / friction results are written to memory
  int r_last;                        // last input row
  size_t col, pitch;
  long long off;                     // r * pitch + column of the current row index (advanced once per row)
};
__host__ __device__ inline StripThr strip_thread(const StripArgs& a, const StripGeo& g, int tid) {
  StripThr t;
  t.tid = tid;
  const int i = g.i0 - 4 + tid;
  t.f_ok = tid >= 1 && tid <= a.nt - 2;
  t.t_ok = tid >= 2 && tid <= a.nt - 3;
  t.u_ok = tid >= 3 && tid <= a.nt - 4;
  t.o_ok = tid >= 4 && tid <= a.nt - 4 && i < a.cb1;
  t.r_f = (g.j0 - 3 > 1 ? g.j0 - 3 : 1) + 2;          // jf = r - 2 >= max(1, j0 - 3)
  t.r_t = (g.j0 - 2 > 2 ? g.j0 - 2 : 2) + 4;          // jt = r - 4 >= max(2, j0 - 2)
  t.r_u = (g.j0 - 1 > 3 ? g.j0 - 1 : 3) + 6;          // jr = r - 6 >= max(3, j0 - 1)
  t.r_t0 = g.j0 + 4; t.r_t1 = g.j1 + 4;
  t.r_o0 = g.j0 + 7; t.r_o1 = g.j1 + 7;
  t.r_last = a.p.ny - 1;
  t.col = (size_t)strip_col(a.p, g, tid);
  t.pitch = (size_t)a.p.pitch;
  t.off = 0;
  return t;
}

// ---- the phases of row index r; Q = r & 7 is a template parameter (the row loop is unrolled by eight),
// so every ring slot is a compile-time constant and a shared-memory access costs one instruction ----
// asynchronous 4-byte copy global -> shared (the host emulation copies at once)
#ifdef __CUDA_ARCH__
__device__ __forceinline__ void strip_cp(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void strip_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void strip_wait3() { asm volatile("cp.async.wait_group 3;" ::: "memory"); }
#else
static inline void strip_cp(float* dst, const float* src) { *dst = *src; }
static inline void strip_commit() {}
static inline void strip_wait3() {}
#endif
#define STRIP_PF 3      // rows in flight ahead of the row being assembled
// group of iteration r: inputs of row r + 3 and the old tendencies of row r - 1 (this thread's column);
// afterwards everything issued three iterations ago (row r, tendencies of row r - 4) has landed
template <int Q, class SM>
__device__ __forceinline__ void strip_issue(const StripArgs& a, SM& s, const StripThr& t, int r) {
  const int tid = t.tid;
  if (r + STRIP_PF >= 0 && r + STRIP_PF <= t.r_last) {
    const long long off = t.off + STRIP_PF * (long long)t.pitch;
    constexpr int w = (Q + STRIP_PF) & 7;
    strip_cp(&s.H[w][tid], a.h + off);
    strip_cp(&s.U[w][tid], a.u + off);
    strip_cp(&s.V[w][tid], a.v + off);
  }
  if (!a.p.first_step && t.t_ok && r + 3 >= t.r_t && r - 1 <= t.r_last) {
    const long long off = t.off - (long long)t.pitch;
    constexpr int w = (Q + 3) & 3;
    strip_cp(&s.DH[w][tid], a.dh + off);
    strip_cp(&s.DU[w][tid], a.du + off);
    strip_cp(&s.DV[w][tid], a.dv + off);
  }
  strip_commit();
  strip_wait3();
}
// fluxes of row jf = r - 2 (rows 1 .. ny-3: no wall rule applies)
template <int Q, class SM>
__device__ __forceinline__ void strip_flux(const StripArgs& a, SM& s, const StripThr& t, int r) {
  if (!t.f_ok || r < t.r_f) return;
  const int tid = t.tid;
  constexpr int c = (Q + 6) & 7, n = (Q + 7) & 7, m = (Q + 5) & 7, w = (Q + 2) & 3;
  const float h_c = s.H[c][tid], h_e = s.H[c][tid + 1], h_n = s.H[n][tid], h_ne = s.H[n][tid + 1];
  const float u_c = s.U[c][tid], v_c = s.V[c][tid];
  s.FE[w][tid] = swe_fe(h_c, h_e, u_c);
  s.FN[w][tid] = swe_fn(h_c, h_n, v_c);
  s.Q[w][tid] = swe_q(a.p, a.p.coriolis[r - 2], s.V[c][tid + 1], v_c, s.U[n][tid], u_c, h_c, h_e, h_n, h_ne);
  s.KE[w][tid] = swe_ke(u_c, s.U[c][tid - 1], v_c, s.V[m][tid]);
}
// tendencies + update of row jt = r - 4
template <int Q, class SM>
__device__ __forceinline__ void strip_tend(const StripArgs& a, SM& s, const StripThr& t, int r) {
  if (!t.t_ok || r < t.r_t) return;
  const int tid = t.tid;
  const B2SweParams& p = a.p;
  constexpr int c = Q & 3, n = (Q + 1) & 3, m = (Q + 3) & 3, hc = (Q + 4) & 7, hn = (Q + 5) & 7;
  SweK2In in;
  in.fe_c = s.FE[c][tid]; in.fe_w = s.FE[c][tid - 1]; in.fen_c = s.FE[n][tid]; in.fen_w = s.FE[n][tid - 1];
  in.fn_c = s.FN[c][tid]; in.fn_e = s.FN[c][tid + 1]; in.fns_c = s.FN[m][tid]; in.fns_e = s.FN[m][tid + 1];
  in.q_c = s.Q[c][tid]; in.q_w = s.Q[c][tid - 1]; in.qs_c = s.Q[m][tid];
  in.ke_c = s.KE[c][tid]; in.ke_e = s.KE[c][tid + 1]; in.ken_c = s.KE[n][tid];
  in.h_c = s.H[hc][tid]; in.h_e = s.H[hc][tid + 1]; in.h_n = s.H[hn][tid];
  in.u_o = s.U[hc][tid]; in.v_o = s.V[hc][tid];
  in.dh_o = in.du_o = in.dv_o = 0.f;
  if (!p.first_step) { in.dh_o = s.DH[c][tid]; in.du_o = s.DU[c][tid]; in.dv_o = s.DV[c][tid]; }
  const SweK2Out o = swe_k2_cell(p, in);
  s.UP[c][tid] = o.u;
  s.VP[c][tid] = o.v;
  if (t.o_ok && r >= t.r_t0 && r < t.r_t1) {
    const long long off = t.off - 4 * (long long)t.pitch;
    a.h_o[off] = o.h; a.dh_o[off] = o.dh; a.du_o[off] = o.du; a.dv_o[off] = o.dv;
  }
}
// friction-u of row jr = r - 6
template <int Q, class SM>
__device__ __forceinline__ void strip_fric_u(const StripArgs& a, SM& s, const StripThr& t, int r) {
  if (!t.u_ok || r < t.r_u) return;
  const int tid = t.tid;
  constexpr int c = (Q + 2) & 3, n = (Q + 3) & 3, m = (Q + 1) & 3;
  s.UPP[c][tid] = swe_friction_u(a.p, s.UP[c][tid], s.UP[c][tid + 1], s.UP[c][tid - 1], s.UP[n][tid], s.UP[m][tid],
                                 false, false);
}
// friction-v of row jv = r - 7, and the step's u'', v'' to memory
template <int Q, class SM>
__device__ __forceinline__ void strip_fric_v(const StripArgs& a, SM& s, const StripThr& t, int r) {
  if (!t.o_ok || r < t.r_o0 || r >= t.r_o1) return;
  const int tid = t.tid;
  const B2SweParams& p = a.p;
  constexpr int c = (Q + 1) & 3, n = (Q + 2) & 3, k = (Q + 1) & 3, km = Q & 3;
  const float upp = s.UPP[k][tid], v_c = s.VP[c][tid];
  const float fe2_c = swe_visc_flux(p.c_nux, s.VP[c][tid + 1], upp);
  const float fe2_w = swe_visc_flux(p.c_nux, v_c, s.UPP[k][tid - 1]);
  const float fn2_c = swe_visc_flux(p.c_nuy, s.VP[n][tid], upp);
  const float fn2_s = swe_visc_flux(p.c_nuy, v_c, s.UPP[km][tid]);
  const long long off = t.off - 7 * (long long)t.pitch;
  a.u_o[off] = upp;
  a.v_o[off] = swe_apply_div(p, v_c, fe2_c, fe2_w, fn2_c, fn2_s);
}

// One row index (Q = r & 7, r >= 0).  `each(f)` runs f(thread state) for the CTA's threads and then
// synchronises them: on the device it is the calling thread + __syncthreads, the host emulation loops
// over all threads.  No phase reads what another phase of the same iteration writes.
template <int Q, class SM, class Each>
__device__ __forceinline__ void strip_row(const StripArgs& a, SM& s, int r, Each&& each) {
  each([&](StripThr& t) {
    strip_issue<Q>(a, s, t, r);
    strip_flux<Q>(a, s, t, r);
    strip_tend<Q>(a, s, t, r);
    strip_fric_u<Q>(a, s, t, r);
    strip_fric_v<Q>(a, s, t, r);
    t.off += (long long)t.pitch;
  });
}
// all rows of one CTA: r runs from the multiple of eight at or below j0 - 4 to j1 + 6
template <class SM, class Each>
__device__ __forceinline__ void strip_cta(const StripArgs& a, SM& s, const StripGeo& g, Each&& each) {
  const int r0 = (g.j0 - 4) & ~7;
  each([&](StripThr& t) {                                           // prime the pipeline: rows r0 .. r0 + 2
    t.off = (long long)(r0 - 3) * (long long)t.pitch + (long long)t.col;
    strip_issue<5>(a, s, t, r0 - 3);                                // (r0 is a multiple of 8: (r0 - 3) & 7 = 5)
    t.off += (long long)t.pitch;
    strip_issue<6>(a, s, t, r0 - 2);
    t.off += (long long)t.pitch;
    strip_issue<7>(a, s, t, r0 - 1);
    t.off += (long long)t.pitch;
  });
  for (int rb = r0; rb <= g.j1 + 6; rb += 8) {
    strip_row<0>(a, s, rb, each);
    strip_row<1>(a, s, rb + 1, each);
    strip_row<2>(a, s, rb + 2, each);
    strip_row<3>(a, s, rb + 3, each);
    strip_row<4>(a, s, rb + 4, each);
    strip_row<5>(a, s, rb + 5, each);
    strip_row<6>(a, s, rb + 6, each);
    strip_row<7>(a, s, rb + 7, each);
  }
}
