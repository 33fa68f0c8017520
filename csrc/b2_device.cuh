// mpi4jax_b200 -- device-side building blocks shared by every kernel:
//   * system-scope acquire/release accessors for flags that live in peer HBM
//   * spin-wait with a %globaltimer watchdog (deadlock -> diagnostics + trap,
//     the analogue of the reference's abort_on_error, mpi_ops_common.h:60-78)
//   * the block-paired cross-GPU barrier (CTA b on every rank <-> CTA b)
//   * the per-kernel ticket (staging parity + unique flag values, advanced on
//     the device so CUDA-graph replays stay correct)
//   * 16-byte vector movers and the typed reduction functors (10 ops x 15 dtypes,
//     fp32 accumulation for f16/bf16)
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "b2_common.h"

#define B2_THREADS 512

// ---------------------------------------------------------------------------
// memory-model helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void b2_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void b2_st_relaxed_sys(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned b2_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned b2_ld_relaxed_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned b2_ld_volatile(const unsigned* p) {
  unsigned v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void b2_st_volatile(unsigned* p, unsigned v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void b2_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// 16-byte accessors.  Peer data is read with strong.sys loads (no stale L1 line
// can be returned) and written with plain 16-byte stores.
__device__ __forceinline__ uint4 b2_ld_peer16(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ uint4 b2_ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void b2_st16(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void b2_st16_sys(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ unsigned long long b2_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------
// watchdog: fatal diagnostics then trap (host side prints "r<rank> | ... aborting")
// ---------------------------------------------------------------------------
static __device__ __noinline__ void b2_fatal(const B2DevComm& c, int code, int opcode, int peer,
                                      unsigned expected, unsigned observed, int aux) {
  B2ErrorRecord* e = c.err;
  if (e != nullptr && atomicCAS((int*)&e->code, 0, code) == 0) {
    e->rank = c.rank;
    e->opcode = opcode;
    e->peer = peer;
    e->expected = expected;
    e->observed = observed;
    e->block = (int)blockIdx.x;
    e->aux = aux;
    __threadfence_system();
  }
  __trap();
}

// Spin until `(int)(load(p) - want) >= 0` (monotone epochs, wrap-safe).
__device__ __forceinline__ void b2_wait_ge(const B2DevComm& c, const unsigned* p, unsigned want,
                                           int opcode, int peer) {
  unsigned v = b2_ld_acquire_sys(p);
  if ((int)(v - want) >= 0) return;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (true) {
    v = b2_ld_acquire_sys(p);
    if ((int)(v - want) >= 0) return;
    if ((++spins & 0xfffu) == 0) {
      unsigned long long now = b2_gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > c.timeout_ns) b2_fatal(c, B2_ERR_TIMEOUT, opcode, peer, want, v, 0);
    }
  }
}

// Spin until `load(p) == want` exactly (used where stale values may be larger).
__device__ __forceinline__ void b2_wait_eq(const B2DevComm& c, const unsigned* p, unsigned want,
                                           int opcode, int peer) {
  unsigned v = b2_ld_acquire_sys(p);
  if (v == want) return;
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (true) {
    v = b2_ld_acquire_sys(p);
    if (v == want) return;
    if ((++spins & 0xfffu) == 0) {
      unsigned long long now = b2_gtime();
      if (t0 == 0) t0 = now;
      else if (now - t0 > c.timeout_ns) b2_fatal(c, B2_ERR_TIMEOUT, opcode, peer, want, v, 1);
    }
  }
}

// ---------------------------------------------------------------------------
// block-paired cross-GPU barrier.
// CTA b of this rank signals CTA b of every peer and waits for all of them.
// All threads of the CTA must call it.  `e` is the CTA's next epoch value.
// release/acquire at .sys scope + bar.sync make every write issued by the CTA
// before the barrier visible to the peers' CTAs after it.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void b2_barrier_all(const B2DevComm& c, unsigned e, int opcode) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < c.size) {
    unsigned* remote =
        (unsigned*)(c.heap[t] + c.lay.flags_off) + (size_t)blockIdx.x * B2_MAX_RANKS + c.rank;
    b2_st_release_sys(remote, e);
    const unsigned* local =
        (const unsigned*)(c.heap[c.rank] + c.lay.flags_off) + (size_t)blockIdx.x * B2_MAX_RANKS + t;
    b2_wait_ge(c, local, e, opcode, t);
  }
  __syncthreads();
}

// The two halves of b2_barrier_all, for software pipelining: local work placed between the
// arrive and the wait (staging the next chunk, copying the previous one out) hides the NVLink
// round trip of the flags.  Epochs of one CTA must still be used in increasing order.
__device__ __forceinline__ void b2_barrier_arrive(const B2DevComm& c, unsigned e) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < c.size) {
    unsigned* remote =
        (unsigned*)(c.heap[t] + c.lay.flags_off) + (size_t)blockIdx.x * B2_MAX_RANKS + c.rank;
    b2_st_release_sys(remote, e);
  }
}
__device__ __forceinline__ void b2_barrier_wait(const B2DevComm& c, unsigned e, int opcode) {
  const int t = threadIdx.x;
  if (t < c.size) {
    const unsigned* local =
        (const unsigned*)(c.heap[c.rank] + c.lay.flags_off) + (size_t)blockIdx.x * B2_MAX_RANKS + t;
    b2_wait_ge(c, local, e, opcode, t);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// ticket: one value per kernel launch, identical on every CTA, advanced by the
// LAST CTA TO FINISH (so every CTA has already read it).  Lives in local device
// memory -> graph replays see fresh values without host involvement.
// Usage:  unsigned t = b2_ticket_read(ptr);  ...  b2_ticket_finish(ptr, ctr, 1);
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned b2_ticket_read(const unsigned* ticket) {
  __shared__ unsigned s_ticket;
  if (threadIdx.x == 0) s_ticket = b2_ld_volatile(ticket);
  __syncthreads();
  return s_ticket;
}

// All threads call; adds `delta` to *target once every one of `nblocks` CTAs is done.
__device__ __forceinline__ void b2_finish_bump(unsigned* target, unsigned* ctr, unsigned delta,
                                               unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned old = atomicAdd(ctr, 1u);
    if (old == nblocks - 1) {
      b2_st_volatile(ctr, 0u);
      b2_st_volatile(target, b2_ld_volatile(target) + delta);
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------
// byte movers (CTA-cooperative).  Fast path: both pointers 16-byte aligned.
// ---------------------------------------------------------------------------
template <bool PEER_SRC>
__device__ __forceinline__ uint4 b2_ld16(const void* p) {
  if (PEER_SRC) return b2_ld_peer16(p);
  return b2_ld_stream16(p);
}

template <bool PEER_SRC>
__device__ __forceinline__ void b2_copy_bytes(void* __restrict__ dst, const void* __restrict__ src,
                                              size_t n) {
  if (n == 0) return;
  const int t = threadIdx.x, nt = blockDim.x;
  if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
    const size_t nv = n >> 4;
    const char* s = (const char*)src;
    char* d = (char*)dst;
    size_t i = t;
    for (; i + 3 * (size_t)nt < nv; i += 4 * (size_t)nt) {
      uint4 v0 = b2_ld16<PEER_SRC>(s + (i << 4));
      uint4 v1 = b2_ld16<PEER_SRC>(s + ((i + nt) << 4));
      uint4 v2 = b2_ld16<PEER_SRC>(s + ((i + 2 * (size_t)nt) << 4));
      uint4 v3 = b2_ld16<PEER_SRC>(s + ((i + 3 * (size_t)nt) << 4));
      b2_st16(d + (i << 4), v0);
      b2_st16(d + ((i + nt) << 4), v1);
      b2_st16(d + ((i + 2 * (size_t)nt) << 4), v2);
      b2_st16(d + ((i + 3 * (size_t)nt) << 4), v3);
    }
    for (; i < nv; i += nt) b2_st16(d + (i << 4), b2_ld16<PEER_SRC>(s + (i << 4)));
    const size_t tail = n & 15;
    if ((size_t)t < tail) {
      const volatile unsigned char* sb = (const volatile unsigned char*)src + (nv << 4);
      ((unsigned char*)dst)[(nv << 4) + t] = sb[t];
    }
  } else {
    const volatile unsigned char* sb = (const volatile unsigned char*)src;
    unsigned char* db = (unsigned char*)dst;
    for (size_t i = t; i < n; i += nt) db[i] = sb[i];
  }
}

// ---------------------------------------------------------------------------
// typed reductions
// ---------------------------------------------------------------------------
struct b2_c64 { float re, im; };
struct b2_c128 { double re, im; };
struct b2_boolean { unsigned char v; };   // tag type: logical semantics on bytes

template <typename T> struct B2Traits {
  using Acc = T;
  static __device__ __forceinline__ Acc up(T v) { return v; }
  static __device__ __forceinline__ T down(Acc v) { return v; }
};
template <> struct B2Traits<__half> {
  using Acc = float;
  static __device__ __forceinline__ Acc up(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half down(Acc v) { return __float2half_rn(v); }
};
template <> struct B2Traits<__nv_bfloat16> {
  using Acc = float;
  static __device__ __forceinline__ Acc up(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 down(Acc v) { return __float2bfloat16_rn(v); }
};
template <> struct B2Traits<b2_boolean> {
  using Acc = unsigned char;
  static __device__ __forceinline__ Acc up(b2_boolean v) { return v.v != 0 ? 1 : 0; }
  static __device__ __forceinline__ b2_boolean down(Acc v) { b2_boolean r; r.v = v; return r; }
};

template <typename A> struct b2_is_complex { static constexpr bool value = false; };
template <> struct b2_is_complex<b2_c64> { static constexpr bool value = true; };
template <> struct b2_is_complex<b2_c128> { static constexpr bool value = true; };
template <typename A> struct b2_is_float { static constexpr bool value = false; };
template <> struct b2_is_float<float> { static constexpr bool value = true; };
template <> struct b2_is_float<double> { static constexpr bool value = true; };

template <int OP, typename A>
__device__ __forceinline__ A b2_apply(A a, A b) {
  if constexpr (b2_is_complex<A>::value) {
    A r;
    if constexpr (OP == B2_SUM) { r.re = a.re + b.re; r.im = a.im + b.im; }
    else { r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; }
    return r;
  } else if constexpr (b2_is_float<A>::value) {
    if constexpr (OP == B2_SUM) return a + b;
    else if constexpr (OP == B2_PROD) return a * b;
    else if constexpr (OP == B2_MIN) return (b < a || b != b) ? b : a;   // NaN-propagating like numpy
    else return (b > a || b != b) ? b : a;
  } else {
    if constexpr (OP == B2_SUM) return (A)(a + b);
    else if constexpr (OP == B2_PROD) return (A)(a * b);
    else if constexpr (OP == B2_MIN) return b < a ? b : a;
    else if constexpr (OP == B2_MAX) return b > a ? b : a;
    else if constexpr (OP == B2_LAND) return (A)((a != 0) && (b != 0));
    else if constexpr (OP == B2_LOR) return (A)((a != 0) || (b != 0));
    else if constexpr (OP == B2_LXOR) return (A)((a != 0) != (b != 0));
    else if constexpr (OP == B2_BAND) return (A)(a & b);
    else if constexpr (OP == B2_BOR) return (A)(a | b);
    else return (A)(a ^ b);
  }
}

// A 16-byte vector viewed as N elements of T, accumulated in B2Traits<T>::Acc.
template <typename T> struct B2Vec {
  static constexpr int N = 16 / (int)sizeof(T);
  using Acc = typename B2Traits<T>::Acc;
  Acc a[N];
  __device__ __forceinline__ void load(uint4 v) {
    alignas(16) T tmp[N];
    *reinterpret_cast<uint4*>(tmp) = v;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = B2Traits<T>::up(tmp[i]);
  }
  template <int OP> __device__ __forceinline__ void accumulate(uint4 v) {
    alignas(16) T tmp[N];
    *reinterpret_cast<uint4*>(tmp) = v;
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = b2_apply<OP, Acc>(a[i], B2Traits<T>::up(tmp[i]));
  }
  __device__ __forceinline__ uint4 store() const {
    alignas(16) T tmp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = B2Traits<T>::down(a[i]);
    return *reinterpret_cast<const uint4*>(tmp);
  }
};

// Store the first `nbytes` (<16) bytes of a vector to an arbitrary address.
__device__ __forceinline__ void b2_store_partial(void* dst, uint4 v, int nbytes) {
  alignas(16) unsigned char tmp[16];
  *reinterpret_cast<uint4*>(tmp) = v;
  unsigned char* d = (unsigned char*)dst;
  for (int i = 0; i < nbytes; ++i) d[i] = tmp[i];
}
