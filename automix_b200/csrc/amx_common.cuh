// amx_common.cuh -- shared host/device utilities for the automix-b200 kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "amx.h"

namespace amx {

// ---- host-side runtime state (amx_api.cu) -------------------------------------------
int fail(int code, const char *fmt, ...);
cudaStream_t stream();
bool defer_sync();  // amx_set_deferred_sync: host-buffer state transfers only enqueue
void count_launch(unsigned n = 1);
int require_device();

#define AMX_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return amx::fail(AMX_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,             \
                       cudaGetErrorString(e_));                                          \
  } while (0)

#define AMX_CUDA_PTR(call)                                                               \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      amx::fail(AMX_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,                    \
                cudaGetErrorString(e_));                                                 \
      return nullptr;                                                                    \
    }                                                                                    \
  } while (0)

// ---- small device helpers --------------------------------------------------------------
// The reference's max/min are macros `(A) > (B) ? (A) : (B)` (automix.c:10-11); their NaN
// behaviour differs from fmax/fmin, and the accept rule depends on it, so mirror them.
__device__ __forceinline__ double max_m(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double min_m(double a, double b) { return a < b ? a : b; }
// Metropolis-Hastings acceptance probability exp(max(-30, min(0, dl))) (:612, :1063, :1247)
__device__ __forceinline__ double mh_prob(double dl) { return exp(max_m(-30.0, min_m(0.0, dl))); }
// u < mh_prob(dl), without the exponential when dl >= 0 (exp(0) = 1 > u for every u in [0,1)); NaN -> reject
__device__ __forceinline__ bool mh_accept(double u, double dl) { return (dl >= 0.0) ? true : (u < exp(max_m(-30.0, dl))); }

// Small arrays live in registers only if every index is a compile-time constant; these
// accessors turn a run-time index into a select chain for small N and a plain (local
// memory) access for large N.
constexpr int kRegArrayMax = 8;

template <int N>
__device__ __forceinline__ double aget(const double (&a)[N], int i) {
  if constexpr (N <= kRegArrayMax) {
    double r = a[0];
#pragma unroll
    for (int j = 1; j < N; j++) r = (i == j) ? a[j] : r;
    return r;
  } else {
    return a[i];
  }
}
template <int N>
__device__ __forceinline__ void aset(double (&a)[N], int i, double v) {
  if constexpr (N <= kRegArrayMax) {
#pragma unroll
    for (int j = 0; j < N; j++) a[j] = (i == j) ? v : a[j];
  } else {
    a[i] = v;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- counter-based per-chain RNG: Philox4x32-10 ---------------------------------------
// key = 64-bit seed, counter = (block index of this chain's stream, chain id).  One block yields four
// uniforms u = (w + 0.5) * 2^-32 in (0,1), one per 32-bit word: the resolution of the reference's own
// generator (sdrand returns a 31-bit integer times 2^-31, automix.c:1300-1305), but never 0 or 1 (the
// reference's can return exactly 0, which sends log(0) into Box-Muller).
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// (w + 0.5) * 2^-32 without an integer-to-double conversion: the word becomes the top of the mantissa of a
// double in [1, 2), and subtracting 1 - 2^-33 is exact.
__device__ __forceinline__ double u32_to_unit(uint32_t w) {
  return __hiloint2double((int)(0x3ff00000u | (w >> 12)), (int)(w << 20)) - (1.0 - 1.0 / 8589934592.0);
}

struct PhiloxStream {
  uint32_t k0, k1, id0, id1;
  unsigned long long n;  // uniforms consumed so far by this chain
  uint32_t w1, w2, w3;   // words of the current block not handed out yet, next first
  __device__ __forceinline__ void block(unsigned long long b, uint32_t (&c)[4]) const {
    c[0] = (uint32_t)b;
    c[1] = (uint32_t)(b >> 32);
    c[2] = id0;
    c[3] = id1;
    philox4x32_10(c, k0, k1);
  }
  __device__ __forceinline__ void open(unsigned long long seed, unsigned long long chain,
                                       unsigned long long consumed) {
    k0 = (uint32_t)seed;
    k1 = (uint32_t)(seed >> 32);
    id0 = (uint32_t)chain;
    id1 = (uint32_t)(chain >> 32);
    n = consumed;
    w1 = w2 = w3 = 0u;
    const int r = (int)(n & 3ull);
    if (r) {  // resume in the middle of a block
      uint32_t c[4];
      block(n >> 2, c);
      w1 = r == 1 ? c[1] : (r == 2 ? c[2] : c[3]);
      w2 = r == 1 ? c[2] : c[3];
      w3 = c[3];
    }
  }
  __device__ __forceinline__ double next() {
    uint32_t w;
    if ((n & 3ull) == 0) {
      uint32_t c[4];
      block(n >> 2, c);
      w = c[0];
      w1 = c[1];
      w2 = c[2];
      w3 = c[3];
    } else {
      w = w1;
      w1 = w2;
      w2 = w3;
    }
    n++;
    return u32_to_unit(w);
  }
  // random access: uniform number i of this chain's stream (what the i-th next() returns)
  __device__ __forceinline__ double at(unsigned long long i) const {
    uint32_t c[4];
    block(i >> 2, c);
    const int r = (int)(i & 3ull);
    return u32_to_unit(r == 0 ? c[0] : (r == 1 ? c[1] : (r == 2 ? c[2] : c[3])));
  }
  __device__ __forceinline__ bool overrun() const { return false; }
};

// Parity mode: the chain reads an injected tape in exactly the order in which the
// reference calls sdrand().
struct TapeStream {
  const double *p;
  unsigned long long n, len;
  bool over;
  __device__ __forceinline__ void open(const double *tape, unsigned long long stride,
                                       unsigned long long chain, unsigned long long consumed) {
    p = tape + chain * stride;
    len = stride;
    n = consumed;
    over = false;
  }
  __device__ __forceinline__ double next() {
    double r = 0.5;
    if (n < len) r = p[n];
    else over = true;
    n++;
    return r;
  }
  __device__ __forceinline__ double at(unsigned long long i) {
    if (i < len) return p[i];
    over = true;
    return 0.5;
  }
  __device__ __forceinline__ bool overrun() const { return over; }
};

// Box-Muller exactly as gauss() (automix.c:1639-1661): radius uniform first, angle second.
// (Keeping this transcendental part out of line was measured: 4.65e9 vs 4.76e9 chain-sweeps/s -- the call costs more
// than the instruction fetch it saves.)
__device__ __forceinline__ double box_muller_sin(double a, double b) {
  const double r = sqrt(-2.0 * log(a));
  return r * sin(6.283185307179586476925 * b);
}
__device__ __forceinline__ double2 box_muller_pair(double a, double b) {
  const double r = sqrt(-2.0 * log(a));
  double s, c;
  sincos(6.283185307179586476925 * b, &s, &c);
  return make_double2(r * s, r * c);
}
template <class U>
__device__ __forceinline__ double gauss_single(U &u) {
  const double a = u.next(), b = u.next();
  return box_muller_sin(a, b);
}
template <class U>
__device__ __forceinline__ void gauss_pair(U &u, double &z0, double &z1) {
  const double a = u.next(), b = u.next();
  const double2 z = box_muller_pair(a, b);
  z0 = z.x;
  z1 = z.y;
}

// ---- out-of-line forms for the wide sweep kernels -------------------------------------------------------------------
// The wide K3 configurations are bound by instruction fetch (ncu: "no instruction" is their largest stall; ~20 000 SASS
// instructions with a Philox block inlined at every draw and a Box-Muller at every normal).  There one copy of each,
// reached by a call, is the better trade; in the small configuration the inlined forms measured faster (above).
// Values in, values out: no address of the stream escapes, so its state stays in registers.
#ifndef AMX_RJ_WIDE_OUTLINE
#define AMX_RJ_WIDE_OUTLINE 1
#endif
static __device__ __noinline__ uint4 philox_block_ol(uint32_t b0, uint32_t b1, uint32_t id0, uint32_t id1, uint32_t k0,
                                                      uint32_t k1) {
  uint32_t c[4] = {b0, b1, id0, id1};
  philox4x32_10(c, k0, k1);
  return make_uint4(c[0], c[1], c[2], c[3]);
}
static __device__ __noinline__ double box_muller_sin_ol(double a, double b) { return box_muller_sin(a, b); }
static __device__ __noinline__ double2 box_muller_pair_ol(double a, double b) { return box_muller_pair(a, b); }
struct PhiloxStreamOL : PhiloxStream {
  __device__ __forceinline__ double next() {
    uint32_t w;
    if ((n & 3ull) == 0) {
      const uint4 c = philox_block_ol((uint32_t)(n >> 2), (uint32_t)(n >> 34), id0, id1, k0, k1);
      w = c.x;
      w1 = c.y;
      w2 = c.z;
      w3 = c.w;
    } else {
      w = w1;
      w1 = w2;
      w2 = w3;
    }
    n++;
    return u32_to_unit(w);
  }
};
__device__ __forceinline__ double gauss_single(PhiloxStreamOL &u) {
  const double a = u.next(), b = u.next();
  return box_muller_sin_ol(a, b);
}
__device__ __forceinline__ void gauss_pair(PhiloxStreamOL &u, double &z0, double &z1) {
  const double a = u.next(), b = u.next();
  const double2 z = box_muller_pair_ol(a, b);
  z0 = z.x;
  z1 = z.y;
}

// ---- Student-t proposals and random permutation (optional modes of the sampler) -----------------------
// Gamma(s,1) by rejection, three regimes, exactly the draw order of rgamma() (automix.c:1585-1637).
template <class U>
__device__ __forceinline__ double rgamma_dev(double s, U &u) {
  const double e1 = 2.718281828459045;  // exp(1.0)
  double out;
  if (s < 1.0) {
    const double b = (s + e1) / e1, inv = 1.0 / s;
    for (;;) {
      const double bu = b * u.next();
      if (bu <= 1.0) {
        const double t = inv * log(bu);
        out = exp(t < -30.0 ? -30.0 : t);
        if (u.next() < exp(-out)) break;
      } else {
        out = -log((b - bu) / s);
        if (u.next() < pow(out, s - 1.0)) break;
      }
      if (u.overrun()) break;
    }
  } else if (s == 1.0) {
    out = -log(u.next());
  } else {
    const double c1 = s - 1.0, c2 = (s - 1.0 / (6.0 * s)) / c1, c3 = 2.0 / c1, c4 = c3 + 2.0, c5 = 1.0 / sqrt(s);
    double w = 1.0;
    for (;;) {
      double u1 = u.next();
      const double u2 = u.next();
      if (s > 2.5) u1 = u2 + c5 * (1.0 - 1.86 * u1);
      if (u.overrun()) break;
      if (u1 <= 0.0 || u1 >= 1.0) continue;
      w = c2 * u2 / u1;
      if ((c3 * u1 + w + 1.0 / w) <= c4) break;
      if ((c3 * log(u1) - log(w) + w) >= 1.0) continue;
      break;
    }
    out = c1 * w;
  }
  return out;
}
// the common divisor of one rt() call: sqrt(Gamma(dof/2) / (dof/2)) (automix.c:1672-1677)
template <class U>
__device__ __forceinline__ double t_divisor(int dof, U &u) {
  const double s = 0.5 * dof;
  return sqrt(rgamma_dev(s, u) / s);
}
// log density of a t variate (automix.c:1717-1725); lt_const = lgamma((dof+1)/2) - lgamma(dof/2) - log(dof*pi)/2
__device__ __forceinline__ double ltprob_dev(int dof, double lt_const, double z) {
  return lt_const - 0.5 * (dof + 1) * log(1.0 + (z * z) / dof);
}

}  // namespace amx
