// tcgen05 BPTT kernel, second generation (sm_100a): LAYER-PIPELINED warp specialisation.
//
// The first-generation kernel (cwlstm_tc_bwd.cuh) walks one 128-coordinate tile through
//   P0 (checkpoint rows -> A) -> Z MMAs -> layer-2 backward -> dX2 MMAs -> layer-1 backward -> dX1 MMAs
// with ALL epilogue warps in every phase, so each step pays three exposed MMA round trips and the epilogue warps idle
// through them (ncu r01: warps_active 19 %, tensor pipe 30 %).  The dependency graph of truncated BPTT is looser:
//   layer-2 chain:  dX2(t+1) -> [A2(t) -> Z2(t)] -> dZ2(t) -> dX2(t) -> ...         (needs nothing from layer 1)
//   layer-1 chain:  dX1(t+1), dX2(t) -> [A1(t) -> Z1(t)] -> dZ1(t) -> dX1(t) -> ...
// so here the two chains run CONCURRENTLY on their own warps (8 warps per layer, a thread pair per coordinate owning
// hidden units 0..9 | 10..19 of ONE layer), layer 1 trailing layer 2 by about half a step, and a polling issuer warp
// feeds the tensor pipe from whichever chain is ready.  What makes it fit:
//   * dW^T = dZ^T.X from bf16 hi/lo operands (tcgen05.mma.kind::f16, fp32 accumulate): x = hi + lo with
//     hi = bf16_rn(x), lo = bf16_rn(x - hi) (16 mantissa bits; products hi.hi + hi.lo + lo.hi).  dW is a pure
//     accumulation over N*T terms, so the 2^-17 representation error averages out (scripts/tc_accuracy.py).  The
//     staging drops from 128 KB (tf32 hi/lo, one buffer shared by both layers) to 64 KB PER LAYER (MN-major
//     SWIZZLE_128B, the canonical 16-bit layout), and the SS-mode MMA count halves (K = 16 per instruction).
//   * Z and dX stay on error-compensated 3xTF32 (they feed the nonlinear recurrences).
//   * TMEM (496 of 512 columns): dZ_hi is written IN PLACE over the Z accumulators a thread has just read, dZ_lo over
//     the dead A operand; dX results have their own columns so the next Z can be issued behind them.
//       Z2|dZ2hi 80 | A2|dZ2lo 96 | dX2 48 | Z1|dZ1hi 80 | A1|dZ1lo 80 | dX1 32 | dW2^T 48 | dW1^T 32
//   * The dW staging (bf16 conversion + shared-memory stores) runs AFTER dZ has been handed to the dX MMAs, in the
//     shadow of that round trip, and the dW MMAs are issued only when no Z / dX work is waiting (in-order tensor pipe).
// Semantics: SURVEY.md Appendix B (derived from DM/meta.py:319-376, second_derivatives=False); imitation mode
// DM/meta_dm_train.py:472-475.
#pragma once
#include <cuda_bf16.h>

#include "cwlstm_tc_bwd.cuh"

namespace l2o {
namespace tcb2 {

using namespace tc;

constexpr int kEpiThreads2 = 512;              // 16 epilogue warps: 0-7 layer 2, 8-15 layer 1
constexpr int kThreads2 = kEpiThreads2 + 128;  // + the warpgroup holding the issuer warp
constexpr int kEpiRegs2 = 112, kIssRegs2 = 32; // setmaxnreg targets (pool: 640 x 96)
constexpr int kSoloRegs2 = 152;                // workers of a single-chain pass (the idle chain's warps drop to 24)
constexpr int kNU = 10;                        // hidden units per thread
// TMEM columns
constexpr int cZ2 = 0, cR2 = 80, cX2 = 176, cZ1 = 224, cR1 = 304, cX1 = 384, cW2 = 416, cW1 = 464;
static_assert(cW1 + 32 <= kTmemCols, "TMEM budget");
// A2 (48 columns, the order of the BPTT weight image B2' = ext_weight_bwd rows 16..63):
//   [0..3 zero-weight | 4..7 feature chunk (only the constant 1 at 4+F meets a non-zero row: b2) | 8..27 h1n | 28..47 h2p]
constexpr int kA2One = 4, kA2H1N = 8, kA2H2P = 28, kA2Cols = 48;
// A1 (24 columns): [0..19 h1p | 20..23 feature chunk (u, 1)]
constexpr int kA1Chunk = 20, kA1Cols = 24;
// staged bf16 operand Y_l = [X_l (slots 0..47) | dZ_l (slots 48..127)] per coordinate, MN-major SWIZZLE_128B:
//   byte(slot, c) = (slot/64)*kLBO16 + (c/8)*kSBO16 + (c%8)*128 + (((slot%64)/8) ^ (c%8))*16 + (slot%8)*2
// X2 slots: [h1n 0..19 | h2p 20..39 | constant one 40 | zero 41..47];  X1 slots: [h1p 0..19 | chunk 20..23 | zero]
constexpr int kYZ16 = 48, kX2H2P = 20, kX2One = 40, kX1Chunk = 20;
constexpr uint32_t kLBO16 = 1024, kSBO16 = 2048;
constexpr int kY16Elems = 128 * 128;           // 32 KB per hi / lo buffer

#ifdef L2O_TC_PROF2
// timeline instrumentation (scripts/tc_bwd2_prof.cu only): clock64 stamps of CTA 0; role 0 = layer-2 worker (warp 0 lane 0),
// 1 = layer-1 worker (warp 8 lane 0), 2 = issuer (event id in the low 3 bits)
__device__ long long g_prof2[3 * 4096];
#define L2O_PROF2(role, idx, tag) \
  do { if (blockIdx.x == 0 && (idx) < 4096) g_prof2[(role) * 4096 + (idx)] = (clock64() << 3) | (tag); } while (0)
#else
#define L2O_PROF2(role, idx, tag) do { } while (0)
#endif

struct SmemB2 {
  uint16_t y2h[kY16Elems];     // 1024-B aligned (first member)
  uint16_t y2l[kY16Elems];
  uint16_t y1h[kY16Elems];
  uint16_t y1l[kY16Elems];
  float img[kImgAllFloats];    // B1'h|B1'l|B2'h|B2'l (K-major) | T1h|T1l|T2h|T2l (transposed, K-major)
  float wo[kH + 4];
  float win[64];               // fc nets: Win [2][20] | bin [20]
  uint64_t wbar;
  uint64_t a_ready[2], dz_ready[2];               // index 0 = layer 2, 1 = layer 1 (epilogue -> issuer, 256 arrivals)
  uint64_t z_done[2], x_done[2], w_done[2];       // issuer (tcgen05.commit) -> epilogue
  uint64_t x2_taken;                              // layer-1 threads have read dX2(t)[h1n] (256 arrivals)
  uint64_t flushed[2];                            // the layer's half-0 threads have drained the dW accumulators of a tile (128)
  uint32_t tmem_slot, pad;
};
static_assert(sizeof(SmemB2) + 1024 <= 227 * 1024, "shared memory budget");

__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// ---- 10-unit chunking -------------------------------------------------------------------------------------------
// A thread owns 10 hidden units of one layer, walked as three chunks of 4 | 4 | 2 units.  Half 0 owns units 0..9 as
// [0-3][4-7][8-9]; half 1 owns units 10..19 walked as [12-15][16-19][10-11], i.e. the SAME chunk shapes with different
// base units, so both halves (and every lane quarter) execute ONE instruction stream: the half is a run-time value
// (ncu r02e: 23 % of the stall samples of the per-half-specialised kernel were instruction-fetch stalls, the four
// role-specialised epilogues being 4 x 27 KB of code).  Every chunk stays naturally aligned for the 16/8-byte global
// accesses, the x4/x2 (one column per unit) and x16/x8 (four gate columns per unit) TMEM accesses and the 16-byte
// staging groups.  f(K0, NC, ub): K0 = index of the chunk's first unit in the thread's arrays, NC = units in the
// chunk (both compile-time), ub = the chunk's first hidden unit (run-time).
struct UnitMap {
  int u0, u1, u2;   // base units of the three chunks
  __device__ __forceinline__ explicit UnitMap(int half) : u0(half ? 12 : 0), u1(half ? 16 : 4), u2(half ? 10 : 8) {}
};
template <class F>
__device__ __forceinline__ void chunks10(const UnitMap& um, F&& f) {
  f(IC<0>{}, IC<4>{}, um.u0);
  f(IC<4>{}, IC<4>{}, um.u1);
  f(IC<8>{}, IC<2>{}, um.u2);
}
#define L2O_CHUNK3(K0, NC, k0c, ncc) L2O_CHUNK(K0, NC, k0c, ncc)
__device__ __forceinline__ void load10(const UnitMap& um, const float* __restrict__ p, float* v) {  // p -> unit 0 of the row
#ifdef L2O_BWD_NOLOAD   // timing experiment only (scripts/tc_bwd2_prof.cu): how much of a step the strided row loads cost
#pragma unroll
  for (int k = 0; k < kNU; ++k) v[k] = 0.01f * (float)(k + 1);
  return;
#endif
  chunks10(um, [&](auto k0c, auto ncc, int ub) {
    L2O_CHUNK(K0, NC, k0c, ncc);
    if constexpr (NC == 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + ub);
      v[K0] = t.x; v[K0 + 1] = t.y; v[K0 + 2] = t.z; v[K0 + 3] = t.w;
    } else {
      const float2 t = *reinterpret_cast<const float2*>(p + ub);
      v[K0] = t.x; v[K0 + 1] = t.y;
    }
  });
}
__device__ __forceinline__ void store10(const UnitMap& um, float* __restrict__ p, const float* v) {  // p -> unit 0 of the row
  chunks10(um, [&](auto k0c, auto ncc, int ub) {
    L2O_CHUNK(K0, NC, k0c, ncc);
    if constexpr (NC == 4) *reinterpret_cast<float4*>(p + ub) = make_float4(v[K0], v[K0 + 1], v[K0 + 2], v[K0 + 3]);
    else *reinterpret_cast<float2*>(p + ub) = make_float2(v[K0], v[K0 + 1]);
  });
}
__device__ __forceinline__ int unit_of(const UnitMap& um, int k) {   // array index -> hidden unit
  return (k < 4 ? um.u0 : (k < 8 ? um.u1 - 4 : um.u2 - 8)) + k;
}
// column of the constant 1 inside the 4-wide feature chunk of A2: behind the F <= 3 features, or first for fc nets
// (their layer-1 inputs are the 20 fc outputs, not a chunk)
template <class C>
struct Bias { static constexpr int kCol = C::FC ? 0 : C::F; };
// 10 per-unit values -> TMEM columns col0 + unit as 3xTF32 hi (at t_hi) / lo (at t_lo)   (col0 = column of unit 0)
__device__ __forceinline__ void st_split10(const UnitMap& um, uint32_t t_hi, uint32_t t_lo, int col0, const float* v) {
  chunks10(um, [&](auto k0c, auto ncc, int ub) {
    L2O_CHUNK(K0, NC, k0c, ncc);
    if constexpr (NC == 4) st_split4(t_hi, t_lo, col0 + ub, v + K0);
    else st_split2(t_hi, t_lo, col0 + ub, v + K0);
  });
}
// per-unit columns col0 + unit of this thread's lane -> registers (loads in flight behind one wait)
__device__ __forceinline__ void ld10(const UnitMap& um, uint32_t t_base, int col0, float* v) {
  chunks10(um, [&](auto k0c, auto ncc, int ub) {
    L2O_CHUNK(K0, NC, k0c, ncc);
    tmem_ldn<NC>(t_base + col0 + ub, v + K0);
  });
  tc_wait_ld();
}

// ---- bf16 hi/lo staging ------------------------------------------------------------------------------------------
// x = hi + lo, hi = bf16_rn(x), lo = bf16_rn(x - hi): pack two values per 32-bit word (first value in the low half)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - ha, b - hb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// byte offset of slot s (s % NS == 0 for an NS-slot store) of coordinate c inside a Y buffer
__device__ __forceinline__ uint32_t y16_off(int c, int s) {
  return (uint32_t)((s >> 6) * (int)kLBO16 + (c >> 3) * (int)kSBO16 + (c & 7) * 128 + ((((s & 63) >> 3) ^ (c & 7)) << 4) +
                    (s & 7) * 2);
}
// explicit shared-window stores (a generic store to a shared address pays the window check on every access)
__device__ __forceinline__ void sts128(uint32_t sa, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t sa, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sa), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t sa, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(sa), "r"(a) : "memory");
}
// NV values (2, 4 or 8) -> slots [s, s+NV) of coordinate c in the hi / lo buffers (yh, yl: shared-window addresses)
template <int NV>
__device__ __forceinline__ void stage16(uint32_t yh, uint32_t yl, int c, int s, const float* v) {
  const uint32_t off = y16_off(c, s);
  uint32_t h[NV / 2], l[NV / 2];
#pragma unroll
  for (int k = 0; k < NV / 2; ++k) split_bf16x2(v[2 * k], v[2 * k + 1], h[k], l[k]);
  if constexpr (NV == 8) {
    sts128(yh + off, h[0], h[1], h[2], h[3]);
    sts128(yl + off, l[0], l[1], l[2], l[3]);
  } else if constexpr (NV == 4) {
    sts64(yh + off, h[0], h[1]);
    sts64(yl + off, l[0], l[1]);
  } else {
    sts32(yh + off, h[0]);
    sts32(yl + off, l[0]);
  }
}
// the thread's 10 per-unit values -> slots s0 + unit   (s0 = slot of unit 0)
__device__ __forceinline__ void stage_units10(const UnitMap& um, uint32_t yh, uint32_t yl, int c, int s0, const float* v) {
  chunks10(um, [&](auto k0c, auto ncc, int ub) {
    L2O_CHUNK(K0, NC, k0c, ncc);
    stage16<NC>(yh, yl, c, s0 + ub, v + K0);
  });
}

// activated gates -> dz for the NC units of one chunk; shared by both layers.
//   z: 4*NC accumulator columns (i, j, f, o per unit), overwritten with dz
//   cprev / dh / dc: per-unit arrays of the chunk;  hn (optional): h' = tanh(c') o
template <int NC>
__device__ __forceinline__ void chunk_bwd(float* z, const float* cprev, const float* dh, float* dc, float* hn) {
#pragma unroll
  for (int u = 0; u < NC; ++u) {
    float* g = z + 4 * u;
    g[0] = sigmoid_fast(g[0]);
    g[1] = tanh_fast(g[1]);
    g[2] = sigmoid_fast(g[2] + 1.0f);
    g[3] = sigmoid_fast(g[3]);
    const float cn = fmaf(g[2], cprev[u], g[0] * g[1]);
    const float tcn = tanh_fast(cn);
    if (hn) hn[u] = tcn * g[3];
    tcb::unit_bwd(g, cprev[u], tcn, dh[u], dc[u]);
  }
}
// dz chunk -> TMEM: hi IN PLACE over the accumulator columns just read, lo into the dead A region
template <int NC>
__device__ __forceinline__ void put_dz(uint32_t t_hi, uint32_t t_lo, int col, const float* dz) {
#pragma unroll
  for (int u = 0; u < NC; ++u) st_split4(t_hi, t_lo, col + 4 * u, dz + 4 * u);
}

// ---- dW^T accumulator flush -------------------------------------------------------------------------------------------
// The tensor core adds into its fp32 accumulators with truncation, so an accumulation that runs over the CTA's whole
// share of the problem (53 tiles x 100 steps x 24 MMAs at 1 M coordinates) drifts: measured 2.4e-5 of max|dtheta|
// against a chunked fp64 reference (scripts/tc_accuracy_large.py), 9.7e-7 at 3.5 tiles per CTA.  The accumulators are
// therefore drained into the fp64 dtheta after EVERY tile (5,040 fp64 atomics per tile: ~0.2 ms per 1 M x 100 unroll)
// and the next tile's first MMA overwrites instead of accumulating.  Lane = slot: gate row m - 48 = 4u + g.
template <class C>
__device__ __forceinline__ void flush_dw2(const l2o_bwd_args& a, uint32_t tl, int c) {
  const int m = c - kYZ16;
  const int col = (m & 3) * kH + (m >> 2);
#pragma unroll
  for (int k4 = 0; k4 < 48 / 4; ++k4) {
    float v[4];
    tcb::tmem_ld4(tl + cW2 + 4 * k4, v);
    if (m >= 0 && m < kN) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 4 * k4 + e;
        int idx = -1;
        if (k < 2 * kH) idx = C::O_W2 + k * C::G2 + col;   // rows of lstm_2/w_gates: h1 (0..19) then h2 (20..39)
        else if (k == kX2One) idx = C::O_B2 + col;
        if (idx >= 0) atomicAdd(&a.dtheta[idx], (double)v[e]);
      }
    }
  }
}
template <class C>
__device__ __forceinline__ void flush_dw1(const l2o_bwd_args& a, uint32_t tl, int c) {
  const int m = c - kYZ16;
  const int col = (m & 3) * kH + (m >> 2);
#pragma unroll
  for (int k4 = 0; k4 < 32 / 4; ++k4) {
    float v[4];
    tcb::tmem_ld4(tl + cW1 + 4 * k4, v);
    if (m >= 0 && m < kN) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 4 * k4 + e;
        int idx = -1;
        if (k < kH) idx = C::O_W1 + (C::F + k) * C::G1 + col;
        else if (k < kX1Chunk + C::F) idx = C::O_W1 + (k - kX1Chunk) * C::G1 + col;
        else if (k == kX1Chunk + C::F) idx = C::O_B1 + col;
        if (idx >= 0) atomicAdd(&a.dtheta[idx], (double)v[e]);
      }
    }
  }
}

// fc nets, layer-1 pass: the chain runs on the layer-2 resources; X slots = [h1p 0..19 | fc outputs e 20..39 | one 40]
template <class C>
__device__ __forceinline__ void flush_dwfc(const l2o_bwd_args& a, uint32_t tl, int c) {
  const int m = c - kYZ16;
  const int col = (m & 3) * kH + (m >> 2);
#pragma unroll
  for (int k4 = 0; k4 < 48 / 4; ++k4) {
    float v[4];
    tcb::tmem_ld4(tl + cW2 + 4 * k4, v);
    if (m >= 0 && m < kN) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = 4 * k4 + e;
        int idx = -1;
        if (k < kH) idx = C::O_W1 + (C::F + k) * C::G1 + col;          // lstm_1/w_gates rows: inputs (F) first, then h1
        else if (k < 2 * kH) idx = C::O_W1 + (k - kH) * C::G1 + col;
        else if (k == kX2One) idx = C::O_B1 + col;
        if (idx >= 0) atomicAdd(&a.dtheta[idx], (double)v[e]);
      }
    }
  }
}

// =====================================================================================================================
// layer-2 workers: output layer + layer-2 LSTM backward
// =====================================================================================================================
// EXPORT (single-chain pass A of fc nets): dX2(t)[h1n], which the layer-1 threads of the pipelined kernel read straight
// from TMEM, goes to a.scratch [T][n][20] for the layer-1 pass.
template <class C, bool EXPORT>
__device__ __forceinline__ void layer2_worker(const l2o_bwd_args& a, const NetRt& rt, SmemB2& S, uint32_t tmem_base, int warp,
                                              int lane) {
  const int half = (warp >> 2) & 1;   // run-time: both halves share one instruction stream
  const UnitMap um(half);
  const int q = warp & 3;
  const int c = q * 32 + lane;
  const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t tZ = tl + cZ2, tRh = tl + cR2, tRl = tl + cR2 + kA2Cols, tDl = tl + cR2, tX = tl + cX2;
  const uint32_t yh = smem_u32(S.y2h), yl = smem_u32(S.y2l);
  const int T = a.T;
  const int64_t n = a.n, slot = n * C::SF, ntiles = (n + 127) / 128;
  const bool imit = a.labels != nullptr;
  const float inv_nt = imit ? 1.0f / (float)a.n_total : 0.f;
  uint32_t pz = 0, px = 0, pw = 0;
  bool have_prev = false;
  int pi = 0;
  const bool prof = (half == 0 && q == 0 && lane == 0);
  (void)pi; (void)prof;
  int64_t exp_row = -1;   // EXPORT: scratch row (t * n + i) of the step whose dX2 is in flight, -1 = inactive lane
  float acc_wo[kNU], acc_bo = 0.f;
#pragma unroll
  for (int k = 0; k < kNU; ++k) acc_wo[k] = 0.f;
  if (half == 0) {  // zero the persistent dW2^T accumulators (lane = slot row); layer-1 half 0 does dW1^T
#pragma unroll
    for (int k = 0; k < 48 / 4; ++k) tmem_st4(tl + cW2 + 4 * k, 0.f, 0.f, 0.f, 0.f);
    tc_wait_st();
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t i = tile * 128 + c;
    const bool act = i < n;
    float dc2[kNU], dh2c[kNU];
#pragma unroll
    for (int k = 0; k < kNU; ++k) { dc2[k] = 0.f; dh2c[k] = 0.f; }
    float lam = (act && !imit) ? a.g_rec[(int64_t)T * n + i] : 0.f;
    for (int t = T - 1; t >= 0; --t) {
      const float* ck = a.ckpt + (int64_t)t * slot;
      // ---- P0: checkpoint rows (h1n(t) IS the checkpointed h1 of slot t+1), A2 = [0 | 0..1..0 | h1n | h2p] --------
      float c2p[kNU];
      float g_t = 0.f, dtanh = 1.0f;
      if (prof) { L2O_PROF2(0, pi, 0); ++pi; }
      {
        float h1n[kNU], h2p[kNU];
#pragma unroll
        for (int k = 0; k < kNU; ++k) { h1n[k] = 0.f; h2p[k] = 0.f; c2p[k] = 0.f; }
        if (act) {
          load10(um, ck + slot + i * kH, h1n);
          load10(um, ck + 2 * n * kH + i * kH, h2p);
          load10(um, ck + 2 * n * kH + (n + i) * kH, c2p);
          if (t > 0) {  // pull the following step's rows towards L2
            const float* nk = ck - slot;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nk + 2 * n * kH + i * kH + um.u2));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nk + 2 * n * kH + (n + i) * kH + um.u2));
          }
        }
        if (imit && act) lam = (a.delta_seq[(int64_t)t * n + i] - a.labels[(int64_t)t * n + i]) * inv_nt;
        if (!imit && act) g_t = a.g_rec[(int64_t)t * n + i];   // consumed at the end of the step (suffix sum)
        if (rt.tanh_output && act) {   // delta = scale tanh(y): the forward pass's recorded delta gives tanh' without y
          const float th = a.delta_seq[(int64_t)t * n + i] / rt.scale;
          dtanh = fmaf(-th, th, 1.0f);
        }
        if (have_prev) {  // dX2 of the previous step: the h2p columns are this chain's carry, and the aliased
          mbar_wait(&S.x_done[0], px);  // dZ2-lo / A2 region becomes writable
          px ^= 1;
          tc_fence_after();
          if (t != T - 1) ld10(um, tX, kA2H2P, dh2c);
          if constexpr (EXPORT) {
            float v[kNU];
            ld10(um, tX, kA2H1N, v);
            if (exp_row >= 0) store10(um, a.scratch + exp_row * kH, v);
          }
        }
        if (prof) { L2O_PROF2(0, pi, 1); ++pi; }
        if (t == T - 1) {
#pragma unroll
          for (int k = 0; k < kNU; ++k) dh2c[k] = 0.f;
        }
        st_split10(um, tRh, tRl, kA2H1N, h1n);
        st_split10(um, tRh, tRl, kA2H2P, h2p);
        if (half == 1) {  // constant columns 0..7: zero-weight rows + the bias 1 at 4+F
          float one[4] = {0.f, 0.f, 0.f, 0.f};
          one[Bias<C>::kCol] = 1.0f;
          tmem_st4(tRh, 0.f, 0.f, 0.f, 0.f);
          tmem_st4(tRl, 0.f, 0.f, 0.f, 0.f);
          tmem_st4(tRh + kA2One, one[0], one[1], one[2], one[3]);
          tmem_st4(tRl + kA2One, 0.f, 0.f, 0.f, 0.f);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&S.a_ready[0]);
        if (prof) { L2O_PROF2(0, pi, 2); ++pi; }
        // X2 row of the dW2^T operand, from the same registers, in the shadow of the Z2 round trip.  The previous step's
        // dW2 MMAs were issued right behind its dX2 MMAs; by now they have drained (the wait is a formality).
        if (have_prev) {
          mbar_wait(&S.w_done[0], pw);
          pw ^= 1;
          if (t == T - 1 && half == 0) {   // first step of a new tile: drain the previous tile's dW2^T accumulators
            tc_fence_after();
            flush_dw2<C>(a, tl, c);
            tc_fence_before();
            mbar_arrive(&S.flushed[0]);
          }
        }
        if (prof) { L2O_PROF2(0, pi, 5); ++pi; }
        stage_units10(um, yh, yl, c, 0, h1n);
        stage_units10(um, yh, yl, c, kX2H2P, h2p);
      }
      // ---- layer-2 gates, output layer, layer-2 backward: dZ2 -> TMEM (hi in place, lo over A2) + bf16 staging ------
      const float dy = rt.scale * lam * dtanh;
      if (half == 1) acc_bo += dy;
      mbar_wait(&S.z_done[0], pz);
      pz ^= 1;
      tc_fence_after();
      if (prof) { L2O_PROF2(0, pi, 3); ++pi; }
      chunks10(um, [&](auto k0c, auto ncc, int ub) {
        L2O_CHUNK(K0, NC, k0c, ncc);
        float z[4 * NC], hn[NC], dh[NC];
        tmem_ldn<4 * NC>(tZ + 4 * ub, z);
#pragma unroll
        for (int u = 0; u < NC; ++u) dh[u] = fmaf(S.wo[ub + u], dy, dh2c[K0 + u]);
        tc_wait_ld();
        chunk_bwd<NC>(z, c2p + K0, dh, dc2 + K0, hn);
#pragma unroll
        for (int u = 0; u < NC; ++u) acc_wo[K0 + u] = fmaf(hn[u], dy, acc_wo[K0 + u]);
        put_dz<NC>(tZ, tDl, 4 * ub, z);
#pragma unroll
        for (int g8 = 0; g8 < NC / 2; ++g8) stage16<8>(yh, yl, c, kYZ16 + 4 * ub + 8 * g8, z + 8 * g8);
      });
      fence_proxy_async();
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&S.dz_ready[0]);
      if (prof) { L2O_PROF2(0, pi, 4); ++pi; }
      have_prev = true;
      if constexpr (EXPORT) exp_row = act ? (int64_t)t * n + i : -1;
      lam += g_t;
    }
  }
  if constexpr (EXPORT) {   // dX2 of the very last step
    if (have_prev) {
      mbar_wait(&S.x_done[0], px);
      px ^= 1;
      tc_fence_after();
      float v[kNU];
      ld10(um, tX, kA2H1N, v);
      if (exp_row >= 0) store10(um, a.scratch + exp_row * kH, v);
    }
  }
  // ---- flush: output-layer gradient from registers -------------------------------------------------------------------
#pragma unroll
  for (int k = 0; k < kNU; ++k) {
    float v = acc_wo[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(&a.dtheta[C::O_WO + unit_of(um, k)], (double)v);
  }
  if (half == 1) {
    float v = acc_bo;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) atomicAdd(&a.dtheta[C::O_BO], (double)v);
  }
  // ---- flush: the last tile's dW2^T accumulators ---------------------------------------------------------------------
  if (have_prev) {
    mbar_wait(&S.w_done[0], pw);
    pw ^= 1;
  }
  tc_fence_after();
  if (half == 0) flush_dw2<C>(a, tl, c);
}

// =====================================================================================================================
// layer-1 workers
// =====================================================================================================================
template <class C>
__device__ __forceinline__ void layer1_worker(const l2o_bwd_args& a, const NetRt& rt, SmemB2& S, uint32_t tmem_base, int warp,
                                              int lane) {
  const int half = (warp >> 2) & 1;   // run-time: both halves share one instruction stream
  const UnitMap um(half);
  const int q = warp & 3;
  const int c = q * 32 + lane;
  const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t tZ = tl + cZ1, tRh = tl + cR1, tRl = tl + cR1 + kA1Cols, tDl = tl + cR1, tX = tl + cX1, tX2 = tl + cX2;
  const uint32_t yh = smem_u32(S.y1h), yl = smem_u32(S.y1l);
  const int T = a.T;
  const int64_t n = a.n, slot = n * C::SF, ntiles = (n + 127) / 128;
  uint32_t pz = 0, px = 0, pw = 0, px2 = 0;
  bool have_prev = false;
  int pi = 0;
  const bool prof = (half == 0 && q == 0 && lane == 0);
  (void)pi; (void)prof;
  if (half == 0) {
#pragma unroll
    for (int k = 0; k < 32 / 4; ++k) tmem_st4(tl + cW1 + 4 * k, 0.f, 0.f, 0.f, 0.f);
    tc_wait_st();
  }
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t i = tile * 128 + c;
    const bool act = i < n;
    float dc1[kNU], dh1[kNU];
#pragma unroll
    for (int k = 0; k < kNU; ++k) { dc1[k] = 0.f; dh1[k] = 0.f; }
    for (int t = T - 1; t >= 0; --t) {
      const float* ck = a.ckpt + (int64_t)t * slot;
      // ---- P0: A1 = [h1p | u, 1] -----------------------------------------------------------------------------------------
      float c1p[kNU];
      float u4[4] = {0.f, 0.f, 0.f, 0.f};
      if (prof) { L2O_PROF2(1, pi, 0); ++pi; }
      {
        float h1p[kNU];
#pragma unroll
        for (int k = 0; k < kNU; ++k) { h1p[k] = 0.f; c1p[k] = 0.f; }
        float raw0 = 0.f;
        if (act) {
          load10(um, ck + i * kH, h1p);
          load10(um, ck + (n + i) * kH, c1p);
          if (half == 1) raw0 = a.in_seq[(int64_t)t * n + i];
          if (t > 0) {
            const float* nk = ck - slot;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nk + i * kH + um.u2));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nk + (n + i) * kH + um.u2));
          }
        }
        if (half == 1) {
          float uu[C::F];
          preprocess<C>(nullptr, rt, raw0, 0.f, uu);
#pragma unroll
          for (int k = 0; k < C::F; ++k) u4[k] = uu[k];
          u4[C::F] = 1.0f;
        }
        if (have_prev) {  // dX1 of the previous step: carry dh1c, and the aliased dZ1-lo / A1 region becomes writable
          mbar_wait(&S.x_done[1], px);
          px ^= 1;
          tc_fence_after();
          if (t != T - 1) ld10(um, tX, 0, dh1);
        }
        if (prof) { L2O_PROF2(1, pi, 1); ++pi; }
        if (t == T - 1) {
#pragma unroll
          for (int k = 0; k < kNU; ++k) dh1[k] = 0.f;
        }
        st_split10(um, tRh, tRl, 0, h1p);
        if (half == 1) st_split4(tRh, tRl, kA1Chunk, u4);
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&S.a_ready[1]);
        if (prof) { L2O_PROF2(1, pi, 2); ++pi; }
        // X1 row of the dW1^T operand from the same registers, in the shadow of the Z1 round trip
        if (have_prev) {
          mbar_wait(&S.w_done[1], pw);
          pw ^= 1;
          if (t == T - 1 && half == 0) {   // first step of a new tile: drain the previous tile's dW1^T accumulators
            tc_fence_after();
            flush_dw1<C>(a, tl, c);
            tc_fence_before();
            mbar_arrive(&S.flushed[1]);
          }
        }
        if (prof) { L2O_PROF2(1, pi, 6); ++pi; }
        stage_units10(um, yh, yl, c, 0, h1p);
        if (half == 1) stage16<4>(yh, yl, c, kX1Chunk, u4);
      }
      // ---- dh1 += dX2(t)[h1n]  (layer 2 has finished step t) -----------------------------------------------------------
      mbar_wait(&S.x_done[0], px2);
      px2 ^= 1;
      tc_fence_after();
      {
        float v[kNU];
        ld10(um, tX2, kA2H1N, v);
#pragma unroll
        for (int k = 0; k < kNU; ++k) dh1[k] += v[k];
      }
      tc_fence_before();
      mbar_arrive(&S.x2_taken);
      if (prof) { L2O_PROF2(1, pi, 3); ++pi; }
      // ---- layer-1 gates + backward: dZ1 -> TMEM + bf16 staging --------------------------------------------------------
      mbar_wait(&S.z_done[1], pz);
      pz ^= 1;
      tc_fence_after();
      if (prof) { L2O_PROF2(1, pi, 4); ++pi; }
      chunks10(um, [&](auto k0c, auto ncc, int ub) {
        L2O_CHUNK(K0, NC, k0c, ncc);
        float z[4 * NC];
        tmem_ldn<4 * NC>(tZ + 4 * ub, z);
        tc_wait_ld();
        chunk_bwd<NC>(z, c1p + K0, dh1 + K0, dc1 + K0, nullptr);
        put_dz<NC>(tZ, tDl, 4 * ub, z);
#pragma unroll
        for (int g8 = 0; g8 < NC / 2; ++g8) stage16<8>(yh, yl, c, kYZ16 + 4 * ub + 8 * g8, z + 8 * g8);
      });
      fence_proxy_async();
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&S.dz_ready[1]);
      if (prof) { L2O_PROF2(1, pi, 5); ++pi; }
      have_prev = true;
    }
  }
  if (have_prev) {
    mbar_wait(&S.w_done[1], pw);
    pw ^= 1;
  }
  tc_fence_after();
  if (half == 0) flush_dw1<C>(a, tl, c);
}


// =====================================================================================================================
// fc(20) nets (RNNProp, DM/networks.py:279-300), layer-1 pass.  Layer 1 of these nets has the SAME operand geometry as
// layer 2 (K = 48: [h1p | e | 1 | 0], dX N = 48, dW^T N = 48), so the chain runs on the layer-2 columns, staging buffers,
// barriers and issuer branch, with its own weight image in the B2'/T2 slots.  dX2(t)[h1n] comes from a.scratch (pass
// A).  The fc layer's own gradient (dWin [2 x 20], dbin [20]) needs de = dX1[20..39], which arrives one step late with
// the chain carry: it is folded into per-thread fp32 sums then (a thread owns the same 10 units of e as of h1).
// =====================================================================================================================
template <class C>
__device__ __forceinline__ void fc1_worker(const l2o_bwd_args& a, const NetRt& rt, SmemB2& S, uint32_t tmem_base, int warp,
                                           int lane, const float* __restrict__ win) {
  static_assert(C::FC && C::NIN == 2 && C::F == kH, "fc(2 -> 20) preprocessing");
  const int half = (warp >> 2) & 1;
  const UnitMap um(half);
  const int q = warp & 3;
  const int c = q * 32 + lane;
  const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
  const uint32_t tZ = tl + cZ2, tRh = tl + cR2, tRl = tl + cR2 + kA2Cols, tDl = tl + cR2, tX = tl + cX2;
  const uint32_t yh = smem_u32(S.y2h), yl = smem_u32(S.y2l);
  constexpr int kAE = kH, kAOne = 2 * kH;   // A columns / X slots of e and of the constant 1
  const int T = a.T;
  const int64_t n = a.n, slot = n * C::SF, ntiles = (n + 127) / 128;
  uint32_t pz = 0, px = 0, pw = 0;
  bool have_prev = false;
  float aw0[kNU], aw1[kNU], ab[kNU], ep[kNU];
#pragma unroll
  for (int k = 0; k < kNU; ++k) { aw0[k] = 0.f; aw1[k] = 0.f; ab[k] = 0.f; ep[k] = 0.f; }
  float r0p = 0.f, r1p = 0.f;   // the inputs of the step whose de is still in flight
  if (half == 0) {
#pragma unroll
    for (int k = 0; k < 48 / 4; ++k) tmem_st4(tl + cW2 + 4 * k, 0.f, 0.f, 0.f, 0.f);
    tc_wait_st();
  }
  auto take_de = [&]() {   // da = de * elu'(a);  dWin += [m~, g~]^T da, dbin += da
    float de[kNU];
    ld10(um, tX, kAE, de);
#pragma unroll
    for (int k = 0; k < kNU; ++k) {
      const float da = de[k] * ep[k];
      aw0[k] = fmaf(r0p, da, aw0[k]);
      aw1[k] = fmaf(r1p, da, aw1[k]);
      ab[k] += da;
    }
  };
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t i = tile * 128 + c;
    const bool act = i < n;
    float dc1[kNU], dh1[kNU];
#pragma unroll
    for (int k = 0; k < kNU; ++k) { dc1[k] = 0.f; dh1[k] = 0.f; }
    for (int t = T - 1; t >= 0; --t) {
      const float* ck = a.ckpt + (int64_t)t * slot;
      float c1p[kNU], dimp[kNU];
      {
        float h1p[kNU], e[kNU];
#pragma unroll
        for (int k = 0; k < kNU; ++k) { h1p[k] = 0.f; c1p[k] = 0.f; dimp[k] = 0.f; }
        float raw0 = 0.f, raw1 = 0.f;
        if (act) {
          load10(um, ck + i * kH, h1p);
          load10(um, ck + (n + i) * kH, c1p);
          load10(um, a.scratch + ((int64_t)t * n + i) * kH, dimp);
          raw0 = a.in_seq[((int64_t)t * 2) * n + i];
          raw1 = a.in_seq[((int64_t)t * 2 + 1) * n + i];
          if (t > 0) {
            const float* nk = ck - slot;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nk + i * kH + um.u2));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nk + (n + i) * kH + um.u2));
          }
        }
        if (have_prev) {  // dX1 of the previous step: chain carry + the fc layer's gradient of that step
          mbar_wait(&S.x_done[0], px);
          px ^= 1;
          tc_fence_after();
          take_de();
          if (t != T - 1) ld10(um, tX, 0, dh1);
        }
        if (t == T - 1) {
#pragma unroll
          for (int k = 0; k < kNU; ++k) dh1[k] = 0.f;
        }
        // e = elu([m~, g~] Win + bin) for this thread's units, as the forward kernel computes it (cwlstm_tc.cuh)
#pragma unroll
        for (int k = 0; k < kNU; ++k) {
          const int u = unit_of(um, k);
          const float av = fmaf(raw1, win[kH + u], fmaf(raw0, win[u], win[2 * kH + u]));
          const float em = ex2_approx(kLog2e * av) - 1.0f;
          const float et = av * fmaf(av, fmaf(av, fmaf(av, 1.0f / 24.0f, 1.0f / 6.0f), 0.5f), 1.0f);
          e[k] = av > 0.f ? av : (av > -0.0625f ? et : em);
          ep[k] = av > 0.f ? 1.0f : e[k] + 1.0f;   // elu' = exp(a) on the negative side
        }
        r0p = raw0;
        r1p = raw1;
        st_split10(um, tRh, tRl, 0, h1p);
        st_split10(um, tRh, tRl, kAE, e);
        if (half == 1) {
          tmem_st4(tRh + kAOne, 1.0f, 0.f, 0.f, 0.f);
          tmem_st4(tRl + kAOne, 0.f, 0.f, 0.f, 0.f);
          tmem_st4(tRh + kAOne + 4, 0.f, 0.f, 0.f, 0.f);
          tmem_st4(tRl + kAOne + 4, 0.f, 0.f, 0.f, 0.f);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(&S.a_ready[0]);
        if (have_prev) {
          mbar_wait(&S.w_done[0], pw);
          pw ^= 1;
          if (t == T - 1 && half == 0) {
            tc_fence_after();
            flush_dwfc<C>(a, tl, c);
            tc_fence_before();
            mbar_arrive(&S.flushed[0]);
          }
        }
        stage_units10(um, yh, yl, c, 0, h1p);
        stage_units10(um, yh, yl, c, kAE, e);
      }
#pragma unroll
      for (int k = 0; k < kNU; ++k) dh1[k] += dimp[k];   // dX2(t)[h1n], handed over by pass A
      mbar_wait(&S.z_done[0], pz);
      pz ^= 1;
      tc_fence_after();
      chunks10(um, [&](auto k0c, auto ncc, int ub) {
        L2O_CHUNK(K0, NC, k0c, ncc);
        float z[4 * NC];
        tmem_ldn<4 * NC>(tZ + 4 * ub, z);
        tc_wait_ld();
        chunk_bwd<NC>(z, c1p + K0, dh1 + K0, dc1 + K0, nullptr);
        put_dz<NC>(tZ, tDl, 4 * ub, z);
#pragma unroll
        for (int g8 = 0; g8 < NC / 2; ++g8) stage16<8>(yh, yl, c, kYZ16 + 4 * ub + 8 * g8, z + 8 * g8);
      });
      fence_proxy_async();
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&S.dz_ready[0]);
      have_prev = true;
    }
  }
  if (have_prev) {   // de of the very last step, then the dW1^T accumulators of the last tile
    mbar_wait(&S.x_done[0], px);
    px ^= 1;
    tc_fence_after();
    take_de();
    mbar_wait(&S.w_done[0], pw);
    pw ^= 1;
  }
#pragma unroll
  for (int k = 0; k < kNU; ++k) {
    float v0 = aw0[k], v1 = aw1[k], vb = ab[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, o);
      v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      vb += __shfl_xor_sync(0xffffffffu, vb, o);
    }
    if (lane == 0) {
      const int u = unit_of(um, k);
      atomicAdd(&a.dtheta[C::O_WIN + u], (double)v0);
      atomicAdd(&a.dtheta[C::O_WIN + C::F + u], (double)v1);
      atomicAdd(&a.dtheta[C::O_BIN + u], (double)vb);
    }
  }
  tc_fence_after();
  if (half == 0) flush_dwfc<C>(a, tl, c);
}

// MODE 0: both chains, layer-pipelined.  MODE 1 / 2 (fc nets): the layer-2 chain alone (exporting dX2[h1n]) / the fc layer-1
// chain alone on the layer-2 resources; warps 8-15 idle.
template <class C, int MODE>
__global__ void __launch_bounds__(kThreads2, 1) unroll_bwd2_kernel(l2o_bwd_args a, NetRt rt, const float* __restrict__ img) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  SmemB2& S = *reinterpret_cast<SmemB2*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = a.T;
  const int64_t ntiles = (a.n + 127) / 128;
  constexpr int kIssuerWarp = kEpiThreads2 / 32;

  {  // zero the staging buffers; X2's constant-one slot (bias row of dW2^T) is written once: hi = bf16(1.0), lo = 0
    uint32_t* y = reinterpret_cast<uint32_t*>(S.y2h);
    for (int k = threadIdx.x; k < 4 * kY16Elems / 2; k += blockDim.x) y[k] = 0u;
  }
  if (threadIdx.x < kH) S.wo[threadIdx.x] = a.theta[C::O_WO + threadIdx.x];
  if constexpr (C::FC) {
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 3 * kH) S.win[threadIdx.x - 64] = a.theta[C::O_WIN + threadIdx.x - 64];
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < 128; cc += blockDim.x)
    *reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(S.y2h) + y16_off(cc, kX2One)) = 0x3F80u;
  if (warp == kIssuerWarp) {
    if (lane == 0) {
      mbar_init(&S.wbar, 1);
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        mbar_init(&S.a_ready[l], 256);
        mbar_init(&S.dz_ready[l], 256);
        mbar_init(&S.z_done[l], 1);
        mbar_init(&S.x_done[l], 1);
        mbar_init(&S.w_done[l], 1);
      }
      mbar_init(&S.x2_taken, 256);
      mbar_init(&S.flushed[0], 128);
      mbar_init(&S.flushed[1], 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&S.tmem_slot, kTmemCols);
    tmem_relinquish();
    if (lane == 0) {
      mbar_expect_tx(&S.wbar, kImgAllBytes);
      tma_bulk_g2s(S.img, img, kImgAllBytes, &S.wbar);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_slot;

  if (warp < kIssuerWarp) {
    if constexpr (MODE == 0) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs2));
      if (warp < 8) layer2_worker<C, false>(a, rt, S, tmem_base, warp, lane);
      else layer1_worker<C>(a, rt, S, tmem_base, warp, lane);
    } else if (warp >= 8) {   // single-chain passes: the other chain's warps hand their registers to the workers
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(24));
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kSoloRegs2));
      if constexpr (MODE == 1) layer2_worker<C, true>(a, rt, S, tmem_base, warp, lane);
      else if constexpr (C::FC) fc1_worker<C>(a, rt, S, tmem_base, warp, lane, S.win);
    }
  } else if (warp > kIssuerWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kIssRegs2));
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kIssRegs2));
    mbar_wait(&S.wbar, 0);
    // ---- polling issuer: per layer the events arrive in the fixed order A(t), DZ(t), Y(t), A(t-1), ...  Z and dX
    // MMAs are issued the moment their operands are ready; the dW MMAs (SS-mode, ~60 cycles each on the in-order
    // tensor pipe) only when nothing else is waiting, one K-step (3 MMAs) per poll so a Z / dX request never queues
    // behind more than one of them.
    const uint32_t img_s = smem_u32(S.img);
    const uint64_t b1h = make_bdesc(img_s), b1l = make_bdesc(img_s + kB1Floats * 4);
    const uint64_t b2h = make_bdesc(img_s + 2 * kB1Floats * 4), b2l = make_bdesc(img_s + (2 * kB1Floats + kB2Floats) * 4);
    const uint32_t t_s = img_s + kImgFloats * 4;
    constexpr uint32_t kT1LBO = (kT1Rows / 8) * 128, kT2LBO = (kT2Rows / 8) * 128;
    const uint64_t t1h = tcb::make_desc(t_s, kT1LBO, 128), t1l = tcb::make_desc(t_s + kT1Floats * 4, kT1LBO, 128);
    const uint64_t t2h = tcb::make_desc(t_s + 2 * kT1Floats * 4, kT2LBO, 128);
    const uint64_t t2l = tcb::make_desc(t_s + (2 * kT1Floats + kT2Floats) * 4, kT2LBO, 128);
    const uint64_t y2h = tcb::make_desc(smem_u32(S.y2h), kLBO16, kSBO16, 2), y2l = tcb::make_desc(smem_u32(S.y2l), kLBO16, kSBO16, 2);
    const uint64_t y1h = tcb::make_desc(smem_u32(S.y1h), kLBO16, kSBO16, 2), y1l = tcb::make_desc(smem_u32(S.y1l), kLBO16, kSBO16, 2);
    constexpr uint32_t id_fwd = tcb::make_idesc_ex(kN, 0, 0);
    constexpr uint32_t id_dx2 = tcb::make_idesc_ex(48, 0, 0), id_dx1 = tcb::make_idesc_ex(32, 0, 0);
    constexpr uint32_t id_dw2 = make_idesc_bf16(48, 1, 1), id_dw1 = make_idesc_bf16(32, 1, 1);
    constexpr uint64_t kFwdStep = (2 * kLBO) >> 4;
    constexpr uint64_t kT1Step = (2 * kT1LBO) >> 4, kT2Step = (2 * kT2LBO) >> 4;
    constexpr uint64_t kY16Step = (2 * kSBO16) >> 4;   // K = 16 coordinates = two 8-coordinate groups
    const uint32_t tb = tmem_base;
    int my_tiles = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;
    const int total = my_tiles * T;                   // steps per layer (< 2^31: n*T/128/grid)
    int nA[2] = {0, 0}, nD[2] = {0, 0};               // events handled per layer (index 0 = layer 2)
    uint32_t pA[2] = {0, 0}, pD[2] = {0, 0}, pF[2] = {0, 0}, pT = 0;
    int pi = 0;
    (void)pi;
    if (MODE != 0) nA[1] = nD[1] = total;               // single-chain passes: chain 1 has nothing to do
    while (nD[0] < total || nD[1] < total) {
#pragma unroll
      for (int l = 0; l < (MODE == 0 ? 2 : 1); ++l) {
        // A(t): gate recompute Z_l = A_l . B_l'
        if (nA[l] < total && nA[l] == nD[l] && mbar_test(&S.a_ready[l], pA[l])) {
          pA[l] ^= 1;
          ++nA[l];
          tc_fence_after();
          if (elect_one()) {
            if (l == 0) {
#pragma unroll
              for (int kc = 0; kc < kK2 / 8; ++kc) {
                mma_tf32_ts(tb + cZ2, tb + cR2 + kA2Cols + 8 * kc, b2h + kc * kFwdStep, id_fwd, kc > 0 ? 1u : 0u);
                mma_tf32_ts(tb + cZ2, tb + cR2 + 8 * kc, b2l + kc * kFwdStep, id_fwd, 1u);
                mma_tf32_ts(tb + cZ2, tb + cR2 + 8 * kc, b2h + kc * kFwdStep, id_fwd, 1u);
              }
            } else {
#pragma unroll
              for (int kc = 0; kc < kK1 / 8; ++kc) {
                mma_tf32_ts(tb + cZ1, tb + cR1 + kA1Cols + 8 * kc, b1h + kc * kFwdStep, id_fwd, kc > 0 ? 1u : 0u);
                mma_tf32_ts(tb + cZ1, tb + cR1 + 8 * kc, b1l + kc * kFwdStep, id_fwd, 1u);
                mma_tf32_ts(tb + cZ1, tb + cR1 + 8 * kc, b1h + kc * kFwdStep, id_fwd, 1u);
              }
            }
            tc_commit(&S.z_done[l]);
          }
          __syncwarp();
          L2O_PROF2(2, pi, l == 0 ? 0 : 1); ++pi;
        }
        // DZ(t): dX_l = dZ_l . W_l^T (layer 2: only after layer 1 has taken dX2 of the previous step), then the
        // dW_l^T batch right behind it: the layer's own next request (Z of step t-1) comes a whole P0 later, by which
        // time the 24 SS-mode MMAs have drained
        if (nD[l] < nA[l] && mbar_test(&S.dz_ready[l], pD[l]) &&
            (MODE != 0 || l == 1 || nD[0] == 0 || mbar_test(&S.x2_taken, pT))) {
          if (MODE == 0 && l == 0 && nD[0] > 0) pT ^= 1;
          pD[l] ^= 1;
          // first step of a tile: its dW batch OVERWRITES the accumulators.  The previous tile's values were drained
          // by the workers before they signalled this step's a_ready, so the handshake never blocks here.
          const bool first_of_tile = nD[l] % T == 0;
          if (first_of_tile && nD[l] > 0) {
            mbar_wait(&S.flushed[l], pF[l]);
            pF[l] ^= 1;
          }
          ++nD[l];
          tc_fence_after();
          if (elect_one()) {
            if (l == 0) {
#pragma unroll
              for (int kc = 0; kc < kN / 8; ++kc) {
                mma_tf32_ts(tb + cX2, tb + cR2 + 8 * kc, t2h + kc * kT2Step, id_dx2, kc > 0 ? 1u : 0u);
                mma_tf32_ts(tb + cX2, tb + cZ2 + 8 * kc, t2l + kc * kT2Step, id_dx2, 1u);
                mma_tf32_ts(tb + cX2, tb + cZ2 + 8 * kc, t2h + kc * kT2Step, id_dx2, 1u);
              }
            } else {
#pragma unroll
              for (int kc = 0; kc < kN / 8; ++kc) {
                mma_tf32_ts(tb + cX1, tb + cR1 + 8 * kc, t1h + kc * kT1Step, id_dx1, kc > 0 ? 1u : 0u);
                mma_tf32_ts(tb + cX1, tb + cZ1 + 8 * kc, t1l + kc * kT1Step, id_dx1, 1u);
                mma_tf32_ts(tb + cX1, tb + cZ1 + 8 * kc, t1h + kc * kT1Step, id_dx1, 1u);
              }
            }
            tc_commit(&S.x_done[l]);
            const uint64_t yh0 = l == 0 ? y2h : y1h, yl0 = l == 0 ? y2l : y1l;
            const uint32_t d = tb + (l == 0 ? cW2 : cW1);
            const uint32_t id = l == 0 ? id_dw2 : id_dw1;
#pragma unroll
            for (int kb = 0; kb < 8; ++kb) {
              mma_bf16_ss(d, yl0 + kb * kY16Step, yh0 + kb * kY16Step, id, (kb == 0 && first_of_tile) ? 0u : 1u);
              mma_bf16_ss(d, yh0 + kb * kY16Step, yl0 + kb * kY16Step, id, 1u);
              mma_bf16_ss(d, yh0 + kb * kY16Step, yh0 + kb * kY16Step, id, 1u);
            }
            tc_commit(&S.w_done[l]);
          }
          __syncwarp();
          L2O_PROF2(2, pi, l == 0 ? 2 : 3); ++pi;
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kIssuerWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tcb2

template <class C>
int tc_launch_bwd2(const NetRt& rt, const l2o_bwd_args& a, float* img, cudaStream_t st, int sms) {
  tc::prep_weights_kernel<C><<<8, 256, 0, st>>>(a.theta, img, 1);
  const size_t smem = sizeof(tcb2::SmemB2) + 1024;
  const int64_t ntiles = (a.n + 127) / 128;
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  if constexpr (C::FC) {
    // two single-chain passes over time (layer 2, then layer 1 fed by a.scratch): both chains' K = 48 images plus both
    // staging buffers exceed the shared memory and TMEM of one CTA
    auto ka = tcb2::unroll_bwd2_kernel<C, 1>;
    auto kb = tcb2::unroll_bwd2_kernel<C, 2>;
    if (cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return L2O_E_CUDA;
    if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return L2O_E_CUDA;
    ka<<<grid, tcb2::kThreads2, smem, st>>>(a, rt, img);
    kb<<<grid, tcb2::kThreads2, smem, st>>>(a, rt, img + tc::kImgAllFloats);
  } else {
    auto k = tcb2::unroll_bwd2_kernel<C, 0>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return L2O_E_CUDA;
    k<<<grid, tcb2::kThreads2, smem, st>>>(a, rt, img);
  }
  return cudaGetLastError() == cudaSuccess ? L2O_OK : L2O_E_CUDA;
}

}  // namespace l2o
