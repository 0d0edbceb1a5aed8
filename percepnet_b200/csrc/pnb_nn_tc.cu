// pnb_nn_tc.cu -- the gain network on Blackwell tensor cores (tcgen05 + TMEM + TMA), batched over streams.
//
// Replaces the reference's 35 scalar GEMVs per hop (/root/reference/src/nnet.cpp:59-200, rnn.cpp:42-81)
// for the layers that carry 99.9 % of the MACs: conv1, conv2, the five GRUs and the two output layers.  The rows
// of every contraction are the concurrent streams (M = S), so the operation is genuinely dense.
//
// Precision.  The parity bar (1e-4 relative on g/r, +-1 LSB on PCM) rules out plain 16-bit operands.
// Every fp32 operand is therefore split into two 16-bit terms whose sum reproduces it to ~2^-22, and the three
// significant products are issued as separate tensor-core MMAs into the same fp32 TMEM accumulator:
//   x w ~= xh wh + xh wl + xl wh
//   GRUs / output layers (inputs bounded by tanh): fp16 terms, both sides pre-scaled by powers of two
//   conv1 / conv2 (ReLU inputs, unbounded):        bf16 terms
//
// Two kernels, both persistent, warp-specialised, one CTA per SM, CTAs paired into clusters of two that act as
// ONE tensor-core unit (tcgen05 cta_group::2, M = 256):
//   tc_gemm_kernel   one dense layer over all rows (conv1, conv2, fc_gb, fc_rb run once per call over F*S rows)
//   gru_chain_kernel the five GRUs of all F hops of a call in ONE launch.  A tile (hop, layer, 256 streams,
//                    64 hidden units) depends only on tiles of the same 256 streams (layer below at this hop,
//                    own layer at the previous hop), so the recurrence is a per-(layer, row-pair) counter that
//                    the epilogue warps bump (release) and the TMA producer polls (acquire) -- no grid barrier,
//                    no launch boundary between layers or hops.
// Warp roles per CTA: warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B boxes of 64 k into a 4-stage
// ring; its own 128 activation rows and HALF of the weight rows of the tile), warp 1 (leader CTA) = MMA issuer,
// warp 2 = TMEM allocator, warps 4-11 = epilogue (tcgen05.ld, one stream per thread).  Two accumulator buffers in
// TMEM: the epilogue of tile i overlaps the MMAs of tile i+1.  The issue loops run warp-converged on uniform
// values so that descriptors live in uniform registers (no per-instruction lane election): ~6 SASS instructions
// per tcgen05.mma instead of 23 -- with one thread issuing, that loop was what paced the tensor pipe in round 1.
//
// The GRU tile holds, for 64 hidden units, four accumulators per unit (z, r, W_n x, U_n h: nnet.cpp:136-173
// needs the last two apart), 256 TMEM columns [nx | z | r | nh]; weight rows are packed gate-interleaved per tile
// (input part [n|z|r], recurrent part [z|r|n]) so that one TMA box brings a tile's rows and ONE N = 192 MMA
// per product covers it.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/percepnet_b200.h"
#include "pnb_engine.h"

using namespace pnb;

namespace {

constexpr int TM = 128;   // streams per CTA tile (UMMA M = 256 per pair)
constexpr int BK = 64;    // k per pipeline stage: 128-byte rows, SWIZZLE_128B (64-byte rows made the TMA unit the pacing stage:
                          // one L2 request per row, 448 request cycles per 576 MMA cycles)
constexpr int HT = 64;    // hidden units per GRU tile
constexpr int GRU_BN = 3 * HT;  // weight rows per GRU tile (z|r|n)
// Epilogue warps per CTA: warp 4 + q + 4 k reads TMEM lane quadrant q; the kEpiSets warps of a quadrant share its
// rows' columns 16 at a time.
constexpr int kEpiSets = 2;
constexpr int kEpiWarps = 4 * kEpiSets;
constexpr int kTcThreads = 32 * (4 + kEpiWarps);
constexpr int DENSE_BN = 256;   // output columns per dense (conv) tile: the widest MMA, least operand traffic per MAC
constexpr float kActScale = 1024.f;  // fp16 activations are stored x 2^10 (keeps the low term normal)
// |pre-activation| from which the reference's tansig_approx (vec.h:63) overflows its float->int conversion
constexpr float kTanhDomain = 8.5e7f;

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
// ---- CTA-pair (cta_group::2) forms: the two CTAs of a cluster act as one M = 256 tensor-core unit ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> the leader's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory whose bytes are accounted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(x), "r"(y),
      "l"(0x1000000000000000ull)  // EVICT_NORMAL
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
// D[256 x N] += A[256 x 16] B[N x 16]^T : rows 0..127 from the leader's A tile / TMEM, 128..255 from the peer's;
// each CTA's shared memory supplies N/2 rows of B.  Every MMA accumulates (the epilogue zeroes the buffer).
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
// completion of all MMAs issued so far arrives on `bar` in BOTH CTAs
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 16 consecutive accumulator columns of this thread's TMEM lane; the values are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// Ties 16 loaded registers to the point after tmem_ld_wait(): the empty volatile asm "produces" them, so no use of
// the values can be scheduled ahead of the wait (the wait itself names no registers).
__device__ __forceinline__ void tmem_ld_fence16(uint32_t *r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                    "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO [16,30), SBO [32,46), version=1 [46,48), layout [61,64) with 2 = SW128).
// The high word is a constant; the low word is (address >> 4) | LBO and advances by plain addition, 32 bytes per
// 16-wide k-step inside the 128-byte row (the shared window is < 256 KB, so the 14-bit field never carries).
// Tiles are 1024-byte aligned (the swizzle is a function of the address bits).
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFF) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A/B format (0 fp16, 1 bf16), both K-major
__host__ __device__ constexpr uint32_t idesc_f16(int n, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);  // M = 256: the CTA pair
}

// ------------------------------------------------------------------------------------------ activations
// tansig_approx of vec.h:53-71 on the FMA/ALU pipes only.  The reference forms the table index with
// floor(.5f + 25 x) -> (int); here t + 2^23 rounded toward -inf leaves floor(t) in the low mantissa bits (exact for
// 0 <= t < 2^23), which replaces FRND + F2I + I2F (quarter-rate conversion pipe) by one FADD.RM and one FADD.
// |x| is clamped at 8: from there on the reference returns exactly +-1 (index clamped at 200, where dy = 0).
// NaN propagates (min.NaN); beyond |x| = 8.6e7 the reference's conversion overflows (undefined behaviour, x86
// returns its argument) -- callers whose pre-activations can get there raise the engine's domain flag instead.
__device__ __forceinline__ float tansig_s(float x, const float *tbl) {
  float ax;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(ax) : "f"(fabsf(x)), "f"(8.0f));
  const float t = fmaf(25.f, ax, .5f);
  const float r = __fadd_rd(t, 8388608.f);
  const int idx = __float_as_int(r) & 0xFF;           // floor(t) <= 200 (255 at most for a NaN payload; table has 256 slots)
  const float fi = r - 8388608.f;
  const float xr = fmaf(-.04f, fi, ax);
  const float y = tbl[idx];
  const float dy = fmaf(-y, y, 1.f);
  const float u = fmaf(-y, xr, 1.f);
  const float v = fmaf(xr * dy, u, y);
  return __int_as_float(__float_as_int(v) | (__float_as_int(x) & 0x80000000));
}
__device__ __forceinline__ float sigmoid_s(float x, const float *tbl) { return fmaf(.5f, tansig_s(.5f * x, tbl), .5f); }

// fp16 two-term split of 16 values (x 2^10), written as two 32-byte rows
__device__ __forceinline__ void store_split_h16(const float *v, __half *hi_p, __half *lo_p) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float s0 = v[2 * i] * kActScale, s1 = v[2 * i + 1] * kActScale;
    const __half2 h = __floats2half2_rn(s0, s1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(s0 - hf.x, s1 - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t *>(&h);
    lo[i] = *reinterpret_cast<const uint32_t *>(&l);
  }
  uint4 *ph = reinterpret_cast<uint4 *>(hi_p), *pl = reinterpret_cast<uint4 *>(lo_p);
  ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}
// bf16 split of 16 values into `terms` (2 or 3) planes `plane` elements apart
__device__ __forceinline__ void store_split_b16(const float *v, __nv_bfloat16 *t0_p, size_t plane, int terms) {
  uint32_t a[8], b[8], c3[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h);
    const float r0 = v[2 * i] - hf.x, r1 = v[2 * i + 1] - hf.y;
    const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
    const float2 lf = __bfloat1622float2(l);
    const __nv_bfloat162 m = __floats2bfloat162_rn(r0 - lf.x, r1 - lf.y);
    a[i] = *reinterpret_cast<const uint32_t *>(&h);
    b[i] = *reinterpret_cast<const uint32_t *>(&l);
    c3[i] = *reinterpret_cast<const uint32_t *>(&m);
  }
  uint4 *p0 = reinterpret_cast<uint4 *>(t0_p), *p1 = reinterpret_cast<uint4 *>(t0_p + plane);
  p0[0] = make_uint4(a[0], a[1], a[2], a[3]); p0[1] = make_uint4(a[4], a[5], a[6], a[7]);
  p1[0] = make_uint4(b[0], b[1], b[2], b[3]); p1[1] = make_uint4(b[4], b[5], b[6], b[7]);
  if (terms == 3) {
    uint4 *p2 = reinterpret_cast<uint4 *>(t0_p + 2 * plane);
    p2[0] = make_uint4(c3[0], c3[1], c3[2], c3[3]); p2[1] = make_uint4(c3[4], c3[5], c3[6], c3[7]);
  }
}

// ------------------------------------------------------------------------------------------ shared kernel pieces
template <int BN, int NT = 2>
struct StageLayout {  // NT terms of A (this CTA's 128 rows), then NT terms of this CTA's half of the B rows
  static constexpr int kABytes = TM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;
  static constexpr int kBytes = NT * (kABytes + kBBytes);
};

struct Pipe {  // position in an mbarrier ring
  uint32_t st = 0, ph = 0;
  __device__ __forceinline__ void advance(uint32_t stages) {
    if (++st == stages) { st = 0; ph ^= 1; }
  }
};

struct SmemCarve {
  uint8_t *stages;
  uint32_t full, empty, acc_full, acc_empty;  // shared-window addresses of the barrier arrays
  uint32_t *tmem_slot;
  float *tbl;
};
template <int STAGES, int REGION_BYTES>  // REGION_BYTES: the stage ring (the barriers sit behind it)
__device__ __forceinline__ SmemCarve carve(uint8_t *smem_raw) {
  // the swizzled tiles need 1024-byte (SW128) alignment; align the carve-up by hand
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  SmemCarve c;
  c.stages = smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + REGION_BYTES);
  c.full = smem_u32(bars);
  c.empty = c.full + 8 * STAGES;
  c.acc_full = c.empty + 8 * STAGES;
  c.acc_empty = c.acc_full + 16;
  c.tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * STAGES + 4);
  c.tbl = reinterpret_cast<float *>(c.tmem_slot + 4);
  return c;
}
template <int STAGES, int REGION_BYTES>
constexpr size_t tc_smem_bytes() {
  return (size_t)REGION_BYTES + (2 * STAGES + 4) * 8 + 16 + 256 * 4 + 1024;
}
template <int BN, int ST2, int ST3>
constexpr size_t dense_smem_bytes() {
  constexpr int r2 = ST2 * StageLayout<BN, 2>::kBytes, r3 = ST3 * StageLayout<BN, 3>::kBytes;
  return tc_smem_bytes<(ST2 > ST3 ? ST2 : ST3), (r2 > r3 ? r2 : r3)>();
}

// common prologue: barriers, TMEM allocation (2 accumulator buffers), activation table, zeroed accumulators
template <int STAGES, int TMEM_COLS>
__device__ __forceinline__ uint32_t tc_prologue(const SmemCarve &c, const float *tansig, int warp, int lane) {
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; i++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.full + 8 * i));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.empty + 8 * i));
    }
    for (int i = 0; i < 2; i++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.acc_full + 8 * i));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c.acc_empty + 8 * i), "r"(2 * 32 * kEpiWarps));  // every epilogue thread of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(c.tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) c.tbl[i] = i < 201 ? tansig[i] : 1.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *c.tmem_slot;
  if (warp >= 4) {  // zero both accumulator buffers: every MMA accumulates
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int col = 0; col < TMEM_COLS; col += 16) tmem_st16_zero(tl + col);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM are ready before the leader issues anything
  tc_fence_after();
  return tmem_base;
}
template <int TMEM_COLS>
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base, int warp) {
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair still has work in flight
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// One k-block of a tile on the MMA side: the significant products of both 16-wide k-steps, then the stage is released.
// Runs warp-converged; `a_lo` is the descriptor low word of the stage's first A term (B terms follow the A terms).
// Two terms per operand: hi hi, hi lo, lo hi.  Three terms (t0 t1 t2, each 8 bits below the previous): the six
// products of total order <= 2.
// PASS < 0: all products of the k-block at once (GRU chain: operands bounded by 1, sums of a few units).
// PASS >= 0 (dense layers, whose ReLU inputs are unbounded): the K loop runs once per product order, smallest products
// first (three terms: PASS 0 order 2, PASS 1 order 1, PASS 2 t0 t0; two terms: PASS 1, PASS 2).  The tensor core adds
// into the fp32 accumulator with truncation at the accumulator's magnitude, so correction products added to an already
// large sum lose exactly the bits they were meant to supply; summed among themselves first, they do not.
template <int BN, int NT, int PASS = -1>
__device__ __forceinline__ void mma_kblock(uint32_t dcol, uint32_t a_lo, uint32_t idesc, uint32_t empty_bar) {
  using SL = StageLayout<BN, NT>;
  constexpr uint32_t A = SL::kABytes >> 4, B = SL::kBBytes >> 4;
  const uint32_t b_lo = a_lo + NT * A;
  if (elect_one()) {
#pragma unroll
    for (int ks = 0; ks < BK / 16; ks++) {
      const uint32_t a = a_lo + ks * 2, b = b_lo + ks * 2;
      if (NT == 2 && PASS < 0) {
        umma_f16(dcol, desc64(a), desc64(b), idesc);
        umma_f16(dcol, desc64(a), desc64(b + B), idesc);
        umma_f16(dcol, desc64(a + A), desc64(b), idesc);
      } else if (NT == 2 && PASS == 1) {
        umma_f16(dcol, desc64(a), desc64(b + B), idesc);
        umma_f16(dcol, desc64(a + A), desc64(b), idesc);
      } else if (NT == 2) {
        umma_f16(dcol, desc64(a), desc64(b), idesc);
      } else if (PASS == 0) {
        umma_f16(dcol, desc64(a + A), desc64(b + B), idesc);
        umma_f16(dcol, desc64(a), desc64(b + 2 * B), idesc);
        umma_f16(dcol, desc64(a + 2 * A), desc64(b), idesc);
      } else if (PASS == 1) {
        umma_f16(dcol, desc64(a), desc64(b + B), idesc);
        umma_f16(dcol, desc64(a + A), desc64(b), idesc);
      } else {
        umma_f16(dcol, desc64(a), desc64(b), idesc);
      }
    }
    umma_commit(empty_bar);  // frees the stage in both CTAs once these MMAs have read it
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------ dense layers
struct TcSeg {
  int a_map, b_map;  // indices into TcArgs::maps
  int k_blocks;      // K / 32
  int a_k0, b_k0;    // starting k (elements) in the A / B matrices
  int a_row0;        // row of the A matrix that holds tile row 0 (hop / tap offset inside a multi-hop buffer)
  int a_term_rows;   // row distance between the split terms inside this A buffer
};

struct TcArgs {
  CUtensorMap maps[8];
  TcSeg seg[5];
  int n_seg;
  int M;                // rows
  int tiles_m, tiles_n; // tile grid; a CTA pair walks pair-tiles pair_id, pair_id + n_pairs, ...
  int b_term_rows;      // row distance between the split terms inside a packed weight matrix
  int fmt;              // 0 fp16, 1 bf16
  float out_scale;      // 2^-(activation scale + weight scale)
  const float *tansig;  // 201-entry table (global)
  const float *bias;    // padded to a multiple of 16 entries
  int *status;          // engine status word: bit 0 = an activation left the reference's defined domain
  const int *wide;      // null, or a device word: non-zero = run with three operand terms (six products) instead of two
  int out_terms;        // bf16 terms written to out_b (2 or 3)
  int act;
  int N;                // valid output columns
  int ldc;              // row stride of out_f32
  float *out_f32;       // [M][ldc] or null
  int f32_row0;         // only rows >= f32_row0 are written to out_f32, at row - f32_row0
  __half *out_h;        // fp16 two-term split [2][out_plane_rows][N] (x 2^10) or null
  __nv_bfloat16 *out_b; // bf16 two-term split [2][out_plane_rows][N] or null
  int out_row0;         // first row written inside a term plane of out_h / out_b
  int out_plane_rows;   // rows per term plane of out_h / out_b
};

__host__ __device__ constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }

// ===== TMA producer (both CTAs): own A rows, own half of the B rows; bytes counted on the leader's barrier =====
template <int BN, int NT, int STAGES>
__device__ __forceinline__ void dense_producer(const TcArgs &args, const SmemCarve &c, uint32_t stage0, int crank, int pair_id,
                                               int n_pairs_cta, int total_pairs) {
  using SL = StageLayout<BN, NT>;
  Pipe p;
  for (int pair = pair_id; pair < total_pairs; pair += n_pairs_cta) {
    const int n_tile = pair % args.tiles_n, m0 = (2 * (pair / args.tiles_n) + crank) * TM;
    // one pass over K per product order (see mma_kblock); with q passes still to go after this one, the pass needs
    // the first q + 1 terms of either operand
    for (int pass = 0; pass < NT; pass++) {
      const int nt = NT - pass;
      for (int s = 0; s < args.n_seg; s++) {
        const TcSeg &sg = args.seg[s];
        const CUtensorMap *ma = &args.maps[sg.a_map], *mb = &args.maps[sg.b_map];
        const int arow = sg.a_row0 + m0, brow = n_tile * BN + crank * (BN / 2);
        for (int kb = 0; kb < sg.k_blocks; kb++) {
          mbar_wait(c.empty + 8 * p.st, p.ph ^ 1);  // fresh barrier: passes immediately
          if (elect_one()) {
            const uint32_t fb = c.full + 8 * p.st;
            if (crank == 0) mbar_expect_tx(fb, 2 * nt * (SL::kABytes + SL::kBBytes));
            const uint32_t sp = stage0 + p.st * SL::kBytes;
#pragma unroll
            for (int tm = 0; tm < NT; tm++) {
              if (tm < nt) {
                tma_load_2d_pair(sp + tm * SL::kABytes, ma, fb, sg.a_k0 + kb * BK, tm * sg.a_term_rows + arow);
                tma_load_2d_pair(sp + NT * SL::kABytes + tm * SL::kBBytes, mb, fb, sg.b_k0 + kb * BK, tm * args.b_term_rows + brow);
              }
            }
          }
          __syncwarp();
          p.advance(STAGES);
        }
      }
    }
  }
}
// ===== MMA issuer: the leader CTA's warp 1 drives both SMs' tensor cores =====
template <int BN, int NT, int STAGES, int ACC_COLS>
__device__ __forceinline__ void dense_mma(const TcArgs &args, const SmemCarve &c, uint32_t stage0, uint32_t tmem_base, int pair_id,
                                          int n_pairs_cta, int total_pairs) {
  using SL = StageLayout<BN, NT>;
  const uint32_t idesc = idesc_f16(BN, args.fmt);
  const uint32_t a_lo0 = desc_lo(stage0);
  Pipe p;
  int j = 0;
  for (int pair = pair_id; pair < total_pairs; pair += n_pairs_cta, j++) {
    const int buf = j & 1;
    mbar_wait(c.acc_empty + 8 * buf, ((j >> 1) & 1) ^ 1);  // both epilogues drained (and re-zeroed) this buffer
    tc_fence_after();
    const uint32_t acc = tmem_base + buf * ACC_COLS;
    int nkb = 0;
    for (int s = 0; s < args.n_seg; s++) nkb += args.seg[s].k_blocks;
    if (NT == 3) {
      for (int kb = 0; kb < nkb; kb++) {
        mbar_wait(c.full + 8 * p.st, p.ph);
        tc_fence_after();
        mma_kblock<BN, NT, 0>(acc, a_lo0 + p.st * (SL::kBytes >> 4), idesc, c.empty + 8 * p.st);
        p.advance(STAGES);
      }
    }
    for (int kb = 0; kb < nkb; kb++) {
      mbar_wait(c.full + 8 * p.st, p.ph);
      tc_fence_after();
      mma_kblock<BN, NT, 1>(acc, a_lo0 + p.st * (SL::kBytes >> 4), idesc, c.empty + 8 * p.st);
      p.advance(STAGES);
    }
    for (int kb = 0; kb < nkb; kb++) {
      mbar_wait(c.full + 8 * p.st, p.ph);
      tc_fence_after();
      mma_kblock<BN, NT, 2>(acc, a_lo0 + p.st * (SL::kBytes >> 4), idesc, c.empty + 8 * p.st);
      p.advance(STAGES);
    }
    if (elect_one()) umma_commit(c.acc_full + 8 * buf);  // both epilogues
    __syncwarp();
  }
}

// ST2 stages of two-term operands; ST3 > 0 adds a three-term mode (ST3 stages) chosen at run time from *args.wide
template <int BN, int ST2, int ST3>
__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_kernel(const __grid_constant__ TcArgs args) {
  constexpr int kStagesMax = ST2 > ST3 ? ST2 : ST3;
  constexpr int kRegion = (ST2 * StageLayout<BN, 2>::kBytes > ST3 * StageLayout<BN, 3>::kBytes) ? ST2 * StageLayout<BN, 2>::kBytes
                                                                                               : ST3 * StageLayout<BN, 3>::kBytes;
  constexpr int kAccCols = pow2_cols(BN);  // TMEM columns per accumulator buffer
  constexpr int kTmemCols = 2 * kAccCols;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemCarve c = carve<kStagesMax, kRegion>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();  // 0 = leader
  const int n_pairs_cta = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const int total_pairs = ((args.tiles_m + 1) >> 1) * args.tiles_n;  // work unit: one B tile x two row tiles
  const uint32_t tmem_base = tc_prologue<kStagesMax, kTmemCols>(c, args.tansig, warp, lane);
  const uint32_t stage0 = smem_u32(c.stages);
  // the same word for every thread of the launch (written by an earlier launch): both CTAs take the same branch
  const bool wide = ST3 > 0 && args.wide != nullptr && *args.wide != 0;

  if (warp == 0) {
    if (ST3 > 0 && wide) dense_producer<BN, 3, ST3 ? ST3 : 1>(args, c, stage0, crank, pair_id, n_pairs_cta, total_pairs);
    else dense_producer<BN, 2, ST2>(args, c, stage0, crank, pair_id, n_pairs_cta, total_pairs);
  } else if (warp == 1) {
    if (crank == 0) {
      if (ST3 > 0 && wide) dense_mma<BN, 3, ST3 ? ST3 : 1, kAccCols>(args, c, stage0, tmem_base, pair_id, n_pairs_cta, total_pairs);
      else dense_mma<BN, 2, ST2, kAccCols>(args, c, stage0, tmem_base, pair_id, n_pairs_cta, total_pairs);
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31, one row per thread =====
    const int wq = warp & 3, eset = (warp - 4) >> 2;
    const float sc = args.out_scale;
    const float *tbl = c.tbl;
    const int N = args.N, act = args.act;
    int j = 0;
    for (int pair = pair_id; pair < total_pairs; pair += n_pairs_cta, j++) {
      const int n_tile = pair % args.tiles_n, m0 = (2 * (pair / args.tiles_n) + crank) * TM;
      const int buf = j & 1;
      mbar_wait(c.acc_full + 8 * buf, (j >> 1) & 1);
      tc_fence_after();
      const int row = m0 + wq * 32 + lane;
      const bool row_ok = row < args.M;
      const uint32_t tlane = tmem_base + buf * kAccCols + ((uint32_t)(wq * 32) << 16);
      for (int col = 16 * eset; col < BN; col += 16 * kEpiSets) {
        uint32_t d[16];
        tmem_ld16_nowait(tlane + col, d);
        const int j0 = n_tile * BN + col;
        float4 bq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) bq[q] = __ldg(reinterpret_cast<const float4 *>(args.bias + j0) + q);
        tmem_ld_wait();
        tmem_ld_fence16(d);
        if (row_ok && j0 < N) {
          const float *bs = reinterpret_cast<const float *>(bq);
          float v[16];
          bool bad = false;
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float x = fmaf(__uint_as_float(d[i]), sc, bs[i]);
            if (act == PNB_ACT_TANH) { v[i] = tansig_s(x, tbl); bad |= fabsf(x) >= kTanhDomain; }
            else if (act == PNB_ACT_RELU) v[i] = x < 0.f ? 0.f : x;
            else if (act == PNB_ACT_SIGMOID) { v[i] = sigmoid_s(x, tbl); bad |= fabsf(x) >= 2.f * kTanhDomain; }
            else v[i] = x;
          }
          if (bad) atomicOr(args.status, 1);
          if (args.out_f32 && row >= args.f32_row0) {
            float *op = args.out_f32 + (size_t)(row - args.f32_row0) * args.ldc + j0;
            if ((reinterpret_cast<uintptr_t>(op) & 15) == 0 && j0 + 16 <= N) {
              float4 *o = reinterpret_cast<float4 *>(op);
#pragma unroll
              for (int q = 0; q < 4; q++) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; i++)
                if (j0 + i < N) op[i] = v[i];
            }
          }
          const size_t orow = (size_t)args.out_row0 + row;
          if (args.out_h)
            store_split_h16(v, args.out_h + orow * N + j0, args.out_h + ((size_t)args.out_plane_rows + orow) * N + j0);
          if (args.out_b)
            store_split_b16(v, args.out_b + orow * N + j0, (size_t)args.out_plane_rows * N, args.out_terms);
        }
        tmem_st16_zero(tlane + col);  // this thread's columns are read: zero them for the buffer's next tile
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_leader(c.acc_empty + 8 * buf);
    }
  }
  tc_teardown<kTmemCols>(tmem_base, warp);
}

// ------------------------------------------------------------------------------------------ the GRU chain
struct ChainLayer {
  int n_x;              // input segments (2 for gru_rb: [gru3 state, conv2 out])
  int x_map[2];         // A tensor maps of the input segments
  int x_kb[2];          // their k-blocks
  int x_bk0[2];         // starting k in the packed input weights
  int x_row1[2];        // 0: the segment is indexed by hop (conv2 out, slot t); 1: state slots (slot t+1 = after hop t)
  int x_term_rows[2];   // row distance between the two split terms in that buffer
  int h_map, w_map, u_map;
  int h_kb;             // H / 32
  int H;
  int tiles_log2;       // log2(H / 64)
  int b_term_rows;      // rows of one term plane of the packed weights
  int dep;              // layer whose output at the same hop is this layer's input (-1: produced before the launch)
  int unit0;            // first unit of this layer inside a hop
  int par0;             // which fp32 state buffer holds the state before hop 0
  float scale;          // 2^-(10 + weight scale)
  const float4 *bias4;  // per unit {b_z + b'_z, b_r + b'_r, b_n, b'_n}
  float *h32[2];        // fp32 state, ping-pong by hop
  __half *h_hl;         // fp16 terms [2][(Fmax+1) S][H]
  int h_term_rows;      // (Fmax+1) S
};
struct ChainArgs {
  CUtensorMap maps[16];  // 0: conv2 out, 1..5: state slots of the five GRUs, 6..10: input weights, 11..15: recurrent weights
  ChainLayer L[5];
  int S, F;              // streams; hops this launch covers
  int t0;                // first hop of this launch inside the call (slot / counter arithmetic is per call)
  int tiles_mp;          // row pairs (256 streams each)
  int units_per_hop;
  int diag;              // unit order: 0 = hop-major (layer by layer inside a hop), 1 = anti-diagonals (few streams)
  unsigned *cnt;         // [5][tiles_mp] epilogue-warp completions, zero at launch
  const float *tansig;
};

struct Unit { int t, l, mp, nt; };
// Anti-diagonal unit order (small batches): (hop t, layer l) sits on diagonal t + depth(l), depth = 0,1,2,3,3 (gru_rb reads
// gru3 like gru_gb does).  Everything a unit waits for lies on an earlier diagonal, and a diagonal holds up to five
// layer-hops of mutually independent units -- with few streams one layer-hop alone cannot occupy the CTA pairs.
__device__ __forceinline__ int diag_size(const ChainArgs &a, int d) {
  int n = 0;
#pragma unroll
  for (int l = 0; l < 5; l++) {
    const int t = d - (l < 4 ? l : 3);
    n += (t >= 0 && t < a.F) ? (a.tiles_mp << a.L[l].tiles_log2) : 0;
  }
  return n;
}
__device__ __forceinline__ Unit decode_unit_diag(const ChainArgs &a, int u) {
  int d = 0, r = u;
  bool found = false;
  for (; d < 3 && !found; d++) {  // ramp-up diagonals
    const int sz = diag_size(a, d);
    if (r < sz) { found = true; break; }
    r -= sz;
  }
  if (!found && a.F > 3) {        // full diagonals 3 .. F-1 all hold units_per_hop units
    const int q = r / a.units_per_hop;
    if (q < a.F - 3) { d = 3 + q; r -= q * a.units_per_hop; found = true; }
    else { r -= (a.F - 3) * a.units_per_hop; d = a.F; }
  }
  for (; !found; d++) {           // ramp-down diagonals
    const int sz = diag_size(a, d);
    if (r < sz) break;
    r -= sz;
  }
  Unit x;
  x.l = 0; x.t = 0;
  bool hit = false;
#pragma unroll
  for (int l = 0; l < 5; l++) {
    const int t = d - (l < 4 ? l : 3);
    const int nl = (t >= 0 && t < a.F) ? (a.tiles_mp << a.L[l].tiles_log2) : 0;
    if (!hit) {
      if (r < nl) { x.l = l; x.t = t; hit = true; }
      else r -= nl;
    }
  }
  const int sh = a.L[x.l].tiles_log2;
  x.t += a.t0;
  x.mp = r >> sh;
  x.nt = r & ((1 << sh) - 1);
  return x;
}
__device__ __forceinline__ Unit decode_unit(const ChainArgs &a, int u) {
  if (a.diag) return decode_unit_diag(a, u);
  Unit x;
  const int tl = u / a.units_per_hop;
  int r = u - tl * a.units_per_hop;
  x.t = a.t0 + tl;
  x.l = 0;
#pragma unroll
  for (int i = 1; i < 5; i++) x.l = (r >= a.L[i].unit0) ? i : x.l;
  r -= a.L[x.l].unit0;
  const int sh = a.L[x.l].tiles_log2;
  x.mp = r >> sh;
  x.nt = r & ((1 << sh) - 1);
  return x;
}
// all epilogue warps of both CTAs of every tile of (layer, row pair) have finished `hops` hops
__device__ __forceinline__ void wait_layer(const ChainArgs &a, int layer, int mp, int hops) {
  const unsigned target = (unsigned)hops * (2u * kEpiWarps) << a.L[layer].tiles_log2;
  const unsigned *p = a.cnt + layer * a.tiles_mp + mp;
  unsigned spins = 0;
  while (ld_acquire_gpu(p) < target)
    if (++spins > (1u << 26)) __trap();  // tens of seconds: a broken schedule becomes a launch error, not a hang
}

constexpr int GRU_STAGES = 4;

// Per-role cycle accounting of the chain kernel (debug builds only: -DPNB_CHAIN_TIMING, tools/chain_roles.py).
#ifdef PNB_CHAIN_TIMING
__device__ unsigned long long g_chain_cycles[16];
#define CT_DECL long long ct_t = clock64(); unsigned long long ct_acc[6] = {0, 0, 0, 0, 0, 0}
#define CT_TICK(k) do { long long _n = clock64(); ct_acc[k] += (unsigned long long)(_n - ct_t); ct_t = _n; } while (0)
#define CT_FLUSH(base, n) do { for (int _i = 0; _i < (n); _i++) atomicAdd(&g_chain_cycles[(base) + _i], ct_acc[_i]); } while (0)
#else
#define CT_DECL do { } while (0)
#define CT_TICK(k) do { } while (0)
#define CT_FLUSH(base, n) do { } while (0)
#endif

__global__ void __launch_bounds__(kTcThreads, 1) gru_chain_kernel(const __grid_constant__ ChainArgs args) {
  using SL = StageLayout<GRU_BN>;
  constexpr int STAGES = GRU_STAGES;
  constexpr int kAccCols = 4 * HT, kTmemCols = 2 * kAccCols;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const SmemCarve c = carve<STAGES, STAGES * SL::kBytes>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int n_pairs_cta = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const int total_units = args.F * args.units_per_hop;
  const int S = args.S;
  const uint32_t tmem_base = tc_prologue<STAGES, kTmemCols>(c, args.tansig, warp, lane);
  const uint32_t stage0 = smem_u32(c.stages);

  if (warp == 0) {
    // ===== TMA producer: waits for the rows it is about to read, then streams the tile's operands =====
    Pipe p;
    CT_DECL;
    for (int u = pair_id; u < total_units; u += n_pairs_cta) {
      const Unit un = decode_unit(args, u);
      const ChainLayer &L = args.L[un.l];
      CT_TICK(0);  // 0: decode + issue
      if (elect_one()) {  // the same lane issues the TMA loads below
        if (L.dep >= 0) wait_layer(args, L.dep, un.mp, un.t + 1);  // input: layer below, this hop
        if (un.t > 0) wait_layer(args, un.l, un.mp, un.t);        // own state after the previous hop
        fence_proxy_async();  // the rows were written with st.global by other SMs; TMA reads them through the async proxy
      }
      __syncwarp();
      CT_TICK(1);  // 1: dependency wait
      const int m0 = (2 * un.mp + crank) * TM;
      const int brow = un.nt * GRU_BN + crank * (GRU_BN / 2);
      for (int s = 0; s <= L.n_x; s++) {
        const bool rec = s == L.n_x;
        const CUtensorMap *ma = &args.maps[rec ? L.h_map : L.x_map[s]], *mb = &args.maps[rec ? L.u_map : L.w_map];
        const int nkb = rec ? L.h_kb : L.x_kb[s];
        const int arow = (rec ? un.t : un.t + L.x_row1[s]) * S + m0;
        const int aterm = rec ? L.h_term_rows : L.x_term_rows[s];
        const int bk0 = rec ? 0 : L.x_bk0[s];
        for (int kb = 0; kb < nkb; kb++) {
          CT_TICK(0);
          mbar_wait(c.empty + 8 * p.st, p.ph ^ 1);
          CT_TICK(2);  // 2: waiting for a free stage
          if (elect_one()) {
            const uint32_t fb = c.full + 8 * p.st;
            if (crank == 0) mbar_expect_tx(fb, 2 * SL::kBytes);
            const uint32_t sp = stage0 + p.st * SL::kBytes;
            tma_load_2d_pair(sp, ma, fb, kb * BK, arow);
            tma_load_2d_pair(sp + SL::kABytes, ma, fb, kb * BK, aterm + arow);
            tma_load_2d_pair(sp + 2 * SL::kABytes, mb, fb, bk0 + kb * BK, brow);
            tma_load_2d_pair(sp + 2 * SL::kABytes + SL::kBBytes, mb, fb, bk0 + kb * BK, L.b_term_rows + brow);
          }
          __syncwarp();
          p.advance(STAGES);
        }
      }
    }
    CT_TICK(0);
    if (lane == 0 && crank == 0) CT_FLUSH(0, 3);
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (crank == 0) {
      const uint32_t idesc = idesc_f16(GRU_BN, 0);
      const uint32_t a_lo0 = desc_lo(stage0);
      Pipe p;
      int j = 0;
      CT_DECL;
      for (int u = pair_id; u < total_units; u += n_pairs_cta, j++) {
        const Unit un = decode_unit(args, u);
        const ChainLayer &L = args.L[un.l];
        const int buf = j & 1;
        CT_TICK(0);  // 0: decode + issue
        mbar_wait(c.acc_empty + 8 * buf, ((j >> 1) & 1) ^ 1);
        CT_TICK(1);  // 1: waiting for the epilogue to hand the accumulator back
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * kAccCols;
        // accumulator columns [nx | z | r | nh]: input part (rows [n|z|r]) at column 0, recurrent part (rows [z|r|n]) at 64
        int nkb = 0;
        for (int s = 0; s < L.n_x; s++) nkb += L.x_kb[s];
        for (int kb = 0; kb < nkb; kb++) {
          CT_TICK(0);
          mbar_wait(c.full + 8 * p.st, p.ph);
          CT_TICK(2);  // 2: waiting for operands
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + p.st * (SL::kBytes >> 4);
          mma_kblock<GRU_BN, 2>(acc, a_lo, idesc, c.empty + 8 * p.st);
          p.advance(STAGES);
        }
        for (int kb = 0; kb < L.h_kb; kb++) {
          CT_TICK(0);
          mbar_wait(c.full + 8 * p.st, p.ph);
          CT_TICK(2);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + p.st * (SL::kBytes >> 4);
          mma_kblock<GRU_BN, 2>(acc + HT, a_lo, idesc, c.empty + 8 * p.st);
          p.advance(STAGES);
        }
        if (elect_one()) umma_commit(c.acc_full + 8 * buf);
        __syncwarp();
      }
      CT_TICK(0);
      if (lane == 0) CT_FLUSH(4, 3);
    }
  } else if (warp >= 4) {
    // ===== epilogue: gate math of nnet.cpp:136-178 (reset_after) on this thread's stream, 16 units at a time =====
    const int wq = warp & 3, eset = (warp - 4) >> 2;
    const float *tbl = c.tbl;
    int j = 0;
    CT_DECL;
    for (int u = pair_id; u < total_units; u += n_pairs_cta, j++) {
      const Unit un = decode_unit(args, u);
      const ChainLayer &L = args.L[un.l];
      const int H = L.H, buf = j & 1;
      CT_TICK(0);  // 0: work
      const int row = (2 * un.mp + crank) * TM + wq * 32 + lane;
      const bool row_ok = row < S;
      const int pold = (L.par0 + un.t) & 1;
      const float *h_old = L.h32[pold] + (size_t)row * H + un.nt * HT;
      float *h_new = L.h32[pold ^ 1] + (size_t)row * H + un.nt * HT;
      // the previous hop's state of these rows is complete (other SMs wrote it): acquire before reading it
      if (un.t > 0) {
        if (lane == 0) wait_layer(args, un.l, un.mp, un.t);
        __syncwarp();
      }
      float4 hold[2][4];
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int q = 0; q < 4; q++)
          hold[b][q] = row_ok ? __ldcg(reinterpret_cast<const float4 *>(h_old + 16 * eset + 32 * b) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      CT_TICK(1);  // 1: dependency wait + state prefetch issue
      mbar_wait(c.acc_full + 8 * buf, (j >> 1) & 1);
      CT_TICK(2);  // 2: waiting for the accumulator
      tc_fence_after();
      const uint32_t tlane = tmem_base + buf * kAccCols + ((uint32_t)(wq * 32) << 16);
      const float sc = L.scale;
      const size_t orow = (size_t)(un.t + 1) * S + row;
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int col = 16 * eset + 32 * b;
        uint32_t zs[16], rs[16], nx[16], nh[16];
        tmem_ld16_nowait(tlane + HT + col, zs);      // accumulator columns: [nx | z | r | nh]
        tmem_ld16_nowait(tlane + 2 * HT + col, rs);
        tmem_ld16_nowait(tlane + col, nx);
        tmem_ld16_nowait(tlane + 3 * HT + col, nh);
        const float4 *b4 = L.bias4 + un.nt * HT + col;
        tmem_ld_wait();
        tmem_ld_fence16(zs); tmem_ld_fence16(rs); tmem_ld_fence16(nx); tmem_ld_fence16(nh);
        // this thread's columns are in registers: zero them for the buffer's next tile
        tmem_st16_zero(tlane + col); tmem_st16_zero(tlane + HT + col);
        tmem_st16_zero(tlane + 2 * HT + col); tmem_st16_zero(tlane + 3 * HT + col);
        if (b == 1) {  // hand the buffer back to the MMA thread before the second half of the gate math
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive_leader(c.acc_empty + 8 * buf);
        }
        const float *ho = reinterpret_cast<const float *>(hold[b]);
        float hn[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float4 bb = __ldg(b4 + i);
          const float z = sigmoid_s(fmaf(__uint_as_float(zs[i]), sc, bb.x), tbl);
          const float r = sigmoid_s(fmaf(__uint_as_float(rs[i]), sc, bb.y), tbl);
          const float tmp = fmaf(__uint_as_float(nh[i]), sc, bb.w);
          float cnd = fmaf(tmp, r, bb.z);
          cnd = fmaf(__uint_as_float(nx[i]), sc, cnd);
          const float n = tansig_s(cnd, tbl);
          hn[i] = fmaf(z, ho[i] - n, n);  // z h + (1 - z) n
        }
        if (row_ok) {
          float4 *hw = reinterpret_cast<float4 *>(h_new + col);
#pragma unroll
          for (int q = 0; q < 4; q++) hw[q] = make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
          __half *ph = L.h_hl + orow * H + un.nt * HT + col;
          store_split_h16(hn, ph, ph + (size_t)L.h_term_rows * H);
        }
      }
      // publish: every store of this warp for this tile is ordered before the counter bump
      CT_TICK(0);
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        red_release_gpu_add(args.cnt + un.l * args.tiles_mp + un.mp, 1u);
      }
      CT_TICK(3);  // 3: publish (fence + counter)
    }
    if (warp == 4 && lane == 0 && crank == 0) CT_FLUSH(8, 4);
  }
  tc_teardown<kTmemCols>(tmem_base, warp);
}

constexpr int DENSE_STAGES = 3, DENSE_STAGES3 = 2, SMALL_BN = 48, SMALL_STAGES = 5;
// conv layers: fc / conv1 outputs and the conv weights are stored as THREE bf16 terms.  The launches use the first
// two (three products, corrections summed before the main product) unless the engine's `wide` word is set
// (PNB_CONV_WIDE at pnb_create: all three terms, six products).  The mode is the same for every row of every call,
// so a stream's result never depends on what its neighbours in the batch carry.
constexpr int kConvTerms = 3;

// fc 70 -> 128 relu in fp32 FMA (0.1 % of the MACs; K = 70 is no tensor-core shape), emitting the three bf16
// terms conv1 consumes.  One block per 16 rows, one thread per output: the 16 feature rows sit in shared memory and are
// read four k at a time (one broadcast 128-bit load per row and four FMAs), the weight column comes through L1.
// Every output is bias + sum over ascending k, one fmaf per term, like the fp32 path's fc kernel.
constexpr int FC_ROWS = 16;
__global__ void __launch_bounds__(128) fc_split_kernel(const float *__restrict__ feat, const float *__restrict__ W,
                                                       const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out,
                                                       int M, size_t plane_rows) {
  __shared__ __align__(16) float f[FC_ROWS][72];
  const int r0 = blockIdx.x * FC_ROWS;
  for (int i = threadIdx.x; i < FC_ROWS * 72; i += blockDim.x) {
    const int rr = i / 72, kk = i - rr * 72;
    f[rr][kk] = (kk < 70 && r0 + rr < M) ? feat[(size_t)(r0 + rr) * 70 + kk] : 0.f;
  }
  __syncthreads();
  const int n = threadIdx.x;
  float a[FC_ROWS];
  const float b0 = bias[n];
#pragma unroll
  for (int rr = 0; rr < FC_ROWS; rr++) a[rr] = b0;
#pragma unroll 1
  for (int k = 0; k < 68; k += 4) {
    const float w0 = __ldg(W + (size_t)k * 128 + n), w1 = __ldg(W + (size_t)(k + 1) * 128 + n);
    const float w2 = __ldg(W + (size_t)(k + 2) * 128 + n), w3 = __ldg(W + (size_t)(k + 3) * 128 + n);
#pragma unroll
    for (int rr = 0; rr < FC_ROWS; rr++) {
      const float4 fv = *reinterpret_cast<const float4 *>(&f[rr][k]);
      a[rr] = fmaf(w0, fv.x, a[rr]);
      a[rr] = fmaf(w1, fv.y, a[rr]);
      a[rr] = fmaf(w2, fv.z, a[rr]);
      a[rr] = fmaf(w3, fv.w, a[rr]);
    }
  }
  {
    const float w0 = __ldg(W + (size_t)68 * 128 + n), w1 = __ldg(W + (size_t)69 * 128 + n);
#pragma unroll
    for (int rr = 0; rr < FC_ROWS; rr++) {
      a[rr] = fmaf(w0, f[rr][68], a[rr]);
      a[rr] = fmaf(w1, f[rr][69], a[rr]);
    }
  }
  const size_t plane = plane_rows * 128;
#pragma unroll
  for (int rr = 0; rr < FC_ROWS; rr++) {
    if (r0 + rr >= M) break;
    float v = a[rr] < 0.f ? 0.f : a[rr];
    __nv_bfloat16 t0 = __float2bfloat16_rn(v);
    float r = v - __bfloat162float(t0);
    __nv_bfloat16 t1 = __float2bfloat16_rn(r);
    __nv_bfloat16 t2 = __float2bfloat16_rn(r - __bfloat162float(t1));
    size_t o = (size_t)(r0 + rr) * 128 + n;
    out[o] = t0; out[plane + o] = t1; out[2 * plane + o] = t2;
  }
}

// End-of-call bookkeeping in one launch: the hop slots the next call reads first (last 4 fc / 2 conv1 outputs,
// last GRU states) move to the front of their buffers, and the chain's dependency counters return to zero.
// A segment shifts `n_slots` slots down by `shift` slots.  One thread owns the same element of every slot and walks
// the slots in ascending order, so a slot that is both source and destination (shift < n_slots) is read before it is
// overwritten -- by the same thread, no cross-thread ordering needed.
struct CarrySeg { uint4 *base; size_t slot16; int n_slots, shift; };
struct CarryArgs {
  CarrySeg seg[16];
  int n_seg;
  unsigned *cnt;
  int n_cnt;
};
__global__ void __launch_bounds__(256) tc_carry_kernel(const __grid_constant__ CarryArgs a) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  for (int s = 0; s < a.n_seg; s++) {
    const CarrySeg sg = a.seg[s];
    for (size_t i = tid; i < sg.slot16; i += nth)
      for (int k = 0; k < sg.n_slots; k++) sg.base[(size_t)k * sg.slot16 + i] = sg.base[(size_t)(k + sg.shift) * sg.slot16 + i];
  }
  for (size_t i = tid; i < (size_t)a.n_cnt; i += nth) a.cnt[i] = 0u;
}

}  // namespace

// ------------------------------------------------------------------------------------------ host state
struct pnb_tc_state {
  // split activations
  // Multi-hop buffers: rows are (hop slot, stream).  The conv layers carry no recurrence, so they run once per
  // call over all F hops (M = F S rows); a tap is the same buffer read `tap` slots later.  Slots 0..3 (0..1) hold
  // the last hops of the previous call.
  __nv_bfloat16 *fc_all = nullptr;   // [3][(Fmax+4) S][128]  fc outputs, bf16 terms
  __nv_bfloat16 *c1_all = nullptr;   // [3][(Fmax+2) S][512]  conv1 outputs
  __half *c2_h = nullptr;            // [2][Fmax S][512]      conv2 outputs, fp16 terms
  __half *h_all[5] = {};             // [2][(Fmax+1) S][H]    GRU states as fp16 terms, slot t+1 = after hop t,
                                     //                       slot 0 = carried from the previous call
  // packed weights
  __nv_bfloat16 *w_conv1 = nullptr, *w_conv2 = nullptr;  // [3][512][K]
  __half *w_gru[5] = {}, *u_gru[5] = {};                 // [2][tiles*192][K]
  float scale_gru[5] = {};                               // 2^-(10 + e)
  float4 *bias4[5] = {};                                 // per unit {b_z + b'_z, b_r + b'_r, b_n, b'_n}
  __half *w_gb = nullptr, *w_rb = nullptr;               // [2][48][2560], [2][48][128] (34 rows used)
  float scale_gb = 0.f, scale_rb = 0.f;
  float *bias_pad = nullptr;                             // conv1 [512] | conv2 [512] | fc_gb [48] | fc_rb [48]
  unsigned *cnt = nullptr;                               // chain dependency counters [5][tiles_mp]
  int *wide = nullptr;                                   // see kConvTerms
  int tiles_mp = 0;
  // tensor maps
  CUtensorMap m_fc_all, m_c1_all, m_c2, m_h[5], m_wconv1, m_wconv2, m_w[5], m_u[5], m_wgb, m_wrb;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::string g_tc_err;

static int make_map(CUtensorMap *m, const void *base, bool bf16, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                        const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

static float pow2_scale_for(const float *a, size_t na, const float *b, size_t nb, int *e_out) {
  float mx = 0.f;
  for (size_t i = 0; i < na; i++) mx = fmaxf(mx, fabsf(a[i]));
  for (size_t i = 0; i < nb; i++) mx = fmaxf(mx, fabsf(b[i]));
  int e = 0;
  if (mx > 0.f && isfinite(mx)) {
    int ex;
    frexpf(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)
    e = 14 - ex;      // mx * 2^e in [2^13, 2^14)
  }
  *e_out = e;
  return ldexpf(1.f, e);
}

extern int tc_fail(int code, const char *msg);  // sets pnb_last_error (pnb_engine.cu)

#define TCK(call)                                                                          \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      g_tc_err = std::string(#call) + ": " + cudaGetErrorString(_e);                       \
      return tc_fail(PNB_ERR_CUDA, g_tc_err.c_str());                                      \
    }                                                                                      \
  } while (0)

// GRU weights: reference layout W[j*3H + g*H + i] -> K-major rows ordered [tile][gate][64 units], scaled, split
// `order` lists the reference gate index (0 z, 1 r, 2 n) of each 64-row group of a tile: input weights are packed
// [n | z | r], recurrent weights [z | r | n], so that either part is ONE N = 192 MMA onto the accumulator columns
// [nx | z | r | nh] (at column 0 and at column 64 respectively).
static void pack_gru(const float *W, int K, int H, float scale, const int order[3], std::vector<__half> &out) {
  const int tiles = H / HT, rows = tiles * GRU_BN;
  out.assign((size_t)2 * rows * K, __float2half(0.f));
  for (int tl = 0; tl < tiles; tl++)
    for (int gq = 0; gq < 3; gq++)
      for (int ii = 0; ii < HT; ii++) {
        const int g = order[gq];
        const int row = tl * GRU_BN + gq * HT + ii, i = tl * HT + ii;
        for (int j = 0; j < K; j++) {
          float w = W[(size_t)j * 3 * H + g * H + i] * scale;
          __half hi = __float2half_rn(w);
          __half lo = __float2half_rn(w - __half2float(hi));
          out[(size_t)row * K + j] = hi;
          out[((size_t)rows + row) * K + j] = lo;
        }
      }
}
// dense/conv weights: W[k*N + n] -> [n][k], three bf16 terms
static void pack_dense_b3(const float *W, int K, int N, std::vector<__nv_bfloat16> &out) {
  out.assign((size_t)3 * N * K, __float2bfloat16(0.f));
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) {
      float w = W[(size_t)k * N + n];
      __nv_bfloat16 t0 = __float2bfloat16_rn(w);
      float r = w - __bfloat162float(t0);
      __nv_bfloat16 t1 = __float2bfloat16_rn(r);
      out[(size_t)n * K + k] = t0;
      out[((size_t)N + n) * K + k] = t1;
      out[((size_t)2 * N + n) * K + k] = __float2bfloat16_rn(r - __bfloat162float(t1));
    }
}

// 34-wide output layers: W[k*34 + n] -> [48 rows (34 used)][K], scaled, two fp16 terms
static void pack_small_h2(const float *W, int K, int N, float scale, std::vector<__half> &out) {
  out.assign((size_t)2 * SMALL_BN * K, __float2half(0.f));
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) {
      float w = W[(size_t)k * N + n] * scale;
      __half hi = __float2half_rn(w);
      out[(size_t)n * K + k] = hi;
      out[((size_t)SMALL_BN + n) * K + k] = __float2half_rn(w - __half2float(hi));
    }
}

// Largest |pre-activation| a layer can form from inputs bounded by 1 (tanh outputs / GRU states): the row sums of
// |W| (+ |U|) plus |b|.  The tensor path is validated inside the reference's defined domain only (tansig_approx
// overflows its float->int conversion from 8.6e7 on, vec.h:63); weights that could leave it are refused up front.
static double max_preact(const float *W, int K, int N, int ld, const float *U, int KU, const float *b0, const float *b1) {
  std::vector<double> acc(N, 0.0);
  for (int k = 0; k < K; k++)
    for (int n = 0; n < N; n++) acc[n] += fabs((double)W[(size_t)k * ld + n]);
  if (U)
    for (int k = 0; k < KU; k++)
      for (int n = 0; n < N; n++) acc[n] += fabs((double)U[(size_t)k * ld + n]);
  double mx = 0.0;
  for (int n = 0; n < N; n++) {
    double v = acc[n] + (b0 ? fabs((double)b0[n]) : 0.0) + (b1 ? fabs((double)b1[n]) : 0.0);
    if (!(v <= mx)) mx = v;  // NaN-propagating max
  }
  return mx;
}

template <typename T>
static cudaError_t dev_upload(T **dst, const std::vector<T> &src) {
  cudaError_t e = cudaMalloc((void **)dst, src.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}
template <typename T>
static cudaError_t dev_zeros(T **dst, size_t n) {
  cudaError_t e = cudaMalloc((void **)dst, n * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemset(*dst, 0, n * sizeof(T));
}

int tc_prepare(pnb_engine *e, const pnb_model *model) {
  if (!g_encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t r = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    if (r != cudaSuccess || !fn) return tc_fail(PNB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    g_encode = (EncodeTiledFn)fn;
  }
  const pnb_gru_layer *g5[5] = {model->gru1, model->gru2, model->gru3, model->gru_gb, model->gru_rb};
  {  // inputs of the GRUs and output layers are bounded by 1 inside the reference's domain: bound their pre-activations
    double mx = 0.0;
    for (int i = 0; i < 5; i++) {
      const int K = g5[i]->nb_inputs, H = g5[i]->nb_neurons;
      for (int g = 0; g < 3; g++) {
        double v = max_preact(g5[i]->input_weights + g * H, K, H, 3 * H, g5[i]->recurrent_weights + g * H, H,
                              g5[i]->bias + g * H, g5[i]->bias + (3 + g) * H);
        if (!(v <= mx)) mx = v;
      }
    }
    double v = max_preact(model->fc_gb->input_weights, 2560, 34, 34, nullptr, 0, model->fc_gb->bias, nullptr);
    if (!(v <= mx)) mx = v;
    v = max_preact(model->fc_rb->input_weights, 128, 34, 34, nullptr, 0, model->fc_rb->bias, nullptr);
    if (!(v <= mx)) mx = v;
    if (!(mx < kTanhDomain))
      return tc_fail(PNB_ERR_ARG, "PNB_NN_TENSOR: a GRU / output layer can form pre-activations >= 8.5e7 (or NaN), outside the "
                                  "domain of the reference's tansig_approx; use PNB_NN_FP32 for this model");
  }
  pnb_tc_state *t = new pnb_tc_state();
  e->tc = t;
  const size_t S = e->S;
  const size_t Fm = e->Fmax;
  TCK(dev_zeros(&t->fc_all, kConvTerms * (Fm + 4) * S * 128));
  TCK(dev_zeros(&t->c1_all, kConvTerms * (Fm + 2) * S * 512));
  TCK(dev_zeros(&t->wide, 1));
  {
    int w = (e->flags & PNB_CONV_WIDE) ? 1 : 0;
    if (const char *ct = getenv("PNB_CONV_TERMS")) w = atoi(ct) == 3 ? 1 : atoi(ct) == 2 ? 0 : w;  // experiments
    TCK(cudaMemcpy(t->wide, &w, sizeof w, cudaMemcpyHostToDevice));
  }
  TCK(dev_zeros(&t->c2_h, 2 * Fm * S * 512));
  for (int i = 0; i < 5; i++) TCK(dev_zeros(&t->h_all[i], 2 * (Fm + 1) * S * e->gru[i].H));
  t->tiles_mp = (int)((S + 2 * TM - 1) / (2 * TM));
  TCK(dev_zeros(&t->cnt, (size_t)5 * t->tiles_mp));
  {
    std::vector<__nv_bfloat16> p;
    pack_dense_b3(model->conv1->input_weights, 640, 512, p);
    TCK(dev_upload(&t->w_conv1, p));
    pack_dense_b3(model->conv2->input_weights, 1536, 512, p);
    TCK(dev_upload(&t->w_conv2, p));
  }
  for (int i = 0; i < 5; i++) {
    const int K = g5[i]->nb_inputs, H = g5[i]->nb_neurons;
    int ex;
    float sc = pow2_scale_for(g5[i]->input_weights, (size_t)K * 3 * H, g5[i]->recurrent_weights, (size_t)H * 3 * H, &ex);
    t->scale_gru[i] = ldexpf(1.f, -(10 + ex));
    std::vector<__half> p;
    const int order_w[3] = {2, 0, 1}, order_u[3] = {0, 1, 2};
    pack_gru(g5[i]->input_weights, K, H, sc, order_w, p);
    TCK(dev_upload(&t->w_gru[i], p));
    pack_gru(g5[i]->recurrent_weights, H, H, sc, order_u, p);
    TCK(dev_upload(&t->u_gru[i], p));
    // nnet.cpp:136-160: z and r start from b + b' (one float add), the candidate keeps b_n and b'_n apart
    const float *b = g5[i]->bias;
    std::vector<float4> b4(H);
    for (int j = 0; j < H; j++) b4[j] = make_float4(b[j] + b[3 * H + j], b[H + j] + b[4 * H + j], b[2 * H + j], b[5 * H + j]);
    TCK(dev_upload(&t->bias4[i], b4));
  }
  {
    int ex;
    std::vector<__half> p;
    float sc = pow2_scale_for(model->fc_gb->input_weights, (size_t)2560 * 34, nullptr, 0, &ex);
    t->scale_gb = ldexpf(1.f, -(10 + ex));
    pack_small_h2(model->fc_gb->input_weights, 2560, 34, sc, p);
    TCK(dev_upload(&t->w_gb, p));
    sc = pow2_scale_for(model->fc_rb->input_weights, (size_t)128 * 34, nullptr, 0, &ex);
    t->scale_rb = ldexpf(1.f, -(10 + ex));
    pack_small_h2(model->fc_rb->input_weights, 128, 34, sc, p);
    TCK(dev_upload(&t->w_rb, p));
    std::vector<float> bp(512 + 512 + 48 + 48, 0.f);
    memcpy(bp.data(), model->conv1->bias, 512 * 4);
    memcpy(bp.data() + 512, model->conv2->bias, 512 * 4);
    memcpy(bp.data() + 1024, model->fc_gb->bias, 34 * 4);
    memcpy(bp.data() + 1072, model->fc_rb->bias, 34 * 4);
    TCK(dev_upload(&t->bias_pad, bp));
  }
  int bad = 0;
  bad |= make_map(&t->m_wgb, t->w_gb, false, 2 * SMALL_BN, 2560, SMALL_BN / 2);
  bad |= make_map(&t->m_wrb, t->w_rb, false, 2 * SMALL_BN, 128, SMALL_BN / 2);
  bad |= make_map(&t->m_fc_all, t->fc_all, true, kConvTerms * (Fm + 4) * S, 128, TM);
  bad |= make_map(&t->m_c1_all, t->c1_all, true, kConvTerms * (Fm + 2) * S, 512, TM);
  bad |= make_map(&t->m_c2, t->c2_h, false, 2 * Fm * S, 512, TM);
  for (int i = 0; i < 5; i++)
    bad |= make_map(&t->m_h[i], t->h_all[i], false, 2 * (Fm + 1) * S, e->gru[i].H, TM);
  bad |= make_map(&t->m_wconv1, t->w_conv1, true, kConvTerms * 512, 640, DENSE_BN / 2);
  bad |= make_map(&t->m_wconv2, t->w_conv2, true, kConvTerms * 512, 1536, DENSE_BN / 2);
  for (int i = 0; i < 5; i++) {
    const int rows = (e->gru[i].H / HT) * GRU_BN;
    bad |= make_map(&t->m_w[i], t->w_gru[i], false, 2 * rows, e->gru[i].M, GRU_BN / 2);
    bad |= make_map(&t->m_u[i], t->u_gru[i], false, 2 * rows, e->gru[i].H, GRU_BN / 2);
  }
  if (bad) return tc_fail(PNB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  TCK(cudaFuncSetAttribute(gru_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)tc_smem_bytes<GRU_STAGES, GRU_STAGES * StageLayout<GRU_BN>::kBytes>()));
  TCK(cudaFuncSetAttribute(tc_gemm_kernel<DENSE_BN, DENSE_STAGES, DENSE_STAGES3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)dense_smem_bytes<DENSE_BN, DENSE_STAGES, DENSE_STAGES3>()));
  TCK(cudaFuncSetAttribute(tc_gemm_kernel<SMALL_BN, SMALL_STAGES, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)dense_smem_bytes<SMALL_BN, SMALL_STAGES, 0>()));
  return PNB_OK;
}

void tc_release(pnb_engine *e) {
  pnb_tc_state *t = e->tc;
  if (!t) return;
  void *ptrs[] = {t->fc_all, t->c1_all, t->c2_h, t->w_conv1, t->w_conv2, t->w_gb, t->w_rb, t->bias_pad, t->cnt, t->wide};
  for (void *p : ptrs) if (p) cudaFree(p);
  for (int i = 0; i < 5; i++) {
    if (t->h_all[i]) cudaFree(t->h_all[i]);
    if (t->w_gru[i]) cudaFree(t->w_gru[i]);
    if (t->u_gru[i]) cudaFree(t->u_gru[i]);
    if (t->bias4[i]) cudaFree(t->bias4[i]);
  }
  delete t;
  e->tc = nullptr;
}

int tc_reset(pnb_engine *e) {
  pnb_tc_state *t = e->tc;
  if (!t) return PNB_OK;
  const size_t S = e->S;
  const size_t Fm = e->Fmax;
  TCK(cudaMemset(t->fc_all, 0, kConvTerms * (Fm + 4) * S * 128 * 2));
  TCK(cudaMemset(t->c1_all, 0, kConvTerms * (Fm + 2) * S * 512 * 2));
  TCK(cudaMemset(t->c2_h, 0, 2 * Fm * S * 512 * 2));
  for (int i = 0; i < 5; i++) TCK(cudaMemset(t->h_all[i], 0, 2 * (Fm + 1) * S * e->gru[i].H * 2));
  TCK(cudaMemset(t->cnt, 0, (size_t)5 * t->tiles_mp * sizeof(unsigned)));
  return PNB_OK;
}

// One stream's conv input history (the last 4 fc outputs, the last 2 conv1 outputs; oldest first) as fp32, for
// pnb_get_state / pnb_set_state.  Between calls the history sits in hop slots 0..3 / 0..1 as two bf16 terms.
int tc_get_stream_hist(pnb_engine *e, int s, float *fc_hist, float *c1_hist, unsigned short *fc_terms, unsigned short *c1_terms) {
  pnb_tc_state *t = e->tc;
  const size_t S = e->S, Fm = e->Fmax;
  static_assert(sizeof(__nv_bfloat16) == sizeof(unsigned short), "raw bf16 terms travel as 16-bit words");
  auto get = [&](const __nv_bfloat16 *base, size_t plane, int slots, int width, float *sum, unsigned short *terms) -> cudaError_t {
    for (int i = 0; i < slots; i++) {
      for (int k = 0; k < width; k++) sum[i * width + k] = 0.f;
      for (int tm = kConvTerms - 1; tm >= 0; tm--) {  // smallest term first: the fp32 sum is then exact up to its last bit
        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(terms + ((size_t)tm * slots + i) * width);
        cudaError_t r = cudaMemcpy(dst, base + (size_t)tm * plane + ((size_t)i * S + s) * width, (size_t)width * 2, cudaMemcpyDeviceToHost);
        if (r != cudaSuccess) return r;
        for (int k = 0; k < width; k++) sum[i * width + k] += __bfloat162float(dst[k]);
      }
    }
    return cudaSuccess;
  };
  TCK(get(t->fc_all, (Fm + 4) * S * 128, 4, 128, fc_hist, fc_terms));
  TCK(get(t->c1_all, (Fm + 2) * S * 512, 2, 512, c1_hist, c1_terms));
  return PNB_OK;
}
// h: the five GRU states of the stream, concatenated (4 x 512 + 128), already stored as fp32 by the caller.
// fc_terms / c1_terms (from a tensor-mode engine) are restored verbatim; without them the fp32 values are split.
int tc_set_stream_hist(pnb_engine *e, int s, const float *fc_hist, const float *c1_hist, const float *h,
                       const unsigned short *fc_terms, const unsigned short *c1_terms) {
  pnb_tc_state *t = e->tc;
  const size_t S = e->S, Fm = e->Fmax;
  auto put = [&](__nv_bfloat16 *base, size_t plane, int slots, int width, const float *sum, const unsigned short *terms) -> cudaError_t {
    std::vector<__nv_bfloat16> buf(width);
    for (int i = 0; i < slots; i++)
      for (int tm = 0; tm < kConvTerms; tm++) {
        for (int k = 0; k < width; k++) {
          if (terms) {
            buf[k] = reinterpret_cast<const __nv_bfloat16 *>(terms + ((size_t)tm * slots + i) * width)[k];
          } else {
            float r = sum[i * width + k];
            __nv_bfloat16 b = __float2bfloat16_rn(r);
            for (int q = 0; q < tm; q++) { r -= __bfloat162float(b); b = __float2bfloat16_rn(r); }
            buf[k] = b;
          }
        }
        cudaError_t r = cudaMemcpy(base + (size_t)tm * plane + ((size_t)i * S + s) * width, buf.data(), (size_t)width * 2, cudaMemcpyHostToDevice);
        if (r != cudaSuccess) return r;
      }
    return cudaSuccess;
  };
  TCK(put(t->fc_all, (Fm + 4) * S * 128, 4, 128, fc_hist, fc_terms));
  TCK(put(t->c1_all, (Fm + 2) * S * 512, 2, 512, c1_hist, c1_terms));
  size_t off = 0;
  for (int li = 0; li < 5; li++) {
    const size_t H = e->gru[li].H;
    std::vector<__half> hi(H), lo(H);
    for (size_t k = 0; k < H; k++) {
      const float sv = h[off + k] * kActScale;
      hi[k] = __float2half_rn(sv);
      lo[k] = __float2half_rn(sv - __half2float(hi[k]));
    }
    __half *p0 = t->h_all[li] + (size_t)s * H;  // slot 0
    TCK(cudaMemcpy(p0, hi.data(), H * 2, cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(p0 + (Fm + 1) * S * H, lo.data(), H * 2, cudaMemcpyHostToDevice));
    off += H;
  }
  return PNB_OK;
}

#ifdef PNB_CHAIN_TIMING
extern "C" int pnb_debug_chain_cycles(unsigned long long *out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_chain_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_chain_cycles, z, sizeof z);
  }
  return 0;
}
#endif

int tc_launches_per_chunk(const pnb_engine *) { return 6; }  // fc, conv1, conv2, GRU chain, fc_gb, fc_rb (+ 1 carry per call)

template <typename K, typename A>
static int tc_launch(pnb_engine *e, K kernel, const A &a, int pairs, size_t smem, cudaStream_t st) {
  int clusters = pairs < e->tc_sms / 2 ? pairs : e->tc_sms / 2;  // persistent: one CTA per SM of the network's share
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TCK(cudaLaunchKernelEx(&cfg, kernel, a));
  return PNB_OK;
}
template <int BN, int ST2, int ST3>
static int tc_launch_dense(pnb_engine *e, TcArgs &a, int rows, int tiles_n, cudaStream_t st) {
  a.M = rows;
  a.tiles_m = (rows + TM - 1) / TM;
  a.tiles_n = tiles_n;
  a.status = e->d_status;
  const int pairs = ((a.tiles_m + 1) / 2) * a.tiles_n;
  return tc_launch(e, tc_gemm_kernel<BN, ST2, ST3>, a, pairs, dense_smem_bytes<BN, ST2, ST3>(), st);
}

static TcSeg mkseg(int a_map, int b_map, int K, int a_k0, int b_k0, int a_row0, int a_term_rows) {
  TcSeg s;
  s.a_map = a_map; s.b_map = b_map; s.k_blocks = K / BK; s.a_k0 = a_k0; s.b_k0 = b_k0;
  s.a_row0 = a_row0; s.a_term_rows = a_term_rows;
  return s;
}

// ---- the network of hops [h0, h0 + n) of a call of F hops, in three phases -----------------------------------
// The non-recurrent front (rnn.cpp:50-52 on n S rows): fc (tc_fc) -> conv1 -> conv2 (tc_front).  Hop t of the call lives in slot t of the
// per-call buffers, so a phase can be run for any hop range once the earlier hops' phases are enqueued.
// Returns the number of launches or a negative error.
int tc_fc(pnb_engine *e, int h0, int n, cudaStream_t st) {
  // fc (fp32 FMA on the CUDA cores: it runs with the DSP kernels when the schedule is split) -> bf16 terms at slots 4 + h0 ..
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax, rows = n * S;
  ProfScope ps(e, PNB_K_TC_AUX, st);
  fc_split_kernel<<<(rows + FC_ROWS - 1) / FC_ROWS, 128, 0, st>>>(e->d_feat + (size_t)h0 * S * kFeat, e->fc.W, e->fc.b,
                                                  t->fc_all + (size_t)(4 + h0) * S * 128, rows, (size_t)(Fm + 4) * S);
  return 1;
}
int tc_front(pnb_engine *e, int h0, int n, int F, cudaStream_t st) {
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax;
  const float *tbl = e->tansig();
  const int rows = n * S;
  int nl = 0, rc;
  TcArgs a;
  // conv1: tap q of hop t reads fc slot t + q (oldest first, nnet.cpp:182-200); output -> c1 slots 2 + t
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_fc_all;
  a.maps[1] = t->m_wconv1;
  for (int q = 0; q < 5; q++) a.seg[q] = mkseg(0, 1, 128, 0, q * 128, (h0 + q) * S, (Fm + 4) * S);
  a.n_seg = 5; a.b_term_rows = 512; a.fmt = 1; a.out_scale = 1.f; a.tansig = tbl;
  a.bias = t->bias_pad; a.act = PNB_ACT_RELU; a.N = 512;
  a.out_b = t->c1_all; a.out_row0 = (2 + h0) * S; a.out_plane_rows = (Fm + 2) * S; a.out_terms = kConvTerms; a.wide = t->wide;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    if ((rc = tc_launch_dense<DENSE_BN, DENSE_STAGES, DENSE_STAGES3>(e, a, rows, 512 / DENSE_BN, st))) return rc;
    nl++;
  }
  // conv2: three taps -> tanh -> fp16 terms (gru1 / gru_rb / fc_gb inputs); fp32 copy of the call's last hop
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_c1_all;
  a.maps[1] = t->m_wconv2;
  for (int q = 0; q < 3; q++) a.seg[q] = mkseg(0, 1, 512, 0, q * 512, (h0 + q) * S, (Fm + 2) * S);
  a.n_seg = 3; a.b_term_rows = 512; a.fmt = 1; a.out_scale = 1.f; a.tansig = tbl;
  a.bias = t->bias_pad + 512; a.act = PNB_ACT_TANH; a.N = 512;
  if (h0 + n == F) { a.out_f32 = e->c2; a.ldc = 512; a.f32_row0 = (n - 1) * S; }
  a.out_h = t->c2_h; a.out_row0 = h0 * S; a.out_plane_rows = Fm * S; a.wide = t->wide;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    if ((rc = tc_launch_dense<DENSE_BN, DENSE_STAGES, DENSE_STAGES3>(e, a, rows, 512 / DENSE_BN, st))) return rc;
    nl++;
  }
  return nl;
}

// The recurrent part: five GRUs per hop (rnn.cpp:58-71) in one persistent launch.  State slot t holds the state
// BEFORE hop t of the call; the fp32 state ping-pongs by hop, starting from e->par (the call's first hop).
int tc_gru_chain(pnb_engine *e, int h0, int n, cudaStream_t st) {
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax;
  ChainArgs a;
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_c2;
  for (int i = 0; i < 5; i++) { a.maps[1 + i] = t->m_h[i]; a.maps[6 + i] = t->m_w[i]; a.maps[11 + i] = t->m_u[i]; }
  a.S = S; a.F = n; a.t0 = h0; a.tiles_mp = t->tiles_mp; a.cnt = t->cnt; a.tansig = e->tansig();
  const int c2_terms = Fm * S, h_terms = (Fm + 1) * S;
  int unit0 = 0;
  for (int li = 0; li < 5; li++) {
    ChainLayer &L = a.L[li];
    const int H = e->gru[li].H, tiles = H / HT;
    L.H = H; L.h_kb = H / BK; L.tiles_log2 = tiles == 8 ? 3 : tiles == 4 ? 2 : tiles == 2 ? 1 : 0;
    L.b_term_rows = tiles * GRU_BN;
    L.h_map = 1 + li; L.w_map = 6 + li; L.u_map = 11 + li;
    L.unit0 = unit0; unit0 += t->tiles_mp * tiles;
    L.par0 = e->par[li];
    L.scale = t->scale_gru[li];
    L.bias4 = t->bias4[li];
    L.h32[0] = e->h[li][0]; L.h32[1] = e->h[li][1];
    L.h_hl = t->h_all[li]; L.h_term_rows = h_terms;
    // each GRU consumes the freshly written state (slot t+1) of the layer below (rnn.cpp:58-71)
    if (li == 0) {          // gru1 <- conv2 out
      L.n_x = 1; L.x_map[0] = 0; L.x_kb[0] = 512 / BK; L.x_bk0[0] = 0; L.x_row1[0] = 0; L.x_term_rows[0] = c2_terms; L.dep = -1;
    } else if (li < 4) {    // gru2 <- gru1, gru3 <- gru2, gru_gb <- gru3
      L.n_x = 1; L.x_map[0] = li; L.x_kb[0] = 512 / BK; L.x_bk0[0] = 0; L.x_row1[0] = 1; L.x_term_rows[0] = h_terms; L.dep = li - 1;
    } else {                // gru_rb <- [gru3 state, conv2 out]
      L.n_x = 2;
      L.x_map[0] = 3; L.x_kb[0] = 512 / BK; L.x_bk0[0] = 0; L.x_row1[0] = 1; L.x_term_rows[0] = h_terms;
      L.x_map[1] = 0; L.x_kb[1] = 512 / BK; L.x_bk0[1] = 512; L.x_row1[1] = 0; L.x_term_rows[1] = c2_terms;
      L.dep = 2;
    }
  }
  a.units_per_hop = unit0;
  // with many streams one layer-hop fills the machine and hop-major order keeps a layer's fresh output in L2 for the
  // layer above; with few, only the anti-diagonal order exposes enough independent units
  a.diag = (t->tiles_mp * 8 < e->tc_sms) ? 1 : 0;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    int rc = tc_launch(e, gru_chain_kernel, a, n * unit0, tc_smem_bytes<GRU_STAGES, GRU_STAGES * StageLayout<GRU_BN>::kBytes>(), st);
    if (rc) return rc;
  }
  return 1;
}

// The two 34-wide output layers (rnn.cpp:73-80; N padded to 48, fp16 two-term split, sigmoid epilogue).
int tc_out(pnb_engine *e, int h0, int n, cudaStream_t st) {
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax;
  const float *tbl = e->tansig();
  const int rows = n * S, h_terms = (Fm + 1) * S;
  int nl = 0, rc;
  TcArgs a;
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_c2;
  for (int q = 0; q < 4; q++) a.maps[1 + q] = t->m_h[q];
  a.maps[5] = t->m_wgb;
  a.seg[0] = mkseg(0, 5, 512, 0, 0, h0 * S, Fm * S);
  for (int q = 1; q < 5; q++) a.seg[q] = mkseg(q, 5, 512, 0, q * 512, (h0 + 1) * S, h_terms);  // state after hop t = slot t+1
  a.n_seg = 5; a.b_term_rows = SMALL_BN; a.fmt = 0; a.out_scale = t->scale_gb; a.tansig = tbl;
  a.bias = t->bias_pad + 1024; a.act = PNB_ACT_SIGMOID; a.N = 34; a.ldc = 68; a.out_f32 = e->d_gr + (size_t)h0 * S * 68;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    if ((rc = tc_launch_dense<SMALL_BN, SMALL_STAGES, 0>(e, a, rows, 1, st))) return rc;
    nl++;
  }
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_h[4];
  a.maps[1] = t->m_wrb;
  a.seg[0] = mkseg(0, 1, 128, 0, 0, (h0 + 1) * S, h_terms);
  a.n_seg = 1; a.b_term_rows = SMALL_BN; a.fmt = 0; a.out_scale = t->scale_rb; a.tansig = tbl;
  a.bias = t->bias_pad + 1072; a.act = PNB_ACT_SIGMOID; a.N = 34; a.ldc = 68; a.out_f32 = e->d_gr + (size_t)h0 * S * 68 + 34;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    if ((rc = tc_launch_dense<SMALL_BN, SMALL_STAGES, 0>(e, a, rows, 1, st))) return rc;
    nl++;
  }
  return nl;
}

// End of a call of F hops: the last 4 fc / 2 conv1 hop slots and the last state slot move to the front, the chain's
// dependency counters return to zero (one launch), and the fp32 state parity advances by F.
int tc_carry(pnb_engine *e, int F, cudaStream_t st) {
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax;
  const int h_terms = (Fm + 1) * S;
  CarryArgs ca;
  memset(&ca, 0, sizeof ca);
  ca.cnt = t->cnt; ca.n_cnt = 5 * t->tiles_mp;
  int ns = 0;
  for (int tm = 0; tm < kConvTerms; tm++) {
    __nv_bfloat16 *bf = t->fc_all + (size_t)tm * (Fm + 4) * S * 128;
    __nv_bfloat16 *bc = t->c1_all + (size_t)tm * (Fm + 2) * S * 512;
    ca.seg[ns++] = CarrySeg{reinterpret_cast<uint4 *>(bf), (size_t)S * 128 * 2 / 16, 4, F};
    ca.seg[ns++] = CarrySeg{reinterpret_cast<uint4 *>(bc), (size_t)S * 512 * 2 / 16, 2, F};
  }
  for (int li = 0; li < 5; li++) {
    const size_t H = e->gru[li].H;
    for (int tm = 0; tm < 2; tm++) {
      __half *base = t->h_all[li] + (size_t)tm * h_terms * H;
      ca.seg[ns++] = CarrySeg{reinterpret_cast<uint4 *>(base), (size_t)S * H * 2 / 16, 1, F};
    }
  }
  ca.n_seg = ns;
  {
    ProfScope ps(e, PNB_K_TC_AUX, st);
    tc_carry_kernel<<<2 * e->tc_sms, 256, 0, st>>>(ca);
  }
  if (F & 1)
    for (int li = 0; li < 5; li++) e->par[li] ^= 1;
  return 1;
}
