// pnb_nn_tc.cu -- the gain network on Blackwell tensor cores (tcgen05 + TMEM + TMA), batched over streams.
//
// Replaces the reference's 35 scalar GEMVs per hop (/root/reference/src/nnet.cpp:59-200, rnn.cpp:42-81)
// for the layers that carry 99 % of the MACs: conv1, conv2 and the five GRUs.  The rows of every
// contraction are the concurrent streams (M = S), so the operation is genuinely dense.
//
// Precision.  The parity bar (1e-4 relative on g/r, +-1 LSB on PCM) rules out plain 16-bit operands.
// Every fp32 operand is therefore split into 16-bit terms whose sum reproduces it, and the product is
// expanded into the tensor-core products that matter, all accumulated in fp32 in TMEM:
//   GRUs (bounded inputs):    x = xh + xl, w = wh + wl in fp16, both pre-scaled by powers of two
//                             x w ~= xh wh + xh wl + xl wh                       (3 MMAs, error ~2^-22)
//   conv1/conv2 (ReLU inputs, unbounded): three bf16 terms each side, the six products >= 2^-16 (6 MMAs)
// Per tile: TMA (SWIZZLE_64B boxes of 32 k) -> 4/5-stage shared-memory ring -> one elected thread issues
// tcgen05.mma (M=128) -> accumulators in TMEM -> four epilogue warps read them back with tcgen05.ld and
// apply bias, the reference's tansig-table activations and the GRU gate arithmetic, writing the new
// state in fp32 and already split for the next contraction.
//
// The GRU tile holds, for 64 hidden units, four accumulators per unit (z, r, W_n x, U_n h: nnet.cpp:136-173
// needs the last two apart), 256 TMEM columns [nx | z | r | nh]; weight rows are packed gate-interleaved per tile
// (input part [n|z|r], recurrent part [z|r|n]) so that one TMA box brings a tile's rows and one N = 192 MMA
// covers it -- the kernel is bound by shared-memory operand reads (A is re-read by every MMA), so fewer, wider
// MMAs matter more than anything else here.
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

constexpr int TM = 128;   // streams per tile (UMMA M)
constexpr int BK = 32;    // k per pipeline stage: 64-byte rows, SWIZZLE_64B
constexpr int HT = 64;    // hidden units per GRU tile
constexpr int GRU_BN = 3 * HT;  // weight rows per GRU tile (z|r|n)
// Epilogue warps per CTA: warp 4 + q + 4 k reads TMEM lane quadrant q; the kEpiSets warps of a quadrant share its
// rows' columns 16 at a time.  The epilogue (gate math, splits, stores) is what a tile costs most, see DESIGN.md 3.2.
constexpr int kEpiSets = 2;
constexpr int kTcThreads = 32 * (4 + 4 * kEpiSets);
constexpr int DENSE_BN = 256;   // output columns per dense (conv) tile: the widest MMA, least operand traffic per MAC
constexpr int kConvTerms = 2;         // bf16 terms per operand on the conv layers (2: 16-bit operands, 3 products)
constexpr float kActScale = 1024.f;  // fp16 activations are stored x 2^10 (keeps the low term normal)

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
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
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(x), "r"(y),
      "l"(0x1000000000000000ull)  // EVICT_NORMAL
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
// D[256 x N] (+)= A[256 x 16] B[N x 16]^T : rows 0..127 from the leader's A tile / TMEM, 128..255 from the peer's;
// each CTA's shared memory supplies N/2 rows of B
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all MMAs issued so far arrives on `bar` in BOTH CTAs
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major operand, SWIZZLE_64B: rows of 64 bytes, 8-row groups 512 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO [16,30), SBO [32,46), version=1 [46,48), layout [61,64) with 4 = SW64)
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A/B format (0 fp16, 1 bf16), both K-major
__host__ __device__ constexpr uint32_t idesc_f16(int n, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * TM) >> 4) << 24);  // M = 256: the CTA pair
}

// ------------------------------------------------------------------------------------------ activations
__device__ __forceinline__ float tansig_s(float x, const float *tbl) {  // vec.h:53-71, table in shared memory
  float sign = 1.f;
  if (x < 0.f) { x = -x; sign = -1.f; }
  float fi = floorf(.5f + 25.f * x);
  int i = (fi < 2147483648.f) ? (int)fi : (int)0x80000000;  // x86 cvttss2si semantics of the reference build
  i = i < 200 ? i : 200;
  i = i > 0 ? i : 0;
  x -= .04f * i;
  float y = tbl[i];
  float dy = 1.f - y * y;
  y = y + x * dy * (1.f - y * x);
  return sign * y;
}
__device__ __forceinline__ float sigmoid_s(float x, const float *tbl) { return .5f + .5f * tansig_s(.5f * x, tbl); }

// ------------------------------------------------------------------------------------------ kernel
struct TcSeg {
  int a_map, b_map;  // indices into TcArgs::maps
  int k_blocks;      // K / 32
  int a_k0, b_k0;    // starting k (elements) in the A / B matrices
  int recurrent;     // GRU: 0 -> the n rows accumulate W_n x, 1 -> U_n h
  int a_row0;        // row of the A matrix that holds tile row 0 (hop / tap offset inside a multi-hop buffer)
  int a_term_rows;   // row distance between the split terms inside this A buffer
};

struct TcArgs {
  CUtensorMap maps[10];
  TcSeg seg[5];
  int n_seg;
  int M;                // streams
  int tiles_m, tiles_n; // tile grid; a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
  int b_term_rows;      // row distance between the split terms inside a packed weight matrix
  int fmt;              // 0 fp16, 1 bf16
  int debug;            // timing diagnostics (PNB_TC_DEBUG): 1 = issue no MMA (operand streaming only), 2 = issue no TMA load
  float out_scale;      // 2^-(activation scale + weight scale)
  const float *tansig;  // 201-entry table (global)
  const float *bias;
  // dense epilogue
  int act;
  int N;                // valid output columns
  int ldc;              // row stride of out_f32
  float *out_f32;       // [M][ldc] or null
  int f32_row0;         // only rows >= f32_row0 are written to out_f32, at row - f32_row0
  __half *out_h;        // fp16 two-term split [2][out_plane_rows][N] (x 2^10) or null
  __nv_bfloat16 *out_b; // bf16 split [terms][out_plane_rows][N] or null
  int out_row0;         // first row written inside a term plane of out_h / out_b
  int out_plane_rows;   // rows per term plane of out_h / out_b
  // GRU epilogue
  int H;
  const float *h_old;   // [M][H] fp32
  float *h_new;         // [M][H] fp32
  __half *h_new_h;      // fp16 split of the new state (x 2^10): [2][out_plane_rows][H], first row out_row0
};

template <int NA, int NB, int BN>
struct StageLayout {
  static constexpr int kABytes = TM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;  // each CTA of the pair holds half of the weight rows
  static constexpr int kBytes = NA * kABytes + NB * kBBytes;
  static constexpr int kPairBytes = kBytes;          // bytes one CTA brings per stage
};

// products of split terms that are kept: (a term, b term), largest first
template <int NA, int NB> struct Products;
template <> struct Products<2, 2> {
  static constexpr int n = 3;
  __device__ static constexpr int a(int i) { return i == 2 ? 1 : 0; }
  __device__ static constexpr int b(int i) { return i == 1 ? 1 : 0; }
};
__device__ __forceinline__ void split_h2(float v, __half &hi, __half &lo) {
  // fp16 operands hold |x| < 64 (x 2^10 < 65504).  GRU inputs are tanh / state values in [-1, 1]; only when
  // the reference's own tansig_approx leaves its defined domain (|pre-activation| >= 8.6e7, see tests) does
  // a "tanh" exceed that, and then this path saturates instead of propagating the reference's garbage.
  float s = fminf(fmaxf(v * kActScale, -65000.f), 65000.f);
  hi = __float2half_rn(s);
  lo = __float2half_rn(s - __half2float(hi));
}
__device__ __forceinline__ void split_b3(float v, __nv_bfloat16 &t0, __nv_bfloat16 &t1, __nv_bfloat16 &t2) {
  t0 = __float2bfloat16_rn(v);
  float r = v - __bfloat162float(t0);
  t1 = __float2bfloat16_rn(r);
  r = r - __bfloat162float(t1);
  t2 = __float2bfloat16_rn(r);
}

__host__ __device__ constexpr int pow2_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue.  Two accumulator buffers in TMEM: the epilogue of tile i overlaps the MMAs of tile i+1.
// CTAs run as pairs (cluster of 2, tcgen05 cta_group::2): the pair works on one weight tile for two adjacent
// 128-stream row tiles as a single M = 256 MMA.  Each CTA loads its own A rows and HALF of the weight rows, the
// leader issues the MMAs and the tensor cores of both SMs read B from both shared memories -- per SM that halves
// the B operand traffic (global->shared and shared->tensor core), which is what bounds this kernel.
// The accumulators are zeroed by the epilogue warps (tcgen05.st) so that every MMA accumulates.
template <int NA, int NB, int BN, int STAGES, bool GRU>
__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_kernel(const __grid_constant__ TcArgs args) {
  using SL = StageLayout<NA, NB, BN>;
  using PR = Products<NA, NB>;
  constexpr int kAccCols = GRU ? 4 * HT : pow2_cols(BN);  // TMEM columns per accumulator buffer
  constexpr int kTmemCols = 2 * kAccCols;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the swizzled tiles need 512-byte (SW64) alignment; align the carve-up to 1024 by hand
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // carve: stages | barriers | tmem slot | tansig table
  uint8_t *stage_base = smem;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * SL::kBytes);
  uint64_t *empty_bar = full_bar + STAGES;
  uint64_t *acc_full = empty_bar + STAGES;   // [2] MMA -> epilogue
  uint64_t *acc_empty = acc_full + 2;        // [2] epilogue -> MMA
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  float *tbl = reinterpret_cast<float *>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();  // 0 = leader
  const int n_pairs_cta = gridDim.x >> 1, pair_id = blockIdx.x >> 1;
  const int total_pairs = ((args.tiles_m + 1) >> 1) * args.tiles_n;  // work unit: one B tile x two row tiles

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * 128 * kEpiSets); }  // every epilogue thread of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < 201; i += blockDim.x) tbl[i] = args.tansig[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 4) {  // zero both accumulator buffers: every MMA accumulates
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = 0; c < kTmemCols; c += 16) tmem_st16_zero(tl + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM are ready before the leader issues anything
  tc_fence_after();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows, own half of the B rows; bytes counted on the leader's barrier =====
    if (lane == 0) {
      int it = 0;
      for (int pair = pair_id; pair < total_pairs; pair += n_pairs_cta) {
        const int n_tile = pair % args.tiles_n, m0 = (2 * (pair / args.tiles_n) + crank) * TM;
        for (int s = 0; s < args.n_seg; s++) {
          const TcSeg sg = args.seg[s];
          const CUtensorMap *ma = &args.maps[sg.a_map], *mb = &args.maps[sg.b_map];
          for (int kb = 0; kb < sg.k_blocks; kb++, it++) {
            const int st = it % STAGES;
            mbar_wait(&empty_bar[st], ((it / STAGES) & 1) ^ 1);  // fresh barrier: passes immediately
            if (args.debug & 2) {  // diagnostic: the MMAs run on whatever the stage holds
              if (crank == 0) mbar_expect_tx(&full_bar[st], 0);
              continue;
            }
            if (crank == 0) mbar_expect_tx(&full_bar[st], 2 * SL::kPairBytes);
            uint8_t *sp = stage_base + st * SL::kBytes;
#pragma unroll
            for (int ta = 0; ta < NA; ta++)
              tma_load_2d_pair(sp + ta * SL::kABytes, ma, &full_bar[st], sg.a_k0 + kb * BK, ta * sg.a_term_rows + sg.a_row0 + m0);
#pragma unroll
            for (int tb = 0; tb < NB; tb++)
              tma_load_2d_pair(sp + NA * SL::kABytes + tb * SL::kBBytes, mb, &full_bar[st], sg.b_k0 + kb * BK,
                               tb * args.b_term_rows + n_tile * BN + crank * (BN / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores =====
    if (lane == 0 && crank == 0) {
      const uint32_t id_main = idesc_f16(GRU ? 3 * HT : BN, args.fmt);
      int it = 0, j = 0;
      for (int pair = pair_id; pair < total_pairs; pair += n_pairs_cta, j++) {
        const int buf = j & 1;
        const uint32_t acc = tmem_base + buf * kAccCols;
        mbar_wait(&acc_empty[buf], ((j >> 1) & 1) ^ 1);  // both epilogues drained (and re-zeroed) this buffer
        tc_fence_after();
        for (int s = 0; s < args.n_seg; s++) {
          const TcSeg sg = args.seg[s];
          // GRU accumulator columns [nx | z | r | nh]: input part (rows [n|z|r]) at column 0, recurrent part
          // (rows [z|r|n]) at column 64; each is one N = 192 MMA
          const uint32_t dcol = (GRU && sg.recurrent) ? acc + HT : acc;
          for (int kb = 0; kb < sg.k_blocks; kb++, it++) {
            const int st = it % STAGES;
            mbar_wait(&full_bar[st], (it / STAGES) & 1);
            tc_fence_after();
            const uint32_t sa = smem_u32(stage_base + st * SL::kBytes);
            const uint32_t sb = sa + NA * SL::kABytes;
#pragma unroll
            for (int ks = 0; ks < BK / 16; ks++) {
#pragma unroll
              for (int p = 0; p < PR::n; p++) {
                const uint64_t ad = smem_desc_sw64(sa + PR::a(p) * SL::kABytes) + (uint64_t)(ks * 2);
                const uint64_t bd = smem_desc_sw64(sb + PR::b(p) * SL::kBBytes) + (uint64_t)(ks * 2);
                if (!(args.debug & 1)) umma_f16(dcol, ad, bd, id_main, 1u);
              }
            }
            umma_commit(&empty_bar[st]);  // frees the stage in both CTAs once these MMAs have read it
          }
        }
        umma_commit(&acc_full[buf]);  // both epilogues
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31, one stream per thread =====
    const int wq = warp & 3, eset = (warp - 4) >> 2;
    const float sc = args.out_scale;
    int j = 0;
    for (int pair = pair_id; pair < total_pairs; pair += n_pairs_cta, j++) {
      const int n_tile = pair % args.tiles_n, m0 = (2 * (pair / args.tiles_n) + crank) * TM;
      const int buf = j & 1;
      mbar_wait(&acc_full[buf], (j >> 1) & 1);
      tc_fence_after();
      const int row = m0 + wq * 32 + lane;
      const bool row_ok = row < args.M;
      const uint32_t tlane = tmem_base + buf * kAccCols + ((uint32_t)(wq * 32) << 16);
      if (GRU) {
        const int H = args.H;
        const float *b = args.bias;
        for (int c = 16 * eset; c < HT; c += 16 * kEpiSets) {
          float zs[16], rs[16], nx[16], nh[16];
          tmem_ld16(tlane + HT + c, zs);      // accumulator columns: [nx | z | r | nh]
          tmem_ld16(tlane + 2 * HT + c, rs);
          tmem_ld16(tlane + c, nx);
          tmem_ld16(tlane + 3 * HT + c, nh);
          const int j0 = n_tile * HT + c;
          if (row_ok) {
            float hn[16];
            const float4 *ho = reinterpret_cast<const float4 *>(args.h_old + (size_t)row * H + j0);
            float hold[16];
#pragma unroll
            for (int q = 0; q < 4; q++) {
              float4 t = ho[q];
              hold[4 * q] = t.x; hold[4 * q + 1] = t.y; hold[4 * q + 2] = t.z; hold[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {  // nnet.cpp:136-178 (reset_after)
              const int jj = j0 + i;
              float z = sigmoid_s((__ldg(b + jj) + __ldg(b + 3 * H + jj)) + zs[i] * sc, tbl);
              float r = sigmoid_s((__ldg(b + H + jj) + __ldg(b + 4 * H + jj)) + rs[i] * sc, tbl);
              float tmp = __ldg(b + 5 * H + jj) + nh[i] * sc;
              float cnd = __ldg(b + 2 * H + jj) + tmp * r;
              cnd = cnd + nx[i] * sc;
              float n = tansig_s(cnd, tbl);
              hn[i] = z * hold[i] + (1.f - z) * n;
            }
            float4 *hw = reinterpret_cast<float4 *>(args.h_new + (size_t)row * H + j0);
#pragma unroll
            for (int q = 0; q < 4; q++) hw[q] = make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
            __align__(16) __half hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; i++) split_h2(hn[i], hi[i], lo[i]);
            const size_t orow = (size_t)args.out_row0 + row;
            uint4 *ph = reinterpret_cast<uint4 *>(args.h_new_h + orow * H + j0);
            uint4 *pl = reinterpret_cast<uint4 *>(args.h_new_h + ((size_t)args.out_plane_rows + orow) * H + j0);
            ph[0] = reinterpret_cast<uint4 *>(hi)[0]; ph[1] = reinterpret_cast<uint4 *>(hi)[1];
            pl[0] = reinterpret_cast<uint4 *>(lo)[0]; pl[1] = reinterpret_cast<uint4 *>(lo)[1];
          }
        }
      } else {
        const int N = args.N;
        for (int c = 16 * eset; c < BN; c += 16 * kEpiSets) {
          float d[16];
          tmem_ld16(tlane + c, d);
          const int j0 = n_tile * BN + c;
          if (row_ok && j0 < N) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
              float x = d[i] * sc + ((j0 + i < N) ? __ldg(args.bias + j0 + i) : 0.f);
              v[i] = args.act == PNB_ACT_TANH ? tansig_s(x, tbl) : args.act == PNB_ACT_RELU ? (x < 0.f ? 0.f : x)
                   : args.act == PNB_ACT_SIGMOID ? sigmoid_s(x, tbl) : x;
            }
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
            if (args.out_h) {
              __align__(16) __half hi[16], lo[16];
#pragma unroll
              for (int i = 0; i < 16; i++) split_h2(v[i], hi[i], lo[i]);
              const size_t orow = (size_t)args.out_row0 + row;
              uint4 *ph = reinterpret_cast<uint4 *>(args.out_h + orow * N + j0);
              uint4 *pl = reinterpret_cast<uint4 *>(args.out_h + ((size_t)args.out_plane_rows + orow) * N + j0);
              ph[0] = reinterpret_cast<uint4 *>(hi)[0]; ph[1] = reinterpret_cast<uint4 *>(hi)[1];
              pl[0] = reinterpret_cast<uint4 *>(lo)[0]; pl[1] = reinterpret_cast<uint4 *>(lo)[1];
            }
            if (args.out_b) {
              __align__(16) __nv_bfloat16 t0[16], t1[16], t2[16];
#pragma unroll
              for (int i = 0; i < 16; i++) split_b3(v[i], t0[i], t1[i], t2[i]);
              const size_t plane = (size_t)args.out_plane_rows * N;
              const size_t orow = (size_t)args.out_row0 + row;
              uint4 *p0 = reinterpret_cast<uint4 *>(args.out_b + orow * N + j0);
              uint4 *p1 = reinterpret_cast<uint4 *>(args.out_b + plane + orow * N + j0);
              p0[0] = reinterpret_cast<uint4 *>(t0)[0]; p0[1] = reinterpret_cast<uint4 *>(t0)[1];
              p1[0] = reinterpret_cast<uint4 *>(t1)[0]; p1[1] = reinterpret_cast<uint4 *>(t1)[1];
              if (kConvTerms == 3) {
                uint4 *p2 = reinterpret_cast<uint4 *>(args.out_b + 2 * plane + orow * N + j0);
                p2[0] = reinterpret_cast<uint4 *>(t2)[0]; p2[1] = reinterpret_cast<uint4 *>(t2)[1];
              }
            }
          }
        }
      }
      // zero the buffer for its next tile and hand it back to the leader's MMA thread
      for (int c = 16 * eset; c < kAccCols; c += 16 * kEpiSets) tmem_st16_zero(tlane + c);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_leader(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the pair still has work in flight
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

constexpr int GRU_STAGES = 7, DENSE_STAGES = 6, SMALL_BN = 48, SMALL_STAGES = 8;
template <int NA, int NB, int BN, int STAGES>
constexpr size_t tc_smem_bytes() {
  return (size_t)STAGES * StageLayout<NA, NB, BN>::kBytes + (2 * STAGES + 4) * 8 + 16 + 208 * 4 + 1024;
}

// fc 70 -> 128 relu in fp32 FMA (0.1 % of the MACs; K = 70 is no tensor-core shape), emitting the three bf16
// terms conv1 consumes.  One block per 4 streams, one thread per output.
__global__ void __launch_bounds__(128) fc_split_kernel(const float *__restrict__ feat, const float *__restrict__ W,
                                                       const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out,
                                                       int M, size_t plane_rows) {
  __shared__ float f[4][72];
  const int r0 = blockIdx.x * 4;
  for (int i = threadIdx.x; i < 4 * 70; i += blockDim.x) {
    int rr = i / 70, kk = i % 70;
    f[rr][kk] = (r0 + rr < M) ? feat[(size_t)(r0 + rr) * 70 + kk] : 0.f;
  }
  __syncthreads();
  const int n = threadIdx.x;
  float a[4] = {bias[n], bias[n], bias[n], bias[n]};
  for (int k = 0; k < 70; k++) {
    float w = __ldg(W + (size_t)k * 128 + n);
#pragma unroll
    for (int rr = 0; rr < 4; rr++) a[rr] = fmaf(w, f[rr][k], a[rr]);
  }
  const size_t plane = plane_rows * 128;
  for (int rr = 0; rr < 4; rr++) {
    if (r0 + rr >= M) break;
    float v = a[rr] < 0.f ? 0.f : a[rr];
    __nv_bfloat16 t0, t1, t2;
    split_b3(v, t0, t1, t2);
    size_t o = (size_t)(r0 + rr) * 128 + n;
    out[o] = t0; out[plane + o] = t1;
    if (kConvTerms == 3) out[2 * plane + o] = t2;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------ host state
struct pnb_tc_state {
  // split activations
  // Multi-hop buffers: rows are (hop slot, stream).  The conv layers carry no recurrence, so they run once per
  // call over all F hops (M = F S rows); a tap is the same buffer read `tap` slots later.  Slots 0..3 (0..1) hold
  // the last hops of the previous call.
  __nv_bfloat16 *fc_all = nullptr;   // [terms][(Fmax+4) S][128]  fc outputs, bf16 terms
  __nv_bfloat16 *c1_all = nullptr;   // [terms][(Fmax+2) S][512]  conv1 outputs
  __half *c2_h = nullptr;            // [2][Fmax S][512]          conv2 outputs, fp16 terms
  __half *h_all[5] = {};             // [2][(Fmax+1) S][H]        GRU states as fp16 terms, slot t+1 = after hop t,
                                     //                           slot 0 = carried from the previous call
  // packed weights
  __nv_bfloat16 *w_conv1 = nullptr, *w_conv2 = nullptr;  // [3][512][K]
  __half *w_gru[5] = {}, *u_gru[5] = {};                 // [2][tiles*192][K]
  float scale_gru[5] = {};                               // 2^-(10 + e)
  __half *w_gb = nullptr, *w_rb = nullptr;               // [2][48][2560], [2][48][128] (34 rows used)
  float scale_gb = 0.f, scale_rb = 0.f;
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
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
  out.assign((size_t)kConvTerms * N * K, __float2bfloat16(0.f));
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++) {
      float w = W[(size_t)k * N + n];
      __nv_bfloat16 t0 = __float2bfloat16_rn(w);
      float r = w - __bfloat162float(t0);
      __nv_bfloat16 t1 = __float2bfloat16_rn(r);
      r -= __bfloat162float(t1);
      __nv_bfloat16 t2 = __float2bfloat16_rn(r);
      out[(size_t)n * K + k] = t0;
      out[((size_t)N + n) * K + k] = t1;
      if (kConvTerms == 3) out[((size_t)2 * N + n) * K + k] = t2;
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
  pnb_tc_state *t = new pnb_tc_state();
  e->tc = t;
  const size_t S = e->S;
  const size_t Fm = e->Fmax;
  TCK(dev_zeros(&t->fc_all, kConvTerms * (Fm + 4) * S * 128));
  TCK(dev_zeros(&t->c1_all, kConvTerms * (Fm + 2) * S * 512));
  TCK(dev_zeros(&t->c2_h, 2 * Fm * S * 512));
  for (int i = 0; i < 5; i++) TCK(dev_zeros(&t->h_all[i], 2 * (Fm + 1) * S * e->gru[i].H));
  {
    std::vector<__nv_bfloat16> p;
    pack_dense_b3(model->conv1->input_weights, 640, 512, p);
    TCK(dev_upload(&t->w_conv1, p));
    pack_dense_b3(model->conv2->input_weights, 1536, 512, p);
    TCK(dev_upload(&t->w_conv2, p));
  }
  const pnb_gru_layer *g5[5] = {model->gru1, model->gru2, model->gru3, model->gru_gb, model->gru_rb};
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
  TCK(cudaFuncSetAttribute(tc_gemm_kernel<2, 2, GRU_BN, GRU_STAGES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)tc_smem_bytes<2, 2, GRU_BN, GRU_STAGES>()));
  TCK(cudaFuncSetAttribute(tc_gemm_kernel<kConvTerms, kConvTerms, DENSE_BN, DENSE_STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)tc_smem_bytes<kConvTerms, kConvTerms, DENSE_BN, DENSE_STAGES>()));
  TCK(cudaFuncSetAttribute(tc_gemm_kernel<2, 2, SMALL_BN, SMALL_STAGES, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)tc_smem_bytes<2, 2, SMALL_BN, SMALL_STAGES>()));
  return PNB_OK;
}

void tc_release(pnb_engine *e) {
  pnb_tc_state *t = e->tc;
  if (!t) return;
  void *ptrs[] = {t->fc_all, t->c1_all, t->c2_h, t->w_conv1, t->w_conv2, t->w_gb, t->w_rb};
  for (void *p : ptrs) if (p) cudaFree(p);
  for (int i = 0; i < 5; i++) {
    if (t->h_all[i]) cudaFree(t->h_all[i]);
    if (t->w_gru[i]) cudaFree(t->w_gru[i]);
    if (t->u_gru[i]) cudaFree(t->u_gru[i]);
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
  return PNB_OK;
}

int tc_launches_per_step(const pnb_engine *) { return 5; }
int tc_launches_per_call(const pnb_engine *) { return 3 + 2; }  // kernels only; the slot carry uses copy engines

template <int NA, int NB, int BN, int STAGES, bool GRU>
static void tc_launch(pnb_engine *e, TcArgs &a, int rows, int tiles_n, cudaStream_t st) {
  // timing diagnostics only (profiles/README.md): with PNB_TC_DEBUG set the network's results are garbage
  static const int dbg = [] {
    const char *v = getenv("PNB_TC_DEBUG");
    const int d = v ? atoi(v) : 0;
    if (d) fprintf(stderr, "percepnet_b200: PNB_TC_DEBUG=%d -- network outputs are INVALID (timing diagnostics)\n", d);
    return d;
  }();
  a.debug = dbg;
  a.M = rows;
  a.tiles_m = (rows + TM - 1) / TM;
  a.tiles_n = tiles_n;
  int pairs = ((a.tiles_m + 1) / 2) * a.tiles_n;
  int clusters = pairs < e->sm_count / 2 ? pairs : e->sm_count / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = tc_smem_bytes<NA, NB, BN, STAGES>();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, tc_gemm_kernel<NA, NB, BN, STAGES, GRU>, a);
}

static TcSeg mkseg(int a_map, int b_map, int K, int a_k0, int b_k0, int rec, int a_row0, int a_term_rows) {
  TcSeg s;
  s.a_map = a_map; s.b_map = b_map; s.k_blocks = K / BK; s.a_k0 = a_k0; s.b_k0 = b_k0; s.recurrent = rec;
  s.a_row0 = a_row0; s.a_term_rows = a_term_rows;
  return s;
}

// The non-recurrent front of the network for all F hops of the call at once (rnn.cpp:50-52 on F S rows):
// fc -> conv1 -> conv2.  Returns the number of launches.
int tc_begin_call(pnb_engine *e, int F, cudaStream_t st) {
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax;
  const float *tbl = e->tansig();
  const int rows = F * S;
  int n = 0;
  // fc (fp32 FMA) for every hop -> bf16 terms at slots 4 .. F+3
  {
    ProfScope ps(e, PNB_K_TC_AUX, st);
    fc_split_kernel<<<(rows + 3) / 4, 128, 0, st>>>(e->d_feat, e->fc.W, e->fc.b, t->fc_all + (size_t)4 * S * 128, rows,
                                                    (size_t)(Fm + 4) * S);
    n++;
  }
  TcArgs a;
  // conv1: tap q of hop t reads fc slot t + q (oldest first, nnet.cpp:182-200); output -> c1 slots 2 .. F+1
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_fc_all;
  a.maps[1] = t->m_wconv1;
  for (int q = 0; q < 5; q++) a.seg[q] = mkseg(0, 1, 128, 0, q * 128, 0, q * S, (Fm + 4) * S);
  a.n_seg = 5; a.b_term_rows = 512; a.fmt = 1; a.out_scale = 1.f; a.tansig = tbl;
  a.bias = e->conv1.b; a.act = e->act_conv1; a.N = 512;
  a.out_b = t->c1_all; a.out_row0 = 2 * S; a.out_plane_rows = (Fm + 2) * S;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    tc_launch<kConvTerms, kConvTerms, DENSE_BN, DENSE_STAGES, false>(e, a, rows, 512 / DENSE_BN, st);
    n++;
  }
  // conv2: three taps -> tanh -> fp16 terms for every hop (gru1 / gru_rb / fc_gb inputs); fp32 copy of the last hop
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_c1_all;
  a.maps[1] = t->m_wconv2;
  for (int q = 0; q < 3; q++) a.seg[q] = mkseg(0, 1, 512, 0, q * 512, 0, q * S, (Fm + 2) * S);
  a.n_seg = 3; a.b_term_rows = 512; a.fmt = 1; a.out_scale = 1.f; a.tansig = tbl;
  a.bias = e->conv2.b; a.act = e->act_conv2; a.N = 512;
  a.out_f32 = e->c2; a.ldc = 512; a.f32_row0 = (F - 1) * S;
  a.out_h = t->c2_h; a.out_row0 = 0; a.out_plane_rows = Fm * S;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    tc_launch<kConvTerms, kConvTerms, DENSE_BN, DENSE_STAGES, false>(e, a, rows, 512 / DENSE_BN, st);
    n++;
  }
  // carry the last 4 (2) hop slots to the front for the next call; slot by slot, ascending (ranges may overlap)
  for (int tm = 0; tm < kConvTerms; tm++) {
    for (int i = 0; i < 4; i++) {
      __nv_bfloat16 *base = t->fc_all + (size_t)tm * (Fm + 4) * S * 128;
      cudaMemcpyAsync(base + (size_t)i * S * 128, base + (size_t)(F + i) * S * 128, (size_t)S * 128 * 2,
                      cudaMemcpyDeviceToDevice, st);
    }
    for (int i = 0; i < 2; i++) {
      __nv_bfloat16 *base = t->c1_all + (size_t)tm * (Fm + 2) * S * 512;
      cudaMemcpyAsync(base + (size_t)i * S * 512, base + (size_t)(F + i) * S * 512, (size_t)S * 512 * 2,
                      cudaMemcpyDeviceToDevice, st);
    }
  }
  return n;
}

// The recurrent part of one hop: the five GRUs (rnn.cpp:58-71).  State slot t holds the state BEFORE hop t.
int tc_step(pnb_engine *e, int tstep, cudaStream_t st) {
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax;
  const float *tbl = e->tansig();
  int n = 0;
  TcArgs a;
  const int c2_row0 = tstep * S, c2_terms = Fm * S;
  const int h_terms = (Fm + 1) * S, h_prev = tstep * S, h_next = (tstep + 1) * S;
  // each GRU consumes the freshly written state (slot t+1) of the layer below and its own previous state (slot t)
  for (int li = 0; li < 5; li++) {
    const int H = e->gru[li].H, p = e->par[li];
    memset(&a, 0, sizeof a);
    int ns = 0;
    if (li == 0) {
      a.maps[0] = t->m_c2;
      a.seg[ns++] = mkseg(0, 2, 512, 0, 0, 0, c2_row0, c2_terms);
    } else if (li < 4) {
      a.maps[0] = t->m_h[li - 1];
      a.seg[ns++] = mkseg(0, 2, 512, 0, 0, 0, h_next, h_terms);
    } else {  // gru_rb input = [gru3 state, conv2 out]
      a.maps[0] = t->m_h[2];
      a.maps[4] = t->m_c2;
      a.seg[ns++] = mkseg(0, 2, 512, 0, 0, 0, h_next, h_terms);
      a.seg[ns++] = mkseg(4, 2, 512, 0, 512, 0, c2_row0, c2_terms);
    }
    a.maps[1] = t->m_h[li];
    a.maps[2] = t->m_w[li];
    a.maps[3] = t->m_u[li];
    a.seg[ns++] = mkseg(1, 3, H, 0, 0, 1, h_prev, h_terms);
    a.n_seg = ns; a.b_term_rows = (H / HT) * GRU_BN; a.fmt = 0;
    a.out_scale = t->scale_gru[li]; a.tansig = tbl; a.bias = e->gru[li].b; a.H = H;
    a.h_old = e->h[li][p]; a.h_new = e->h[li][p ^ 1];
    a.h_new_h = t->h_all[li]; a.out_row0 = h_next; a.out_plane_rows = h_terms;
    {
      ProfScope ps(e, PNB_K_TC_GEMM, st);
      tc_launch<2, 2, GRU_BN, GRU_STAGES, true>(e, a, S, H / HT, st);
      n++;
    }
    e->par[li] ^= 1;
  }
  return n;
}

// The two 34-wide output layers for all F hops at once (rnn.cpp:73-80; N padded to 48, fp16 two-term split,
// sigmoid epilogue), then the carry of the last state slot to slot 0 for the next call.
int tc_end_call(pnb_engine *e, int F, cudaStream_t st) {
  pnb_tc_state *t = e->tc;
  const int S = e->S, Fm = e->Fmax;
  const float *tbl = e->tansig();
  const int rows = F * S, h_terms = (Fm + 1) * S;
  int n = 0;
  TcArgs a;
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_c2;
  for (int q = 0; q < 4; q++) a.maps[1 + q] = t->m_h[q];
  a.maps[5] = t->m_wgb;
  a.seg[0] = mkseg(0, 5, 512, 0, 0, 0, 0, Fm * S);
  for (int q = 1; q < 5; q++) a.seg[q] = mkseg(q, 5, 512, 0, q * 512, 0, S, h_terms);  // state after hop t = slot t+1
  a.n_seg = 5; a.b_term_rows = SMALL_BN; a.fmt = 0; a.out_scale = t->scale_gb; a.tansig = tbl;
  a.bias = e->fc_gb.b; a.act = e->act_gb; a.N = 34; a.ldc = 68; a.out_f32 = e->d_gr;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    tc_launch<2, 2, SMALL_BN, SMALL_STAGES, false>(e, a, rows, 1, st);
    n++;
  }
  memset(&a, 0, sizeof a);
  a.maps[0] = t->m_h[4];
  a.maps[1] = t->m_wrb;
  a.seg[0] = mkseg(0, 1, 128, 0, 0, 0, S, h_terms);
  a.n_seg = 1; a.b_term_rows = SMALL_BN; a.fmt = 0; a.out_scale = t->scale_rb; a.tansig = tbl;
  a.bias = e->fc_rb.b; a.act = e->act_rb; a.N = 34; a.ldc = 68; a.out_f32 = e->d_gr + 34;
  {
    ProfScope ps(e, PNB_K_TC_GEMM, st);
    tc_launch<2, 2, SMALL_BN, SMALL_STAGES, false>(e, a, rows, 1, st);
    n++;
  }
  for (int li = 0; li < 5; li++) {
    const size_t H = e->gru[li].H;
    for (int tm = 0; tm < 2; tm++) {
      __half *base = t->h_all[li] + (size_t)tm * h_terms * H;
      cudaMemcpyAsync(base, base + (size_t)F * S * H, (size_t)S * H * 2, cudaMemcpyDeviceToDevice, st);
    }
  }
  return n;
}
