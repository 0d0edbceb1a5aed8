// pnb_nn_f32.cu -- the gain network in fp32 FMA (BASELINE.json config 2), batched over streams.
//
// The reference evaluates the network one stream at a time as 35 GEMVs per hop
// (/root/reference/src/nnet.cpp:59-72, vec.h:102-135, rnn.cpp:42-81) and re-reads all 31.9 MB of
// weights for every hop.  Here the rows of every contraction are the concurrent streams, so each
// weight tile is reused across 128 streams from shared memory; activations use the reference's
// tansig table approximation (vec.h:53-76).  Per chunk of hops: fc, conv1, conv2 over all rows of the chunk
// (gemm_f32_kernel), the five GRUs of all its hops in one persistent launch (gru_chain_f32_kernel), the output layers,
// and a carry; the schedule is nn_chunk_f32 in pnb_engine.cu.
#include "pnb_kernels.h"
#include "../../include/pnb_nnet_layout.h"

namespace pnb {

__device__ __forceinline__ float tansig_approx(float x, const float *__restrict__ tbl) {  // vec.h:53-71
  float sign = 1.f;
  if (x < 0.f) { x = -x; sign = -1.f; }
  float fi = floorf(.5f + 25.f * x);
  // the reference converts with x86 cvttss2si, which yields INT_MIN when out of range or NaN (then the
  // clamp below lands on 0, not 200): keep that for |x| >= 8.6e7 so both sides leave the table the same way
  int i = (fi < 2147483648.f) ? (int)fi : (int)0x80000000;
  i = i < 200 ? i : 200;
  i = i > 0 ? i : 0;
  x -= .04f * i;
  float y = __ldg(tbl + i);
  float dy = 1.f - y * y;
  y = y + x * dy * (1.f - y * x);
  return sign * y;
}
__device__ __forceinline__ float sigmoid_approx(float x, const float *__restrict__ tbl) {  // vec.h:73-76
  return .5f + .5f * tansig_approx(.5f * x, tbl);
}
__device__ __forceinline__ float activate(float v, int act, const float *__restrict__ tbl) {
  if (act == PNB_ACT_SIGMOID) return sigmoid_approx(v, tbl);
  if (act == PNB_ACT_TANH) return tansig_approx(v, tbl);
  if (act == PNB_ACT_RELU) return v < 0.f ? 0.f : v;
  return v;
}

// Packed fp32 pairs (fma.rn.f32x2: two IEEE fp32 FMAs per instruction, bit for bit what two fmaf give).  A scalar FFMA
// reads three registers from two register banks, so unless the allocator happens to put the accumulator and the B value
// on opposite banks it takes two cycles -- ncu showed "dispatch stall" as the largest stall of the scalar version, with
// half of its FFMAs conflicting.  The packed form reads aligned pairs (one register per bank each): every instruction
// costs the same two cycles for two FMAs, and the loop needs half the issue slots.
typedef unsigned long long f32x2;
__device__ __forceinline__ void ffma2(f32x2 &d, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
__device__ __forceinline__ f32x2 pack2(float x, float y) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

constexpr int BM = 128, BN = 64, BK = 16, APAD = 4;

template <bool VEC_B>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  f32x2 acc[8][2];  // 8 rows x 2 pairs of columns
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i][0] = 0ull; acc[i][1] = 0ull; }

  // global->register staging: A 2 x float4 per thread, B 1 x float4 per thread
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;  // rows a_row and a_row+64
  const int b_k = tid >> 4, b_n = (tid & 15) * 4;
  float4 ra0, ra1, rb;

  int seg = 0, koff = 0;
  int total_tiles = 0;
  for (int i = 0; i < g.n_seg; i++) total_tiles += g.seg[i].K / BK;

  auto fetch = [&](int sg, int k0) {
    const GemmSeg &S = g.seg[sg];
    int r0 = m0 + a_row, r1 = r0 + 64;
    ra0 = r0 < g.M ? *reinterpret_cast<const float4 *>(S.A + (size_t)r0 * S.lda + k0 + a_kq)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    ra1 = r1 < g.M ? *reinterpret_cast<const float4 *>(S.A + (size_t)r1 * S.lda + k0 + a_kq)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    const float *bp = S.B + (size_t)(k0 + b_k) * S.ldb + n0 + b_n;
    if (VEC_B) {
      rb = (n0 + b_n < g.N) ? __ldg(reinterpret_cast<const float4 *>(bp)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      rb.x = (n0 + b_n + 0 < g.N) ? __ldg(bp + 0) : 0.f;
      rb.y = (n0 + b_n + 1 < g.N) ? __ldg(bp + 1) : 0.f;
      rb.z = (n0 + b_n + 2 < g.N) ? __ldg(bp + 2) : 0.f;
      rb.w = (n0 + b_n + 3 < g.N) ? __ldg(bp + 3) : 0.f;
    }
  };
  auto stash = [&](int buf) {
    As[buf][a_kq + 0][a_row] = ra0.x; As[buf][a_kq + 1][a_row] = ra0.y;
    As[buf][a_kq + 2][a_row] = ra0.z; As[buf][a_kq + 3][a_row] = ra0.w;
    As[buf][a_kq + 0][a_row + 64] = ra1.x; As[buf][a_kq + 1][a_row + 64] = ra1.y;
    As[buf][a_kq + 2][a_row + 64] = ra1.z; As[buf][a_kq + 3][a_row + 64] = ra1.w;
    *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n]) = rb;
  };

  fetch(0, 0);
  stash(0);
  __syncthreads();
  for (int tile = 0; tile < total_tiles; tile++) {
    const int buf = tile & 1;
    // advance (seg, koff) to the next tile and prefetch it
    int nseg = seg, nk = koff + BK;
    if (nk >= g.seg[seg].K) { nseg = seg + 1; nk = 0; }
    const bool more = tile + 1 < total_tiles;
    if (more) fetch(nseg, nk);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
      const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][k][tx * 4]);  // columns {0,1} and {2,3} as pairs
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const f32x2 aa = pack2(av[i], av[i]);
        ffma2(acc[i][0], aa, b.x);
        ffma2(acc[i][1], aa, b.y);
      }
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    seg = nseg;
    koff = nk;
  }

#pragma unroll
  for (int i = 0; i < 8; i++) {
    int r = m0 + ty * 8 + i;
    if (r >= g.M) continue;
    const float2 p0 = unpack2(acc[i][0]), p1 = unpack2(acc[i][1]);
    const float accv[4] = {p0.x, p0.y, p1.x, p1.y};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int c = n0 + tx * 4 + j;
      if (c >= g.N) continue;
      float v = accv[j];
      if (g.bias) v = activate(v + __ldg(g.bias + c), g.act, g.tansig);
      g.C[(size_t)r * g.ldc + c] = v;
    }
  }
}

int launch_gemm_f32(const GemmArgs &g, cudaStream_t st) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  bool vec = true;
  for (int i = 0; i < g.n_seg; i++)
    if ((g.seg[i].ldb & 3) || (reinterpret_cast<uintptr_t>(g.seg[i].B) & 15)) vec = false;
  if (vec) gemm_f32_kernel<true><<<grid, 256, 0, st>>>(g);
  else gemm_f32_kernel<false><<<grid, 256, 0, st>>>(g);
  return 1;
}

// fc 70 -> 128 (+relu): one thread per (stream, output); the 70-float feature row is broadcast, the
// weight row W[j*128 + i] is coalesced over i.
__global__ void __launch_bounds__(128) fc_f32_kernel(const float *__restrict__ feat, const float *__restrict__ W,
                                                     const float *__restrict__ bias, float *__restrict__ out,
                                                     int M, int K, int N) {
  __shared__ float f[4][72];
  const int r0 = blockIdx.x * 4;
  for (int i = threadIdx.x; i < 4 * K; i += blockDim.x) {
    int rr = i / K, kk = i % K;
    f[rr][kk] = (r0 + rr < M) ? feat[(size_t)(r0 + rr) * K + kk] : 0.f;
  }
  __syncthreads();
  const int n = threadIdx.x;
  if (n >= N) return;
  float a0 = bias[n], a1 = a0, a2 = a0, a3 = a0;
  for (int k = 0; k < K; k++) {
    float w = __ldg(W + (size_t)k * N + n);
    a0 = fmaf(w, f[0][k], a0);
    a1 = fmaf(w, f[1][k], a1);
    a2 = fmaf(w, f[2][k], a2);
    a3 = fmaf(w, f[3][k], a3);
  }
  float a[4] = {a0, a1, a2, a3};
  for (int rr = 0; rr < 4; rr++)
    if (r0 + rr < M) out[(size_t)(r0 + rr) * N + n] = a[rr] < 0.f ? 0.f : a[rr];
}

int launch_fc_f32(const float *feat, const float *W, const float *bias, float *out, int M, int K, int N,
                  cudaStream_t st) {
  fc_f32_kernel<<<(M + 3) / 4, 128, 0, st>>>(feat, W, bias, out, M, K, N);
  return 1;
}

// ------------------------------------------------------------------------------------------------------------------
// The five GRUs of a chunk of hops in ONE persistent fp32 launch (BASELINE.json config 2: "persistent-kernel GRU, fp32").
//
// Unit of work = (hop t, layer l, block of BM streams, tile of 32 hidden units): the four gate sums z, r, W_n x, U_n h of
// 32 hidden units for BM streams (nnet.cpp:120-180 keeps the last two apart), as one register-tiled FMA contraction over
// [x ; h], then the gate math and the state update in the same thread -- no pre-activation ever goes to memory.
// State slot t of h_all[l] is the state BEFORE hop t of the chunk, slot t+1 the state after it: a layer reads slot t of
// its own buffer and slot t+1 of the layer below (the freshly updated state, rnn.cpp:58-71), so nothing is overwritten
// inside a launch.  A unit depends only on units of the same stream block: all tiles of (t, l-1) and of (t-1, l).  Each
// (layer, stream block) has a completion counter (red.release / ld.acquire); units are walked in anti-diagonal order
// (t + l ascending), where every dependency lies on an earlier diagonal, by a grid that is fully co-resident -- so the
// waits are short, there is no grid barrier and no launch boundary between layers or hops.
// The accumulation order of every sum (ascending k, x segments first, then h, one FMA per term) is the one the
// per-layer kernels above use.
// ------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int CH_BK = 32, CH_HT = 32, CH_BN = 3 * CH_HT, CH_STAGES = 3, CH_THREADS = 256;  // 32-k tiles: one block barrier per 32 k
constexpr int CH_APAD = CH_BK + 4;  // row stride of the A tile in floats: 16-byte aligned rows, conflict-free 128-bit reads

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int RT>  // rows per thread: 8 (128-stream blocks) or 1 (16-stream blocks, for a handful of streams)
struct ChainSmem {
  static constexpr int BM = 16 * RT;
  float A[CH_STAGES][BM][CH_APAD];
  float B[CH_STAGES][CH_BK][CH_BN];
};

// one k-tile of FMAs: acc_z/acc_r/acc_n[i] (pairs of hidden units) += A[row i][k] * B[k][gate][pair]
template <int RT>
__device__ __forceinline__ void chain_tile_fma(const float (*As)[CH_APAD], const float (*Bs)[CH_BN], int ty, int tx,
                                               f32x2 (&az)[RT], f32x2 (&ar)[RT], f32x2 (&an)[RT]) {
#pragma unroll
  for (int kk = 0; kk < CH_BK; kk += 4) {
    float4 a[RT];
#pragma unroll
    for (int i = 0; i < RT; i++) a[i] = *reinterpret_cast<const float4 *>(&As[ty + 16 * i][kk]);  // rows ty, ty+16, ..: a warp's two rows are neighbours (different banks)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const f32x2 bz = *reinterpret_cast<const f32x2 *>(&Bs[kk + q][2 * tx]);
      const f32x2 br = *reinterpret_cast<const f32x2 *>(&Bs[kk + q][CH_HT + 2 * tx]);
      const f32x2 bn = *reinterpret_cast<const f32x2 *>(&Bs[kk + q][2 * CH_HT + 2 * tx]);
#pragma unroll
      for (int i = 0; i < RT; i++) {
        const float av = q == 0 ? a[i].x : q == 1 ? a[i].y : q == 2 ? a[i].z : a[i].w;
        const f32x2 aa = pack2(av, av);
        ffma2(az[i], aa, bz);
        ffma2(ar[i], aa, br);
        ffma2(an[i], aa, bn);
      }
    }
  }
}

template <int RT>
__global__ void __launch_bounds__(CH_THREADS, RT == 8 ? 2 : 4) gru_chain_f32_kernel(const __grid_constant__ F32ChainArgs args) {
  using SM = ChainSmem<RT>;
  constexpr int BM = SM::BM;
  extern __shared__ __align__(16) unsigned char chain_smem_raw[];
  SM &sm = *reinterpret_cast<SM *>(chain_smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int S = args.S;

  for (int u = blockIdx.x; u < args.n_units; u += gridDim.x) {
    // ---- which unit: anti-diagonal order, the (hop, layer) entries and their first units come from the host
    int e = 0;
    while (u >= args.lh_unit0[e + 1]) e++;
    const int t = args.lh[e] >> 3, l = args.lh[e] & 7;
    const F32ChainLayer &L = args.L[l];
    const int tiles = L.H / CH_HT;
    const int local = u - args.lh_unit0[e];
    const int rb = local / tiles, j0 = (local - rb * tiles) * CH_HT, m0 = rb * BM;

    // ---- dependencies (one thread polls; the barrier publishes the result to the block)
    if (tid == 0) {
      unsigned spins = 0;
      if (L.dep >= 0) {
        const unsigned want = (unsigned)(t + 1) * (unsigned)(args.L[L.dep].H / CH_HT);
        while (ld_acquire_gpu_u32(args.cnt + L.dep * args.n_rb + rb) < want)
          if (++spins > (1u << 26)) __trap();  // a broken schedule becomes a launch error, not a hang
      }
      if (t > 0) {
        const unsigned want = (unsigned)t * (unsigned)tiles;
        while (ld_acquire_gpu_u32(args.cnt + l * args.n_rb + rb) < want)
          if (++spins > (1u << 26)) __trap();
      }
    }
    __syncthreads();

    // ---- the contraction over [x segments ; h]
    f32x2 az[RT], ar[RT], anx[RT], anh[RT];  // pairs of neighbouring hidden units
#pragma unroll
    for (int i = 0; i < RT; i++) { az[i] = 0ull; ar[i] = 0ull; anx[i] = 0ull; anh[i] = 0ull; }

    const int n_seg = L.n_x + 1;
    int kt_total = 0;
    for (int sg = 0; sg < n_seg; sg++) kt_total += (sg < L.n_x ? L.x_K[sg] : L.H) / CH_BK;

    // issue the loads of k-tile `kt` (global tile index over all segments) into stage kt % CH_STAGES
    auto issue = [&](int kt) {
      int sg = 0, k0 = kt * CH_BK;
      while (sg < L.n_x && k0 >= L.x_K[sg]) { k0 -= L.x_K[sg]; sg++; }
      const bool rec = sg == L.n_x;
      const float *Ab = rec ? L.h + (size_t)t * S * L.H : L.x[sg] + (size_t)(t + L.x_slot1[sg]) * L.x_slot_stride[sg];
      const int lda = rec ? L.H : L.x_ld[sg];
      const float *Bb = rec ? L.U : L.W + (size_t)L.w_row0[sg] * L.ldw;
      const int st = kt % CH_STAGES;
      // A: BM rows x CH_BK/4 chunks of 16 bytes
      constexpr int CPR = CH_BK / 4;
      for (int c = tid; c < BM * CPR; c += CH_THREADS) {
        const int r = c / CPR, q = c - r * CPR;
        const bool ok = m0 + r < S;
        cp_async16(&sm.A[st][r][q * 4], Ab + (size_t)(ok ? m0 + r : 0) * lda + k0 + q * 4, ok);
      }
      // B: CH_BK k-rows x 3 gates x 8 chunks
      for (int c = tid; c < CH_BK * 24; c += CH_THREADS) {
        const int kr = c / 24, rem = c - kr * 24, g = rem >> 3, q = rem & 7;
        cp_async16(&sm.B[st][kr][g * CH_HT + q * 4], Bb + (size_t)(k0 + kr) * L.ldw + g * L.H + j0 + q * 4, true);
      }
    };
    for (int p = 0; p < CH_STAGES - 1; p++) {
      if (p < kt_total) issue(p);
      cp_async_commit();
    }
    const int kt_x = kt_total - L.H / CH_BK;  // tiles before the recurrent segment
    for (int kt = 0; kt < kt_total; kt++) {
      cp_async_wait<CH_STAGES - 2>();
      __syncthreads();  // tile kt has landed for every thread; the stage read in iteration kt-1 is free again
      if (kt + CH_STAGES - 1 < kt_total) issue(kt + CH_STAGES - 1);
      cp_async_commit();
      const int st = kt % CH_STAGES;
      if (kt < kt_x) chain_tile_fma<RT>(sm.A[st], sm.B[st], ty, tx, az, ar, anx);
      else chain_tile_fma<RT>(sm.A[st], sm.B[st], ty, tx, az, ar, anh);
    }
    cp_async_wait<0>();

    // ---- gates and state update (nnet.cpp:120-180 with reset_after: z, r from the summed sums, n = tanh(b_n + (b'_n + U_n h) r + W_n x))
    {
      const float *b = L.bias;
      const int H = L.H;
      const float *h_old = L.h + (size_t)t * S * H;
      float *h_new = L.h + (size_t)(t + 1) * S * H;
      const int j = j0 + 2 * tx;
      float bz[2], br[2], bn[2], bnh[2];
#pragma unroll
      for (int c = 0; c < 2; c++) {
        bz[c] = __ldg(b + j + c) + __ldg(b + 3 * H + j + c);
        br[c] = __ldg(b + H + j + c) + __ldg(b + 4 * H + j + c);
        bn[c] = __ldg(b + 2 * H + j + c);
        bnh[c] = __ldg(b + 5 * H + j + c);
      }
#pragma unroll
      for (int i = 0; i < RT; i++) {
        const int m = m0 + ty + 16 * i;
        if (m >= S) continue;
        const float2 ho = __ldcg(reinterpret_cast<const float2 *>(h_old + (size_t)m * H + j));
        const float2 sz = unpack2(az[i]), sr = unpack2(ar[i]), snx = unpack2(anx[i]), snh = unpack2(anh[i]);
        float out[2];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const float z = sigmoid_approx(bz[c] + (c == 0 ? sz.x : sz.y), args.tansig);
          const float r = sigmoid_approx(br[c] + (c == 0 ? sr.x : sr.y), args.tansig);
          const float tmp = bnh[c] + (c == 0 ? snh.x : snh.y);
          float cc = bn[c] + tmp * r;
          cc = cc + (c == 0 ? snx.x : snx.y);
          const float n = tansig_approx(cc, args.tansig);
          const float h = c == 0 ? ho.x : ho.y;
          out[c] = z * h + (1.f - z) * n;
        }
        *reinterpret_cast<float2 *>(h_new + (size_t)m * H + j) = make_float2(out[0], out[1]);
      }
    }
    __syncthreads();  // every thread's stores are issued (and the tile buffers are free for the next unit)
    if (tid == 0) {
      __threadfence();
      red_release_gpu_add_u32(args.cnt + l * args.n_rb + rb, 1u);
    }
  }
}

// End of a chunk of n hops: the last 4 fc / 2 conv1 slots and the last state slot move to the front, conv2's output of
// the last hop is kept for the tap, the chain's counters return to zero.  A thread owns the same element of every slot.
__global__ void f32_carry_kernel(F32CarryArgs a) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (int sg = 0; sg < a.n_seg; sg++) {
    const F32CarrySeg &g = a.seg[sg];
    for (size_t i = tid; i < g.slot4; i += stride)
      for (int q = 0; q < g.n_slots; q++) g.dst[(size_t)q * g.slot4 + i] = g.src[(size_t)q * g.slot4 + i];
  }
  for (size_t i = tid; i < (size_t)a.n_cnt; i += stride) a.cnt[i] = 0u;
}
}  // namespace

size_t f32_chain_smem(int rt) { return rt == 8 ? sizeof(ChainSmem<8>) : sizeof(ChainSmem<1>); }

int launch_gru_chain_f32(const F32ChainArgs &a, int rt, int sm_count, cudaStream_t st) {
  static int occ8 = 0, occ1 = 0;
  int &occ = rt == 8 ? occ8 : occ1;
  const size_t smem = f32_chain_smem(rt);
  if (!occ) {
    cudaError_t e1, e2;
    if (rt == 8) {
      e1 = cudaFuncSetAttribute(gru_chain_f32_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gru_chain_f32_kernel<8>, CH_THREADS, smem);
    } else {
      e1 = cudaFuncSetAttribute(gru_chain_f32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gru_chain_f32_kernel<1>, CH_THREADS, smem);
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess || occ < 1) { occ = 0; return -1; }
  }
  // the dependency waits need every block of the grid resident: never more blocks than fit at once
  int grid = occ * sm_count;
  if (grid > a.n_units) grid = a.n_units;
  if (rt == 8) gru_chain_f32_kernel<8><<<grid, CH_THREADS, smem, st>>>(a);
  else gru_chain_f32_kernel<1><<<grid, CH_THREADS, smem, st>>>(a);
  return 1;
}

int launch_f32_carry(const F32CarryArgs &a, cudaStream_t st) {
  f32_carry_kernel<<<296, 256, 0, st>>>(a);
  return 1;
}

}  // namespace pnb
