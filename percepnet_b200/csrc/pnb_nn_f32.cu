// pnb_nn_f32.cu -- the gain network in fp32 FMA (BASELINE.json config 2), batched over streams.
//
// The reference evaluates the network one stream at a time as 35 GEMVs per hop
// (/root/reference/src/nnet.cpp:59-72, vec.h:102-135, rnn.cpp:42-81) and re-reads all 31.9 MB of
// weights for every hop.  Here the rows of every contraction are the concurrent streams, so each
// weight tile is reused across 128 streams from shared memory; activations use the reference's
// tansig table approximation (vec.h:53-76).
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

constexpr int BM = 128, BN = 64, BK = 16, APAD = 4;

template <bool VEC_B>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

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
      float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
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
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int c = n0 + tx * 4 + j;
      if (c >= g.N) continue;
      float v = acc[i][j];
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

// GRU gates (nnet.cpp:120-180 with reset_after):
//   z = sig(b_z + b'_z + zsum)   r = sig(b_r + b'_r + rsum)
//   n = tanh(b_n + (b'_n + nh) * r + nx)   h' = z h + (1 - z) n
__global__ void gru_gates_kernel(GruGateArgs g) {
  const int H = g.H;
  const size_t total = (size_t)g.M * H;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int m = (int)(idx / H), j = (int)(idx % H);
    const float *b = g.bias;
    float zs = g.zr[(size_t)m * 2 * H + j], rs = g.zr[(size_t)m * 2 * H + H + j];
    float z = sigmoid_approx((__ldg(b + j) + __ldg(b + 3 * H + j)) + zs, g.tansig);
    float r = sigmoid_approx((__ldg(b + H + j) + __ldg(b + 4 * H + j)) + rs, g.tansig);
    float tmp = __ldg(b + 5 * H + j) + g.nh[idx];
    float c = __ldg(b + 2 * H + j) + tmp * r;
    c = c + g.nx[idx];
    float n = tansig_approx(c, g.tansig);
    float h = g.h_old[idx];
    g.h_new[idx] = z * h + (1.f - z) * n;
  }
}

int launch_gru_gates(const GruGateArgs &g, cudaStream_t st) {
  size_t total = (size_t)g.M * g.H;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  gru_gates_kernel<<<blocks, 256, 0, st>>>(g);
  return 1;
}

}  // namespace pnb
