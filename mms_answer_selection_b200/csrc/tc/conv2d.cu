// 2-D convolution of the CNN that consumes the similarity tensor S in the reference's net
// (examples/trec_qa_w2v_mms/do_trec_qa_clean.py:470-477: Conv 5x5 (mc -> 32) + BN -> AvePool 4 -> TanH -> Conv 5x5
// (32 -> 64) + BN -> AvePool 5 -> TanH), as implicit GEMMs on tcgen05 (TF32 multiply, fp32 accumulate in TMEM).
// Reference layer: ConvolutionLayer (conv_layer.cpp:25-73) over BaseConvolutionLayer's im2col + gemm
// (base_conv_layer.cpp:257-321), stride 1, pad 0, group 1 -- the geometry of the net; other geometries are refused.
//
// The reference materialises the im2col matrix per sample (kh*kw = 25 copies of every input value) and calls one gemm
// per sample.  Here the im2col matrix never exists: its elements are gathered straight from the NCHW tensor into the
// shared-memory operand tile (separable addressing: element (row, k) of the matrix is src[f(row) + g(k)], g from a
// small table), rounded to TF32 on the way, in the canonical 128-byte-swizzled K-major UMMA layout:
//
//   forward  Y[(n,oy,ox)][o] = sum_k A[(n,oy,ox)][k] Wm[o][k] + b[o]     A gathered from x,  k = (c,ky,kx)   M = N*OH*OW
//   dx       Zt_n[k][p] = sum_o Wm[o][k] dY[n][o][p]  (one batched GEMM, operands in place), then col2im in gather
//            form: dX[n][c][y][x] = sum_(ky,kx) Zt_n[(c,ky,kx)][(y-ky, x-kx)]  -- what the reference does per sample
//   dW       dW[o][k]       += sum_(n,p) dY[n][o][p] A[(n,p)][k]         the reduction runs over positions: dY is the
//                                                                         K-major A operand as it lies in memory, the
//                                                                         gathered matrix is the B operand; images are
//                                                                         split over CTAs, red.global.add into dW
// One persistent kernel (warps 0-3 epilogue, warp 4 MMA issue, warps 5-12 operand staging; mbarrier ring; double-
// buffered TMEM accumulator), two modes.  The outputs leave in NCHW directly from the accumulator rows (a warp's 32
// lanes are 32 consecutive positions of one channel plane: 128 contiguous bytes per store).
#include "../mms_common.cuh"
#include "tc_gemm.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kBM = 128;
constexpr int kBK = 32;
constexpr int kEpiWarps = 4;
constexpr int kLoadWarps = 8;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kThreads = (kEpiWarps + 1 + kLoadWarps) * 32;
constexpr int kMaxStages = 6;
constexpr int kMaxKc = 2048;            // entries of the k-offset table

enum { kFwd = 0, kDw = 2 };

struct ConvArgs {
  const float* src;        // gathered tensor: x (fwd, dW) or dY (dx): (Nimg, Cs, Hs, Ws)
  int Cs, Hs, Ws;
  int Hg, Wg;              // grid the gathered rows walk, per image: OH x OW (fwd, dW) or H x W (dx)
  int sgn;                 // +1: (y + ky, x + kx);  -1: (y - ky, x - kx) with zero padding
  int kh, kw, Kc;          // Kc = Cs * kh * kw
  int Nimg;
  const float* dense;      // fwd / dx: weight matrix [Ncols][Kc];  dW: dY (Nimg, Co, P)
  int Ncols;               // fwd: C_out, dx: C_in
  float* out;              // fwd / dx: (Nimg, Ncols, Hg, Wg);  dW: [Co][Kc], accumulated
  const float* bias;       // fwd: [C_out] or null
  int Co;                  // dW: rows of dW (= channels of dY)
  int BN, n_tiles, isplit, stages;
  long long Mtot;          // fwd / dx: Nimg * Hg * Wg
  unsigned total_tiles;
  uint32_t tmem_cols;
};

struct Smem {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};

struct Tile { long long m0; int n0, img0, img1, nk; };

__device__ __forceinline__ Tile decode(const ConvArgs& g, unsigned t, int mode, int sps) {
  Tile tl;
  if (mode != kDw) {
    tl.m0 = (long long)t * kBM; tl.n0 = 0; tl.img0 = tl.img1 = 0; tl.nk = sps;
  } else {
    const int nt = t % g.n_tiles, sp = t / g.n_tiles;
    const int per = (g.Nimg + g.isplit - 1) / g.isplit;
    tl.m0 = 0; tl.n0 = nt * g.BN;
    tl.img0 = min(g.Nimg, sp * per); tl.img1 = min(g.Nimg, tl.img0 + per);
    tl.nk = (tl.img1 - tl.img0) * sps;
  }
  return tl;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv2d_kernel(const ConvArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = g.BN * 128;
  const int stage_bytes = 16384 + b_bytes;
  Smem* sm = reinterpret_cast<Smem*>(ring + g.stages * stage_bytes);
  int* koff = reinterpret_cast<int*>(sm + 1);              // g(k): offset of gathered column k

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const int plane_g = g.Hg * g.Wg;
  // stages per reduction unit: fwd / dx: k blocks of the gathered matrix; dW: position blocks of one image
  const int sps = MODE == kDw ? (plane_g + kBK - 1) / kBK : (g.Kc + kBK - 1) / kBK;
  for (int k = threadIdx.x; k < g.Kc; k += kThreads) {
    const int c = k / (g.kh * g.kw), r = k - c * g.kh * g.kw, ky = r / g.kw, kx = r - ky * g.kw;
    koff[k] = c * g.Hs * g.Ws + g.sgn * (ky * g.Ws + kx);
  }
  if (warp == kEpiWarps) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) { mbar_init(&sm->full[s], kLoadWarps); mbar_init(&sm->empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], kEpiWarps); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  const long long src_img = (long long)g.Cs * g.Hs * g.Ws;

  if (warp > kEpiWarps) {
    // ------------------------------------------------------------ operand staging
    const int tid = threadIdx.x - (kEpiWarps + 1) * 32;
    const int r0 = tid >> 3, c4 = tid & 7;
    const uint32_t soff0 = swz128(r0, c4);                 // chunk j of this thread: + j * 4096 (32 rows further down)
    int it = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const Tile tl = decode(g, t, MODE, sps);
      // per-tile row geometry of the gathered operand: a loader thread owns ONE row (position) of the tile and half of
      // the stage's k columns, so that the 32 lanes of a warp read 32 consecutive positions of one (c, ky, kx) plane --
      // one 128-byte line per load instruction
      const int grow = tid & 127, ghalf = tid >> 7;
      long long gbase = 0; bool gok = false;
      if (MODE != kDw) {
        const long long m = tl.m0 + grow;
        gok = m < g.Mtot;
        const long long n = gok ? m / plane_g : 0;
        const int p = gok ? (int)(m - n * plane_g) : 0;
        const int gy = p / g.Wg, gx = p - gy * g.Wg;
        gbase = n * src_img + (long long)gy * g.Ws + gx;
      }
      for (int i = 0; i < tl.nk; ++i, ++it) {
        const int s = it % g.stages;
        float4 va[4], vb[8];
        if (MODE != kDw) {
          const int k0 = i * kBK + c4 * 4;
          // gathered A: 4 chunks of 4 consecutive gathered columns for this thread's row
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int kc = i * kBK + (ghalf * 4 + cc) * 4;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (gok && kc + e < g.Kc) ? __ldg(g.src + gbase + koff[kc + e]) : 0.f;
            va[cc] = make_float4(v[0], v[1], v[2], v[3]);
          }
          // dense B: weight matrix rows (output columns), 4 consecutive k
          const int nb = g.BN / 32 + ((g.BN & 31) ? 1 : 0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int row = r0 + 32 * j;
            if (j < nb && row < g.Ncols) {
              const float* p = g.dense + (long long)row * g.Kc + k0;
              if (k0 + 3 < g.Kc && (g.Kc & 3) == 0) vb[j] = __ldg(reinterpret_cast<const float4*>(p));
              else {
                if (k0 < g.Kc) vb[j].x = __ldg(p);
                if (k0 + 1 < g.Kc) vb[j].y = __ldg(p + 1);
                if (k0 + 2 < g.Kc) vb[j].z = __ldg(p + 2);
                if (k0 + 3 < g.Kc) vb[j].w = __ldg(p + 3);
              }
            }
          }
        } else {
          const int img = tl.img0 + i / sps, p0 = (i % sps) * kBK + c4 * 4;
          // dense A: dY of this image, rows = channels, 4 consecutive positions
          const float* dy = g.dense + (long long)img * g.Co * plane_g;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int o = r0 + 32 * j;
            if (o < g.Co) {
              const float* p = dy + (long long)o * plane_g + p0;
              if (p0 < plane_g) va[j].x = __ldg(p);
              if (p0 + 1 < plane_g) va[j].y = __ldg(p + 1);
              if (p0 + 2 < plane_g) va[j].z = __ldg(p + 2);
              if (p0 + 3 < plane_g) va[j].w = __ldg(p + 3);
            }
          }
          // gathered B: rows = gathered columns of this N tile, 4 consecutive positions of the image
          long long pb[4]; bool pok[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int p = p0 + e;
            pok[e] = p < plane_g;
            const int oy = pok[e] ? p / g.Wg : 0, ox = pok[e] ? p - oy * g.Wg : 0;
            pb[e] = (long long)img * src_img + (long long)oy * g.Ws + ox;
          }
          const int nb = g.BN / 32 + ((g.BN & 31) ? 1 : 0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int kk = tl.n0 + r0 + 32 * j;
            if (j < nb && r0 + 32 * j < g.BN && kk < g.Kc) {
              const int ko = koff[kk];
              if (pok[0]) vb[j].x = __ldg(g.src + pb[0] + ko);
              if (pok[1]) vb[j].y = __ldg(g.src + pb[1] + ko);
              if (pok[2]) vb[j].z = __ldg(g.src + pb[2] + ko);
              if (pok[3]) vb[j].w = __ldg(g.src + pb[3] + ko);
            }
          }
        }
        if (it >= g.stages) mbar_wait(&sm->empty[s], ((it / g.stages) - 1) & 1);
        uint8_t* a_dst = ring + s * stage_bytes;
        uint8_t* b_dst = a_dst + 16384;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 o = va[j];
          o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w);
          if (MODE != kDw) *reinterpret_cast<float4*>(a_dst + swz128(grow, ghalf * 4 + j)) = o;   // row `grow`, chunk ghalf*4+j
          else *reinterpret_cast<float4*>(a_dst + soff0 + j * 4096) = o;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (r0 + 32 * j < g.BN) {
            float4 o = vb[j];
            o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w);
            *reinterpret_cast<float4*>(b_dst + soff0 + j * 4096) = o;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->full[s]);
      }
    }
  } else if (warp == kEpiWarps) {
    // ------------------------------------------------------------ MMA issue
    const uint32_t idesc = idesc_tf32(kBM, g.BN, false, false);
    int it = 0, tcount = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x, ++tcount) {
      const Tile tl = decode(g, t, MODE, sps);
      const int buf = tcount & 1;
      if (tcount >= 2) mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) - 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * g.BN;
      for (int i = 0; i < tl.nk; ++i, ++it) {
        const int s = it % g.stages;
        mbar_wait(&sm->full[s], (it / g.stages) & 1);
        tc_fence_after();
        // descriptors as (low, constant high) halves stepped by immediates: the issue stream of this one warp, not
        // the tensor pipe, is what a stage of four small MMAs costs
        const uint32_t a_lo = desc_lo_k(smem_u32(ring + s * stage_bytes));
        const uint32_t b_lo = a_lo + (16384u >> 4);
        if (elect_one_sync()) {
          mma_tf32_ss_lh(acc, a_lo, kDescHiK, b_lo, kDescHiK, idesc, i > 0 ? 1u : 0u);
          mma_tf32_ss_lh(acc, a_lo + 1 * kDescStepK, kDescHiK, b_lo + 1 * kDescStepK, kDescHiK, idesc, 1u);
          mma_tf32_ss_lh(acc, a_lo + 2 * kDescStepK, kDescHiK, b_lo + 2 * kDescStepK, kDescHiK, idesc, 1u);
          mma_tf32_ss_lh(acc, a_lo + 3 * kDescStepK, kDescHiK, b_lo + 3 * kDescStepK, kDescHiK, idesc, 1u);
          mma_commit(&sm->empty[s]);
          if (i == tl.nk - 1) mma_commit(&sm->acc_full[buf]);
        }
        __syncwarp();
      }
      if (tl.nk == 0) {
        if (elect_one_sync()) mma_commit(&sm->acc_full[buf]);
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue
    int tcount = 0;
    for (unsigned t = blockIdx.x; t < g.total_tiles; t += gridDim.x, ++tcount) {
      const Tile tl = decode(g, t, MODE, sps);
      const int buf = tcount & 1;
      mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem + buf * g.BN + ((uint32_t)(warp * 32) << 16);
      const int rl = warp * 32 + lane;
      if (MODE != kDw) {
        const long long m = tl.m0 + rl;
        const bool ok = m < g.Mtot;
        const long long n = ok ? m / plane_g : 0;
        const int p = ok ? (int)(m - n * plane_g) : 0;
        float* orow = g.out + n * (long long)g.Ncols * plane_g + p;
        for (int c0 = 0; c0 < g.BN; c0 += 32) {
          float v[32];
          if (c0 + 16 < g.BN) tmem_ld32(acc + c0, v); else tmem_ld16(acc + c0, v);
          if (!ok) continue;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int o = c0 + i;
            if (o < g.Ncols && o < g.BN) orow[(long long)o * plane_g] = v[i] + (g.bias ? __ldg(g.bias + o) : 0.f);
          }
        }
      } else {
        const int ncols = min(g.BN, g.Kc - tl.n0);
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          float v[32];
          if (tl.nk > 0) {
            if (c0 + 16 < g.BN) tmem_ld32(acc + c0, v); else tmem_ld16(acc + c0, v);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          if (rl >= g.Co || tl.nk == 0) continue;
          float* wrow = g.out + (long long)rl * g.Kc + tl.n0 + c0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < ncols) atomicAdd(wrow + i, v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm->acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, g.tmem_cols);
}

// dbias[o] += sum over (n, p) of dY[n][o][p]   (gemv against the ones vector in the reference, base_conv_layer.cpp:312-316)
template <typename T>
__global__ void __launch_bounds__(256) conv_bias_grad_kernel(const T* __restrict__ dy, T* __restrict__ db, int N, int Co, int P) {
  const int o = blockIdx.x;
  T acc = T(0);
  const bool vec = sizeof(T) == 4 && (P & 3) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
  for (int n = blockIdx.y; n < N; n += gridDim.y) {
    const T* p = dy + ((size_t)n * Co + o) * P;
    if (vec) {
      for (int i = threadIdx.x; i < (P >> 2); i += 256) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(p) + i);
        acc += (T)((v.x + v.y) + (v.z + v.w));
      }
    } else {
      for (int i = threadIdx.x; i < P; i += 256) acc += p[i];
    }
  }
  __shared__ T s[256];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) s[threadIdx.x] += s[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(db + o, s[0]);
}

// ---- image-resident variant -------------------------------------------------------------------------------------
// The planes the CNN convolves are small (4 x 40 x 40 and 32 x 9 x 9 floats per sample: 25.6 and 10.4 KB), so a CTA
// loads each sample ONCE into shared memory (coalesced 16-byte loads, rounded to TF32 on the way: once per value instead
// of once per use) and gathers the kh*kw-fold redundant operand tiles from there -- no global-memory latency inside the
// stage loop, which is what bounded the global-gather kernel above (~3500 cycles per 20 KB stage).  The next sample's
// loads are in flight (registers) while the current one is being consumed.  forward keeps the whole weight matrix in
// shared memory in UMMA layout when it fits (conv0: 16 KB); dW streams dY with the loads of four stages in flight.
constexpr int kImgRegs = 7;             // float4 registers per loader thread for the next sample (<= 28 KB per sample)

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv2d_img_kernel(const ConvArgs g, const int img_floats, const int b_resident, const int tiles_per_img) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int sps_k = (g.Kc + kBK - 1) / kBK;
  const int b_bytes = g.BN * 128;
  const int stage_bytes = 16384 + ((MODE == kFwd && b_resident) ? 0 : b_bytes);
  uint8_t* bres = ring + g.stages * stage_bytes;                          // forward: resident weight blocks
  const int bres_bytes = (MODE == kFwd && b_resident) ? sps_k * b_bytes : 0;
  float* img = reinterpret_cast<float*>(bres + bres_bytes);               // the current sample, TF32-rounded
  Smem* sm = reinterpret_cast<Smem*>(reinterpret_cast<uint8_t*>(img) + ((img_floats * 4 + 127) & ~127));
  int* koff = reinterpret_cast<int*>(sm + 1);
  int* ptab = koff + g.Kc;                                 // offset of grid position p inside a source plane

  const int warp = warp_idx_sync(), lane = threadIdx.x & 31;
  const int plane_g = g.Hg * g.Wg;
  const int sps = MODE == kDw ? (plane_g + kBK - 1) / kBK : sps_k;
  for (int k = threadIdx.x; k < g.Kc; k += kThreads) {
    const int c = k / (g.kh * g.kw), r = k - c * g.kh * g.kw, ky = r / g.kw, kx = r - ky * g.kw;
    koff[k] = c * g.Hs * g.Ws + ky * g.Ws + kx;
  }
  for (int p = threadIdx.x; p < plane_g; p += kThreads) ptab[p] = (p / g.Wg) * g.Ws + p % g.Wg;
  if (warp == kEpiWarps) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) { mbar_init(&sm->full[s], kLoadWarps); mbar_init(&sm->empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&sm->acc_full[b], 1); mbar_init(&sm->acc_empty[b], kEpiWarps); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&sm->tmem_base, g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm->tmem_base;
  // the samples of this CTA: forward: n = blockIdx.x, += gridDim.x;  dW: the slice [img0, img1) of its split
  int n_first, n_step, n_end, n0_tile = 0;
  if (MODE == kFwd) { n_first = blockIdx.x; n_step = gridDim.x; n_end = g.Nimg; }
  else {
    const int nt = blockIdx.x % g.n_tiles, sp = blockIdx.x / g.n_tiles;
    const int per = (g.Nimg + g.isplit - 1) / g.isplit;
    n_first = min(g.Nimg, sp * per); n_end = min(g.Nimg, n_first + per); n_step = 1; n0_tile = nt * g.BN;
  }

  if (warp > kEpiWarps) {
    // ------------------------------------------------------------ operand staging
    const int tid = threadIdx.x - (kEpiWarps + 1) * 32;
    const int r0 = tid >> 3, c4 = tid & 7;
    const uint32_t soff0 = swz128(r0, c4);
    const int grow = tid & 127, ghalf = tid >> 7;
    const int nv = img_floats >> 2;                        // float4 per sample
    float4 pre[kImgRegs];
    auto fetch = [&](int n) {                              // the sample's loads, all in flight at once
      const float4* src = reinterpret_cast<const float4*>(g.src + (long long)n * img_floats);
#pragma unroll
      for (int j = 0; j < kImgRegs; ++j) {
        const int v = tid + j * kLoadThreads;
        if (v < nv) pre[j] = __ldg(src + v);
      }
    };
    auto park = [&]() {                                    // registers -> shared memory, rounded once
#pragma unroll
      for (int j = 0; j < kImgRegs; ++j) {
        const int v = tid + j * kLoadThreads;
        if (v < nv) {
          float4 o = pre[j];
          o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w);
          reinterpret_cast<float4*>(img)[v] = o;
        }
      }
    };
    if (MODE == kFwd && b_resident) {                      // weight matrix -> UMMA K-major blocks, once
      for (int i = 0; i < sps_k; ++i) {
        const int k0 = i * kBK + c4 * 4;
        for (int j = 0; j * 32 + r0 < g.BN; ++j) {
          const int row = r0 + 32 * j;
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < g.Ncols) {
            const float* p = g.dense + (long long)row * g.Kc + k0;
            if (k0 < g.Kc) o.x = to_tf32(__ldg(p));
            if (k0 + 1 < g.Kc) o.y = to_tf32(__ldg(p + 1));
            if (k0 + 2 < g.Kc) o.z = to_tf32(__ldg(p + 2));
            if (k0 + 3 < g.Kc) o.w = to_tf32(__ldg(p + 3));
          }
          *reinterpret_cast<float4*>(bres + i * b_bytes + soff0 + j * 4096) = o;
        }
      }
    }
    const bool fast4 = MODE == kFwd && b_resident && sps == 4;
    int ko4[4][16];
    uint32_t sw4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sw4[j] = swz128(grow, ghalf * 4 + j);
    if (fast4) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int k = i * kBK + ghalf * 16 + e;
          ko4[i][e] = k < g.Kc ? koff[k] : 0;
        }
    }
    int it = 0;
    if (!fast4 && n_first < n_end) fetch(n_first);
    for (int n = n_first; n < n_end; n += n_step) {
      // every loader has finished gathering from the previous sample before it is overwritten
      asm volatile("bar.sync 1, %0;" ::"n"(kLoadThreads));
      if (fast4) fetch(n);                                 // (the register-resident offsets leave no room for a prefetch:
      park();                                              //  one exposed load latency per sample, ~3 % of its time)
      asm volatile("bar.sync 1, %0;" ::"n"(kLoadThreads));
      if (!fast4 && n + n_step < n_end) fetch(n + n_step); // in flight while this sample is consumed
      if (MODE == kFwd && fast4) {
        // 4 k blocks (conv0: K = 100), weights resident: the 64 gather offsets of this thread live in registers and the
        // loop body is 16 shared-memory loads and 4 stores per stage -- no table look-ups, no predicates (columns past K
        // meet zero weight rows, rows past the plane are never stored: both read a valid address)
        for (int mt = 0; mt < tiles_per_img; ++mt) {
          const int p = mt * kBM + grow;
          const float* rowp = img + (p < plane_g ? ptab[p] : 0);
#pragma unroll
          for (int i = 0; i < 4; ++i, ++it) {
            const int s = it % g.stages;
            float4 va[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              va[cc] = make_float4(rowp[ko4[i][cc * 4]], rowp[ko4[i][cc * 4 + 1]], rowp[ko4[i][cc * 4 + 2]], rowp[ko4[i][cc * 4 + 3]]);
            if (it >= g.stages) mbar_wait(&sm->empty[s], ((it / g.stages) - 1) & 1);
            uint8_t* a_dst = ring + s * stage_bytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(a_dst + sw4[j]) = va[j];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->full[s]);
          }
        }
      } else if (MODE == kFwd) {
        for (int mt = 0; mt < tiles_per_img; ++mt) {
          const int p = mt * kBM + grow;
          const bool gok = p < plane_g;
          const float* rowp = img + (gok ? ptab[p] : 0);
          for (int i = 0; i < sps; ++i, ++it) {
            const int s = it % g.stages;
            float4 va[4], vb[8];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int kc = i * kBK + (ghalf * 4 + cc) * 4;
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int ko = (kc + e < g.Kc) ? koff[kc + e] : -1;       // uniform over the warp: a broadcast read
                v[e] = (gok && ko >= 0) ? rowp[ko] : 0.f;
              }
              va[cc] = make_float4(v[0], v[1], v[2], v[3]);
            }
            if (!b_resident) {
              const int k0 = i * kBK + c4 * 4;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                vb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int row = r0 + 32 * j;
                if (row < g.BN && row < g.Ncols) {
                  const float* pw = g.dense + (long long)row * g.Kc + k0;
                  if (k0 + 3 < g.Kc && (g.Kc & 3) == 0) vb[j] = __ldg(reinterpret_cast<const float4*>(pw));
                  else {
                    if (k0 < g.Kc) vb[j].x = __ldg(pw);
                    if (k0 + 1 < g.Kc) vb[j].y = __ldg(pw + 1);
                    if (k0 + 2 < g.Kc) vb[j].z = __ldg(pw + 2);
                    if (k0 + 3 < g.Kc) vb[j].w = __ldg(pw + 3);
                  }
                }
              }
            }
            if (it >= g.stages) mbar_wait(&sm->empty[s], ((it / g.stages) - 1) & 1);
            uint8_t* a_dst = ring + s * stage_bytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(a_dst + swz128(grow, ghalf * 4 + j)) = va[j];
            if (!b_resident) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (r0 + 32 * j < g.BN) {
                  float4 o = vb[j];
                  o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w);
                  *reinterpret_cast<float4*>(a_dst + 16384 + soff0 + j * 4096) = o;
                }
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->full[s]);
          }
        }
      } else {
        // dW: stages = 32-position blocks of this sample; dY loads of four stages in flight
        const float* dy = g.dense + (long long)n * g.Co * plane_g;
        int kod[8];                                          // gather offsets of this thread's rows: constant for the kernel
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int kk = n0_tile + r0 + 32 * j; kod[j] = kk < g.Kc ? koff[kk] : 0; }
        // rows of the dY tile = channels: only ceil(Co / 32) of the 4 row groups exist (rows past Co stay unwritten --
        // they only feed accumulator rows the epilogue never reads)
        const int ja = (g.Co + 31) / 32;
        auto load_group = [&](int ib, float4 (&dst)[4][2]) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p0 = (ib + q) * kBK + c4 * 4;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              dst[q][j] = make_float4(0.f, 0.f, 0.f, 0.f);
              const int o = r0 + 32 * j;
              if (j < ja && ib + q < sps && o < g.Co) {
                const float* p = dy + (long long)o * plane_g + p0;
                if (p0 + 3 < plane_g && (plane_g & 3) == 0) dst[q][j] = __ldg(reinterpret_cast<const float4*>(p));
                else {
                  if (p0 < plane_g) dst[q][j].x = __ldg(p);
                  if (p0 + 1 < plane_g) dst[q][j].y = __ldg(p + 1);
                  if (p0 + 2 < plane_g) dst[q][j].z = __ldg(p + 2);
                  if (p0 + 3 < plane_g) dst[q][j].w = __ldg(p + 3);
                }
              }
            }
          }
        };
        float4 cur[4][2], nxt[4][2];
        load_group(0, cur);
        for (int ib = 0; ib < sps; ib += 4) {
          if (ib + 4 < sps) load_group(ib + 4, nxt);         // in flight while the current four stages are built
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (ib + q >= sps) break;
            const int s = it % g.stages;
            const int p0 = (ib + q) * kBK + c4 * 4;
            // positions past the plane meet zero dY columns and gathered rows past K are never stored: both read offset 0
            int po[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) po[e] = (p0 + e < plane_g) ? ptab[p0 + e] : 0;
            float4 vb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (r0 + 32 * j < g.BN) {
                const float* base = img + kod[j];
                vb[j] = make_float4(base[po[0]], base[po[1]], base[po[2]], base[po[3]]);
              }
            }
            if (it >= g.stages) mbar_wait(&sm->empty[s], ((it / g.stages) - 1) & 1);
            uint8_t* a_dst = ring + s * stage_bytes;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (j < ja) {
                float4 o = cur[q][j];
                o.x = to_tf32(o.x); o.y = to_tf32(o.y); o.z = to_tf32(o.z); o.w = to_tf32(o.w);
                *reinterpret_cast<float4*>(a_dst + soff0 + j * 4096) = o;
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (r0 + 32 * j < g.BN) *reinterpret_cast<float4*>(a_dst + 16384 + soff0 + j * 4096) = vb[j];
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm->full[s]);
            ++it;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) { cur[q][0] = nxt[q][0]; cur[q][1] = nxt[q][1]; }
        }
      }
    }
  } else if (warp == kEpiWarps) {
    // ------------------------------------------------------------ MMA issue
    const uint32_t idesc = idesc_tf32(kBM, g.BN, false, false);
    const uint32_t ring_lo = desc_lo_k(smem_u32(ring)), stage_lo = (uint32_t)stage_bytes >> 4;
    const uint32_t bres_lo = desc_lo_k(smem_u32(bres)), b_lo_step = (uint32_t)b_bytes >> 4;
    int it = 0, tcount = 0;
    // (the resident weight blocks are written by the loader warps before their first arrive on full[0], behind the
    // same proxy fence as the stage itself: waiting for a stage also covers them)
    if (MODE == kFwd) {
      for (int n = n_first; n < n_end; n += n_step) {
        for (int mt = 0; mt < tiles_per_img; ++mt, ++tcount) {
          const int buf = tcount & 1;
          if (tcount >= 2) mbar_wait(&sm->acc_empty[buf], ((tcount >> 1) - 1) & 1);
          tc_fence_after();
          const uint32_t acc = tmem + buf * g.BN;
          for (int i = 0; i < sps; ++i, ++it) {
            const int s = it % g.stages;
            mbar_wait(&sm->full[s], (it / g.stages) & 1);
            tc_fence_after();
            const uint32_t a_lo = ring_lo + s * stage_lo;
            const uint32_t b_lo = b_resident ? bres_lo + i * b_lo_step : a_lo + (16384u >> 4);
            if (elect_one_sync()) {
              mma_tf32_ss_lh(acc, a_lo, kDescHiK, b_lo, kDescHiK, idesc, i > 0 ? 1u : 0u);
              mma_tf32_ss_lh(acc, a_lo + 1 * kDescStepK, kDescHiK, b_lo + 1 * kDescStepK, kDescHiK, idesc, 1u);
              mma_tf32_ss_lh(acc, a_lo + 2 * kDescStepK, kDescHiK, b_lo + 2 * kDescStepK, kDescHiK, idesc, 1u);
              mma_tf32_ss_lh(acc, a_lo + 3 * kDescStepK, kDescHiK, b_lo + 3 * kDescStepK, kDescHiK, idesc, 1u);
              mma_commit(&sm->empty[s]);
              if (i == sps - 1) mma_commit(&sm->acc_full[buf]);
            }
            __syncwarp();
          }
        }
      }
    } else {
      const int nk = (n_end - n_first) * sps;
      const uint32_t acc = tmem;
      for (int i = 0; i < nk; ++i, ++it) {
        const int s = it % g.stages;
        mbar_wait(&sm->full[s], (it / g.stages) & 1);
        tc_fence_after();
        const uint32_t a_lo = ring_lo + s * stage_lo;
        const uint32_t b_lo = a_lo + (16384u >> 4);
        if (elect_one_sync()) {
          mma_tf32_ss_lh(acc, a_lo, kDescHiK, b_lo, kDescHiK, idesc, i > 0 ? 1u : 0u);
          mma_tf32_ss_lh(acc, a_lo + 1 * kDescStepK, kDescHiK, b_lo + 1 * kDescStepK, kDescHiK, idesc, 1u);
          mma_tf32_ss_lh(acc, a_lo + 2 * kDescStepK, kDescHiK, b_lo + 2 * kDescStepK, kDescHiK, idesc, 1u);
          mma_tf32_ss_lh(acc, a_lo + 3 * kDescStepK, kDescHiK, b_lo + 3 * kDescStepK, kDescHiK, idesc, 1u);
          mma_commit(&sm->empty[s]);
          if (i == nk - 1) mma_commit(&sm->acc_full[0]);
        }
        __syncwarp();
      }
      if (nk == 0) {
        if (elect_one_sync()) mma_commit(&sm->acc_full[0]);
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue
    const int rl = warp * 32 + lane;
    if (MODE == kFwd) {
      int tcount = 0;
      for (int n = n_first; n < n_end; n += n_step) {
        for (int mt = 0; mt < tiles_per_img; ++mt, ++tcount) {
          const int buf = tcount & 1;
          mbar_wait(&sm->acc_full[buf], (tcount >> 1) & 1);
          tc_fence_after();
          const uint32_t acc = tmem + buf * g.BN + ((uint32_t)(warp * 32) << 16);
          const int p = mt * kBM + rl;
          const bool ok = p < plane_g;
          float* orow = g.out + (long long)n * g.Ncols * plane_g + p;
          for (int c0 = 0; c0 < g.BN; c0 += 32) {
            float v[32];
            if (c0 + 16 < g.BN) tmem_ld32(acc + c0, v); else tmem_ld16(acc + c0, v);
            if (!ok) continue;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int o = c0 + i;
              if (o < g.Ncols && o < g.BN) orow[(long long)o * plane_g] = v[i] + (g.bias ? __ldg(g.bias + o) : 0.f);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm->acc_empty[buf]);
        }
      }
    } else {
      const int nk = (n_end - n_first) * sps;
      mbar_wait(&sm->acc_full[0], 0);
      tc_fence_after();
      const uint32_t acc = tmem + ((uint32_t)(warp * 32) << 16);
      const int ncols = min(g.BN, g.Kc - n0_tile);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        float v[32];
        if (nk > 0) {
          if (c0 + 16 < g.BN) tmem_ld32(acc + c0, v); else tmem_ld16(acc + c0, v);
        }
        if (rl >= g.Co || nk == 0) continue;
        float* wrow = g.out + (long long)rl * g.Kc + n0_tile + c0;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < ncols) atomicAdd(wrow + i, v[i]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, g.tmem_cols);
}

// col2im_cpu (im2col.cpp:60-90) in gather form: dx[n][c][y][x] = sum over (ky, kx) of Zt[n][(c, ky, kx)][(y - ky, x - kx)]
// where that position exists; every element of Zt is read exactly once, consecutive x read consecutive addresses.
__global__ void __launch_bounds__(256)
conv_col2im_kernel(const float* __restrict__ Zt, float* __restrict__ dx, long long total, int C, int H, int Wd, int kh, int kw,
                   int OH, int OW) {
  const int P = OH * OW, khw = kh * kw;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int x = (int)(e % Wd); long long r = e / Wd;
    const int y = (int)(r % H); r /= H;
    const int c = (int)(r % C); const long long n = r / C;
    const float* z = Zt + ((size_t)n * C + c) * khw * P;
    float acc = 0.f;
    const int ky0 = max(0, y - OH + 1), ky1 = min(kh - 1, y);
    const int kx0 = max(0, x - OW + 1), kx1 = min(kw - 1, x);
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) acc += __ldg(z + (size_t)(ky * kw + kx) * P + (y - ky) * OW + (x - kx));
    dx[e] = acc;
  }
}

// ---- SIMT forms (double blobs, MMS_MATH_FP32, shapes outside the tensor-core kernel): direct sums
template <typename T>
__global__ void __launch_bounds__(256)
conv_fwd_simt(const T* __restrict__ x, const T* __restrict__ W, const T* __restrict__ b, T* __restrict__ y, long long total,
              int C, int H, int Wd, int Co, int kh, int kw, int OH, int OW) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int ox = (int)(e % OW); long long r = e / OW;
    const int oy = (int)(r % OH); r /= OH;
    const int o = (int)(r % Co); const long long n = r / Co;
    T acc = b ? b[o] : T(0);
    for (int c = 0; c < C; ++c)
      for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx)
          acc += x[((n * C + c) * H + oy + ky) * Wd + ox + kx] * W[((o * C + c) * kh + ky) * kw + kx];
    y[e] = acc;
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
conv_dx_simt(const T* __restrict__ dy, const T* __restrict__ W, T* __restrict__ dx, long long total, int C, int H, int Wd,
             int Co, int kh, int kw, int OH, int OW) {
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int xx = (int)(e % Wd); long long r = e / Wd;
    const int yy = (int)(r % H); r /= H;
    const int c = (int)(r % C); const long long n = r / C;
    T acc = T(0);
    for (int o = 0; o < Co; ++o)
      for (int ky = 0; ky < kh; ++ky) {
        const int oy = yy - ky;
        if (oy < 0 || oy >= OH) continue;
        for (int kx = 0; kx < kw; ++kx) {
          const int ox = xx - kx;
          if (ox < 0 || ox >= OW) continue;
          acc += dy[((n * Co + o) * OH + oy) * OW + ox] * W[((o * C + c) * kh + ky) * kw + kx];
        }
      }
    dx[e] = acc;
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
conv_dw_simt(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dW, int N, int C, int H, int Wd, int Co, int kh,
             int kw, int OH, int OW) {
  // one CTA per weight: the (n, oy, ox) sum is spread over the threads
  const int e = blockIdx.x;
  const int kx = e % kw; int r = e / kw;
  const int ky = r % kh; r /= kh;
  const int c = r % C; const int o = r / C;
  T acc = T(0);
  const long long total = (long long)N * OH * OW;
  for (long long i = threadIdx.x; i < total; i += 256) {
    const int ox = (int)(i % OW); long long q = i / OW;
    const int oy = (int)(q % OH); const long long n = q / OH;
    acc += dy[((n * Co + o) * OH + oy) * OW + ox] * x[((n * C + c) * H + oy + ky) * Wd + ox + kx];
  }
  __shared__ T s[256];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (threadIdx.x < st) s[threadIdx.x] += s[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) dW[e] += s[0];
}

int launch_tc(mms_context* ctx, ConvArgs& g, int mode) {
  int stages = kMaxStages;
  const size_t tail = sizeof(Smem) + 2 * sizeof(int) * (size_t)g.Kc + 1024;
  while (stages > 2 && (size_t)stages * (16384 + g.BN * 128) + tail > 200 * 1024) --stages;
  g.stages = stages;
  g.tmem_cols = umma::tmem_cols_pow2(2 * g.BN);
  const size_t smem = (size_t)stages * (16384 + g.BN * 128) + tail;
  static bool configured = false;
  if (!configured) {
    MMS_MAX_SMEM(conv2d_kernel<kFwd>, 201 * 1024);
    MMS_MAX_SMEM(conv2d_kernel<kDw>, 201 * 1024);
    configured = true;
  }
  const unsigned grid = (unsigned)mms_min<long long>(g.total_tiles, ctx->sm_count);
  if (mode == kFwd) { MmsKernelScope ks_(ctx, "conv2d_fwd_kernel"); conv2d_kernel<kFwd><<<grid, kThreads, smem, ctx->stream>>>(g); }
  else { MmsKernelScope ks_(ctx, "conv2d_dw_kernel"); conv2d_kernel<kDw><<<grid, kThreads, smem, ctx->stream>>>(g); }
  MMS_LAUNCH_CHECK();
  return 0;
}

// The image-resident kernel when a sample fits its registers / shared memory; else the global-gather kernel.
int launch_conv(mms_context* ctx, ConvArgs& g, int mode) {
  const long long img_floats = (long long)g.Cs * g.Hs * g.Ws;
  const int plane = g.Hg * g.Wg;
  // per-sample tiles (forward) / stages (dW) must be reasonably full: a 5 x 5 output plane would use 25 of 128 rows
  const bool img_ok = img_floats % 4 == 0 && img_floats <= (long long)kImgRegs * kLoadThreads * 4 &&
                      (reinterpret_cast<uintptr_t>(g.src) & 15) == 0 && plane >= 4 * kBM &&
                      (mode != kDw || g.Co <= 64);          // the image kernel's dW stages two 32-row groups of dY
  if (!img_ok) return launch_tc(ctx, g, mode);
  const int sps_k = mms_ceil_div(g.Kc, kBK);
  const int b_bytes = g.BN * 128;
  const int b_resident = mode == kFwd && (size_t)sps_k * b_bytes <= 48 * 1024;
  const size_t fixed = (size_t)(b_resident ? sps_k * b_bytes : 0) + (((size_t)img_floats * 4 + 127) & ~(size_t)127) + sizeof(Smem) +
                       sizeof(int) * ((size_t)g.Kc + plane) + 1024;
  const int stage_bytes = 16384 + (b_resident ? 0 : b_bytes);
  int stages = kMaxStages;
  while (stages > 2 && (size_t)stages * stage_bytes + fixed > 200 * 1024) --stages;
  if ((size_t)stages * stage_bytes + fixed > 200 * 1024) return launch_tc(ctx, g, mode);
  g.stages = stages;
  g.tmem_cols = umma::tmem_cols_pow2(2 * g.BN);
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  static bool configured = false;
  if (!configured) {
    MMS_MAX_SMEM(conv2d_img_kernel<kFwd>, 201 * 1024);
    MMS_MAX_SMEM(conv2d_img_kernel<kDw>, 201 * 1024);
    configured = true;
  }
  const int tiles_per_img = mms_ceil_div(plane, kBM);
  if (mode == kFwd) {
    const unsigned grid = (unsigned)mms_min(g.Nimg, ctx->sm_count);
    MmsKernelScope ks_(ctx, "conv2d_fwd_kernel");
    conv2d_img_kernel<kFwd><<<grid, kThreads, smem, ctx->stream>>>(g, (int)img_floats, b_resident, tiles_per_img);
  } else {
    MmsKernelScope ks_(ctx, "conv2d_dw_kernel");
    conv2d_img_kernel<kDw><<<g.total_tiles, kThreads, smem, ctx->stream>>>(g, (int)img_floats, 0, tiles_per_img);
  }
  MMS_LAUNCH_CHECK();
  return 0;
}

bool tc_shape_ok(int C, int Co, int kh, int kw) {
  return C * kh * kw <= kMaxKc && Co * kh * kw <= kMaxKc && Co <= 128 && C <= 256;
}

inline int ew_grid(mms_context* ctx, long long n) {
  return (int)mms_max<long long>(1, mms_min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16));
}

}  // namespace

template <typename T>
int mms_conv2d_forward_impl(mms_context* ctx, const T* x, const T* W, const T* bias, T* top, int N, int C, int H, int Wd,
                            int Co, int kh, int kw) {
  MMS_REQUIRE(x && W && top, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && C > 0 && Co > 0 && kh > 0 && kw > 0 && H >= kh && Wd >= kw, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  const int OH = H - kh + 1, OW = Wd - kw + 1;
  if (sizeof(T) == 4 && ctx->math == MMS_MATH_TF32 && tc_shape_ok(C, Co, kh, kw)) {
    ConvArgs g = {};
    g.src = reinterpret_cast<const float*>(x); g.Cs = C; g.Hs = H; g.Ws = Wd; g.Hg = OH; g.Wg = OW; g.sgn = 1;
    g.kh = kh; g.kw = kw; g.Kc = C * kh * kw; g.Nimg = N;
    g.dense = reinterpret_cast<const float*>(W); g.Ncols = Co;
    g.out = reinterpret_cast<float*>(top); g.bias = reinterpret_cast<const float*>(bias);
    g.BN = mms_ceil_div(Co, 16) * 16; g.n_tiles = 1; g.isplit = 1;
    g.Mtot = (long long)N * OH * OW;
    MMS_REQUIRE(mms_ceil_div(g.Mtot, kBM) <= 0x7fffffff, MMS_E_UNSUPPORTED, "too many rows");
    g.total_tiles = (unsigned)mms_ceil_div(g.Mtot, kBM);
    return launch_conv(ctx, g, kFwd);
  }
  const long long total = (long long)N * Co * OH * OW;
  { MmsKernelScope ks_(ctx, "conv2d_fwd_simt");
    conv_fwd_simt<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(x, W, bias, top, total, C, H, Wd, Co, kh, kw, OH, OW); }
  MMS_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int mms_conv2d_backward_impl(mms_context* ctx, const T* x, const T* W, const T* dtop, T* dW, T* dbias, T* dx, int N, int C,
                             int H, int Wd, int Co, int kh, int kw) {
  MMS_REQUIRE(x && W && dtop, MMS_E_INVALID, "null pointer");
  MMS_REQUIRE(N >= 0 && C > 0 && Co > 0 && kh > 0 && kw > 0 && H >= kh && Wd >= kw, MMS_E_INVALID, "bad size");
  if (N == 0) return 0;
  const int OH = H - kh + 1, OW = Wd - kw + 1, P = OH * OW;
  const bool tc = sizeof(T) == 4 && ctx->math == MMS_MATH_TF32 && tc_shape_ok(C, Co, kh, kw);
  if (dbias) {
    dim3 grid(Co, mms_max(1, mms_min(N, mms_ceil_div(4 * ctx->sm_count, Co))));
    MmsKernelScope ks_(ctx, "conv2d_bias_grad_kernel");
    conv_bias_grad_kernel<T><<<grid, 256, 0, ctx->stream>>>(dtop, dbias, N, Co, P);
    MMS_LAUNCH_CHECK();
  }
  if (dW) {
    if (tc) {
      ConvArgs g = {};
      g.src = reinterpret_cast<const float*>(x); g.Cs = C; g.Hs = H; g.Ws = Wd; g.Hg = OH; g.Wg = OW; g.sgn = 1;
      g.kh = kh; g.kw = kw; g.Kc = C * kh * kw; g.Nimg = N;
      g.dense = reinterpret_cast<const float*>(dtop); g.Co = Co;
      g.out = reinterpret_cast<float*>(dW);
      g.n_tiles = mms_ceil_div(g.Kc, 256);
      g.BN = mms_ceil_div(mms_ceil_div(g.Kc, g.n_tiles), 16) * 16;
      g.n_tiles = mms_ceil_div(g.Kc, g.BN);
      g.isplit = mms_max(1, mms_min(N, ctx->sm_count / g.n_tiles));
      g.total_tiles = (unsigned)(g.n_tiles * g.isplit);
      MMS_TRY(launch_conv(ctx, g, kDw));
    } else {
      MmsKernelScope ks_(ctx, "conv2d_dw_simt");
      conv_dw_simt<T><<<Co * C * kh * kw, 256, 0, ctx->stream>>>(x, dtop, dW, N, C, H, Wd, Co, kh, kw, OH, OW);
      MMS_LAUNCH_CHECK();
    }
  }
  if (dx) {
    if (tc) {
      // backward_cpu_gemm + col2im of the reference (base_conv_layer.cpp:289-298), batched: Zt_n [Kc x P] = Wm^T dY_n as
      // ONE batched TF32 GEMM over the images (both operands read in place, MN-major), then a gather-form col2im in
      // which consecutive threads read consecutive positions of one row of Zt.  (The contraction as an implicit GEMM
      // over (o, ky, kx) with C_in output columns re-reads every dY value kh*kw times for a handful of columns: 20 ms
      // at 4096 x 32 x 36 x 36 -> 4 x 40 x 40 against ~1 ms this way.)
      const int Kc = C * kh * kw;
      const long long per_img = (long long)Kc * P;
      long long chunk = mms_max<long long>(1, (long long)(ctx->scratch_cap / sizeof(float)) / per_img);
      chunk = mms_min<long long>(chunk, N);
      void* sp = nullptr;
      MMS_TRY(mms_scratch(ctx, sizeof(float) * (size_t)(chunk * per_img), &sp));
      float* Zt = static_cast<float*>(sp);
      for (long long n0 = 0; n0 < N; n0 += chunk) {
        const int nc = (int)mms_min<long long>(chunk, N - n0);
        TcGemmArgs t = tc_gemm_args(reinterpret_cast<const float*>(W), Kc, 1,
                                    reinterpret_cast<const float*>(dtop) + (size_t)n0 * Co * P, P, 1, Zt, P, Kc, P, Co);
        t.nb1 = nc; t.sA1 = 0; t.sB1 = (long long)Co * P; t.sC1 = per_img;
        MMS_TRY(mms_tc_gemm(ctx, t));
        const long long total = (long long)nc * C * H * Wd;
        { MmsKernelScope ks_(ctx, "conv2d_col2im_kernel");
          conv_col2im_kernel<<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(
              Zt, reinterpret_cast<float*>(dx) + (size_t)n0 * C * H * Wd, total, C, H, Wd, kh, kw, OH, OW); }
        MMS_LAUNCH_CHECK();
      }
    } else {
      const long long total = (long long)N * C * H * Wd;
      MmsKernelScope ks_(ctx, "conv2d_dx_simt");
      conv_dx_simt<T><<<ew_grid(ctx, total), 256, 0, ctx->stream>>>(dtop, W, dx, total, C, H, Wd, Co, kh, kw, OH, OW);
      MMS_LAUNCH_CHECK();
    }
  }
  return 0;
}

template int mms_conv2d_forward_impl<float>(mms_context*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int);
template int mms_conv2d_forward_impl<double>(mms_context*, const double*, const double*, const double*, double*, int, int, int, int, int, int, int);
template int mms_conv2d_backward_impl<float>(mms_context*, const float*, const float*, const float*, float*, float*, float*, int, int, int, int, int, int, int);
template int mms_conv2d_backward_impl<double>(mms_context*, const double*, const double*, const double*, double*, double*, double*, int, int, int, int, int, int, int);
