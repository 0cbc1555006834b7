// Fused forward cell for a dense support that fits one tile (the shipped SF shape: N = 100 regions, C = 5 categories,
// h = 16, Ks = Kc = 2):  one launch computes what /root/reference/framework/STC_GNN.py:65-79 does in two BDG_Dif calls
// (:31-47) -- spatial mode product (:37), categorical mode product (:38), weight contraction (:42), bias/activation
// (:44-46), sigmoid / tanh / GRU blend (:71-78) -- with the spatial terms and the feature tensor never leaving the SM.
//
// One persistent CTA per SM walks samples; a sample's [N, C*(h+Din)] state is processed as C category blocks of 32
// feature columns ([h | x | zero pad]) in two passes (gates, then candidate on [r*H | Xt]):
//
//   workers (8 warps)   load block c' of [Hlike | Xt] from HBM/L2, split hi/lo (3xTF32) and write it twice: as the
//                       MN-major B image of the spatial GEMM and as the K-major A atom (k = 0) of the gate GEMM;
//                       later read Y1 = Gs^T X back from TMEM, store it (backward needs it), split it into the k = 1 atom.
//   MMA warp            spatial GEMM  Y1[m][32] = Gs^T[m][n] X[n][32]: the support's hi part is the A operand RESIDENT IN
//                       TENSOR MEMORY (lane = output node), its lo part an MN-major image in shared memory;
//                       gate GEMM     D_c'[m][(c,o)] = [X_c' | Y1_c'] x W   (one TMEM accumulator block per category).
//   epilogue (workers)  out[m,d,:] = D_d[c=0] + sum_c' T_1(Gc)[c',d] D_c'[c=1]  (the categorical mix commuted to the GEMM
//                       output, every operand already in the thread's own TMEM lane), bias, activation, sigmoid -> u, r, r*H
//                       (pass 0) or tanh and the GRU blend -> c, H' (pass 1).
//
// Everything the multi-kernel backward reads is written on the way: u, r, c, r*H, the spatial terms of H, Xt, r*H and the
// pre-mix partial outputs P_1.  HBM traffic per row: reads Din + h, writes h + (2 Din + 10 h) saved floats.
#include "stc_conv_common.cuh"
#include "stc_tc.cuh"

#include <stdlib.h>

namespace stc {

using namespace tc;

constexpr int CF_WORKER_WARPS = 8, CF_THREADS = 32 * (CF_WORKER_WARPS + 1);
constexpr int CF_BLK = 32;            // feature columns per category block = one K atom
constexpr int CF_MAXC = 5;            // categories: C * 64 accumulator columns must fit TMEM next to the support
constexpr int CF_TM_Y = 320;          // TMEM columns: [0,320) D blocks, [320,384) two Y1 slots, [384, 384+Kp) support hi
constexpr int CF_TM_G = 384;

struct CellFusedArgs {
  int B, N, C, Din, h, act, has_bias, Kp, trace;
  const float* Gs;      // [N][N]
  const float* Q1;      // T_1(Gc) = Gc, [C][C]
  const float* xt;
  long long xt_bs;
  const float* h_prev;
  const float* Wg;      // [(2*2*L)][2h]
  const float* bg;
  const float* Wc;      // [(2*2*L)][h]
  const float* bc;
  float* h_out;
  float *u, *r, *c, *rH, *Yr1, *Yh1, *Yx1, *Pg, *Pc;
  uint32_t off_glo, off_wg, off_wc, off_img, off_k0, off_k1, off_misc, off_bar, smem_bytes, img_bytes;
};

// barrier slots
enum { CB_IMG_FULL = 0, CB_IMG_FREE = 2, CB_Y_FULL = 4, CB_Y_FREE = 6, CB_K0_FULL = 8, CB_K0_FREE, CB_K1_FULL, CB_K1_FREE,
       CB_D_FULL, CB_D_FREE, CB_COUNT };

__device__ __forceinline__ void wait_nth(uint64_t* bar, int n) { mbar_wait(bar, (uint32_t)n & 1u); }          // n-th completion
__device__ __forceinline__ void wait_free(uint64_t* bar, int n) { mbar_wait(bar, ((uint32_t)n & 1u) ^ 1u); }   // before n-th reuse
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * CF_WORKER_WARPS) : "memory"); }

__global__ void __launch_bounds__(CF_THREADS, 1)
tc_cell_fwd_fused_kernel(const CellFusedArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, C = a.C, Din = a.Din, h = a.h, L = a.Din + a.h, Kp = a.Kp;
  uint8_t* Glo = smem + a.off_glo;                       // support lo part, MN-major A image [4 col blocks][Kp][128 B]
  uint8_t* Wg_hi = smem + a.off_wg;                      // gates B atoms [2 k][64 rows (c,o)][128 B], then lo
  uint8_t* Wc_hi = smem + a.off_wc;                      // candidate B atoms [2 k][32 rows][128 B], then lo
  uint8_t* img = smem + a.off_img;                       // [2 slots][hi | lo][Kp][128 B]
  uint8_t* k0_hi = smem + a.off_k0;                      // A atom k = 0 (hi, lo)
  uint8_t* k1_hi = smem + a.off_k1;                      // A atom k = 1 (hi, lo)
  float* Qs = reinterpret_cast<float*>(smem + a.off_misc);   // [C][C]
  float* bias_s = Qs + 32;                               // [3h]: gates then candidate
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + CB_COUNT);
  const uint32_t atomA = 128 * ATOM_ROW_BYTES;           // 16 KB
  const uint32_t colblk = (uint32_t)Kp * ATOM_ROW_BYTES;
  const uint32_t WG_ATOM = 64 * ATOM_ROW_BYTES, WC_ATOM = 32 * ATOM_ROW_BYTES;

  // ---------------- one-time setup ----------------
  if (tid == 0) {
    for (int i = 0; i < CB_COUNT; ++i) {
      const bool by_workers = i == CB_IMG_FULL || i == CB_IMG_FULL + 1 || i == CB_Y_FREE || i == CB_Y_FREE + 1 ||
                              i == CB_K0_FULL || i == CB_K1_FULL || i == CB_D_FREE;
      mbar_init(&bars[i], by_workers ? CF_WORKER_WARPS : 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  // zero the image ring and both A atoms: rows >= N and padding columns are never written again
  for (uint32_t i = tid * 16u; i < 2 * a.img_bytes; i += CF_THREADS * 16u) *reinterpret_cast<float4*>(img + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t i = tid * 16u; i < 4 * atomA; i += CF_THREADS * 16u) *reinterpret_cast<float4*>(k0_hi + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  // support lo image: element (k = input node, m = output node) = Gs[k][m]
  for (int it = tid; it < Kp * 32; it += CF_THREADS) {
    const int k = it >> 5, ch = it & 31;
    float lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = ch * 4 + i;
      const float v = (k < N && m < N) ? a.Gs[(size_t)k * N + m] : 0.f;
      float hi;
      split_tf32(v, hi, lo[i]);
    }
    *reinterpret_cast<float4*>(Glo + (uint32_t)(ch >> 3) * colblk + mn32_chunk_offset(k, ch & 7)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
  // weight atoms: Bt[(c,o)][kb] = W[((k*2 + c)*L + l(kb))*Hout + o], K layout of a block = [h-part | x-part | 0]
  for (int conv = 0; conv < 2; ++conv) {
    const int Hout = conv == 0 ? 2 * h : h, rows = 2 * Hout;
    const float* W = conv == 0 ? a.Wg : a.Wc;
    uint8_t* bh = conv == 0 ? Wg_hi : Wc_hi;
    const uint32_t atom = conv == 0 ? WG_ATOM : WC_ATOM;
    uint8_t* bl = bh + 2 * atom;
    for (int it = tid; it < 2 * rows * 8; it += CF_THREADS) {
      const int k = it / (rows * 8), rem = it - k * rows * 8, n = rem >> 3, qq = rem & 7;
      const int c = n / Hout, o = n - c * Hout;
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kb = qq * 4 + i;
        const int l = kb < h ? Din + kb : (kb - h < Din ? kb - h : -1);
        v[i] = l >= 0 ? W[((size_t)(k * 2 + c) * L + l) * Hout + o] : 0.f;
      }
      store_split4(bh + k * atom, bl + k * atom, atom_chunk_offset(n, qq), make_float4(v[0], v[1], v[2], v[3]));
    }
  }
  for (int i = tid; i < C * C; i += CF_THREADS) Qs[i] = a.Q1[i];
  for (int i = tid; i < 3 * h; i += CF_THREADS) bias_s[i] = a.has_bias ? (i < 2 * h ? a.bg[i] : a.bc[i - 2 * h]) : 0.f;
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tG = tmem_base + CF_TM_G;
  // support hi part -> TMEM (lane m = output node, column k = input node): two warps per lane quarter share the columns
  if (warp < 8) {
    const int q4 = warp & 3, part = warp >> 2;
    const int m = q4 * 32 + lane;
    for (int k0 = part * 8; k0 < Kp; k0 += 16) {
      float hi[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        const float v = (k < N && m < N) ? a.Gs[(size_t)k * N + m] : 0.f;
        hi[i] = to_tf32_rn(v);
      }
      tmem_st8(tG + ((uint32_t)(q4 * 32) << 16) + (uint32_t)k0, hi);
    }
    tmem_st_wait();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();

  const int my_samples = a.B > (int)blockIdx.x ? (a.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const long long RC = (long long)N * C;                  // rows per sample

  if (warp == CF_WORKER_WARPS) {
    // =================================== MMA issuer ===================================
    if (lane == 0) {
      const uint32_t id_sp_t = make_idesc_tf32_atmem_bmn(128, CF_BLK);   // A = support hi in TMEM, B = MN-major image
      const uint32_t id_sp_s = make_idesc_tf32_mn(128, CF_BLK);          // A = support lo image, B = MN-major image
      const uint64_t glo0 = make_smem_desc_mn32(smem_u32(Glo), colblk, MN32_GROUP_BYTES);
      int n_img[2] = {0, 0}, n_y[2] = {0, 0}, n_k0 = 0, n_k1 = 0, n_d = 0;
      // Accumulation order matters: the tensor core truncates the fp32 accumulator after every MMA, so every cross term
      // (hi*lo, lo*hi: 2^-11 of the result) is added while the accumulator is still small, and only K/8 full-magnitude
      // accumulations follow (profiles/r1_tc_precision.txt).
      auto gate_block = [&](uint32_t d_tmem, const uint8_t* w_hi, uint32_t watom, uint32_t idesc) {
        const uint32_t a0h = smem_u32(k0_hi), a0l = a0h + atomA, a1h = smem_u32(k1_hi), a1l = a1h + atomA;
        const uint32_t b0h = smem_u32(w_hi), b0l = b0h + 2 * watom, b1h = b0h + watom, b1l = b1h + 2 * watom;
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = ks * 32;
          mma_tf32(d_tmem, make_smem_desc_sw128(a0l + o), make_smem_desc_sw128(b0h + o), idesc, ks == 0 ? 0u : 1u);
          mma_tf32(d_tmem, make_smem_desc_sw128(a0h + o), make_smem_desc_sw128(b0l + o), idesc, 1u);
        }
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = ks * 32;
          mma_tf32(d_tmem, make_smem_desc_sw128(a1l + o), make_smem_desc_sw128(b1h + o), idesc, 1u);
          mma_tf32(d_tmem, make_smem_desc_sw128(a1h + o), make_smem_desc_sw128(b1l + o), idesc, 1u);
        }
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = ks * 32;
          mma_tf32(d_tmem, make_smem_desc_sw128(a0h + o), make_smem_desc_sw128(b0h + o), idesc, 1u);
        }
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t o = ks * 32;
          mma_tf32(d_tmem, make_smem_desc_sw128(a1h + o), make_smem_desc_sw128(b1h + o), idesc, 1u);
        }
      };
      for (int s = 0; s < my_samples; ++s) {
        for (int pass = 0; pass < 2; ++pass) {
          const int Hout = pass == 0 ? 2 * h : h;
          const int dstride = 2 * Hout;                            // accumulator columns per category block
          const uint32_t id_gate = make_idesc_tf32(128, dstride);
          const uint8_t* w_hi = pass == 0 ? Wg_hi : Wc_hi;
          const uint32_t watom = pass == 0 ? WG_ATOM : WC_ATOM;
          auto gate = [&](int cb) {                                // both atoms of block cb are in place
            wait_nth(&bars[CB_K0_FULL], n_k0);
            wait_nth(&bars[CB_K1_FULL], n_k1);
            if (cb == 0) wait_free(&bars[CB_D_FREE], n_d);         // the previous pass's epilogue has read every D block
            fence_after_sync();
            gate_block(tmem_base + (uint32_t)(cb * dstride), w_hi, watom, id_gate);
            mma_commit(&bars[CB_K0_FREE]);
            mma_commit(&bars[CB_K1_FREE]);
            ++n_k0;
            ++n_k1;
          };
          long long tr[24];
          int ntr = 0;
          const bool tracing = a.trace && blockIdx.x == 0 && s == 1;
#define CF_TRM() do { if (tracing && ntr < 24) tr[ntr++] = clock64(); } while (0)
          CF_TRM();
          for (int cb = 0; cb < C; ++cb) {
            const int sl = cb & 1;
            wait_nth(&bars[CB_IMG_FULL + sl], n_img[sl]);          // image slot written
            wait_free(&bars[CB_Y_FREE + sl], n_y[sl]);             // Y1 slot drained by the converters
            fence_after_sync();
            CF_TRM();
            {
              const uint32_t xhi = smem_u32(img + (size_t)sl * a.img_bytes);
              const uint64_t xh0 = make_smem_desc_mn32(xhi, colblk, MN32_GROUP_BYTES);
              const uint64_t xl0 = make_smem_desc_mn32(xhi + a.img_bytes / 2, colblk, MN32_GROUP_BYTES);
              const uint32_t dY = tmem_base + CF_TM_Y + (uint32_t)(sl * CF_BLK);
#pragma unroll 1
              for (int ks = 0; ks < Kp / 8; ++ks) {
                const uint64_t o = (uint64_t)(ks * ((2 * MN32_GROUP_BYTES) >> 4));
                mma_tf32(dY, glo0 + o, xh0 + o, id_sp_s, ks > 0 ? 1u : 0u);
                mma_tf32_atmem(dY, tG + (uint32_t)(ks * 8), xl0 + o, id_sp_t, 1u);
              }
#pragma unroll 1
              for (int ks = 0; ks < Kp / 8; ++ks) {
                const uint64_t o = (uint64_t)(ks * ((2 * MN32_GROUP_BYTES) >> 4));
                mma_tf32_atmem(dY, tG + (uint32_t)(ks * 8), xh0 + o, id_sp_t, 1u);
              }
            }
            mma_commit(&bars[CB_Y_FULL + sl]);
            mma_commit(&bars[CB_IMG_FREE + sl]);
            ++n_img[sl];
            ++n_y[sl];
            CF_TRM();
            if (cb >= 1) gate(cb - 1);
            CF_TRM();
          }
          gate(C - 1);
          mma_commit(&bars[CB_D_FULL]);
          ++n_d;
          CF_TRM();
          if (tracing) {
            for (int i = 1; i < ntr; ++i) printf("M%d %02d %lld\n", pass, i, tr[i] - tr[i - 1]);
          }
        }
      }
    }
  } else {
    // =================================== workers ===================================
    const int q4 = warp & 3, half = warp >> 2;
    const int m = q4 * 32 + lane;                       // this thread's TMEM lane = node within the sample
    const bool mlive = m < N;
    const uint32_t tl = tmem_base + ((uint32_t)(q4 * 32) << 16);
    // loader map: chunk q of rows n0 + 32 i
    const int lq = tid & 7, ln0 = tid >> 3;
    int n_img[2] = {0, 0}, n_y[2] = {0, 0}, n_k0 = 0, n_k1 = 0, n_d = 0;

    for (int s = 0; s < my_samples; ++s) {
      const long long b = blockIdx.x + (long long)s * gridDim.x;
      const float* xs = a.xt + b * a.xt_bs;                       // [N][C][Din]
      const long long row_b = b * RC;                             // first (node, category) row of the sample
      for (int pass = 0; pass < 2; ++pass) {
        const int Hout = pass == 0 ? 2 * h : h;
        const int dstride = 2 * Hout;
        const float* hsrc = (pass == 0 ? a.h_prev : a.rH) + row_b * h;   // [N][C][h]; pass 1 reads the r*H this CTA wrote
        float* y1h = (pass == 0 ? a.Yh1 : a.Yr1) + row_b * h;

        float4 raw[4];           // a block's values in flight from HBM / L2 (issued one block ahead)
        float4 bh4[4], bl4[4];   // the block's hi / lo values between its two halves (image first, k = 0 atom later)
        auto issue_loads = [&](int cb) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = ln0 + 32 * i;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) {
              if (lq < 4) {   // pass 1 reads r*H written by this CTA a moment ago: L2-coherent load
                v = __ldcg(reinterpret_cast<const float4*>(hsrc + ((long long)n * C + cb) * h + lq * 4));
              } else {
                const int xi = (lq - 4) * 4;
                const float* xp = xs + ((long long)n * C + cb) * Din + xi;
                if (xi + 3 < Din && (Din & 3) == 0) {
                  v = *reinterpret_cast<const float4*>(xp);
                } else {
                  if (xi < Din) v.x = xp[0];
                  if (xi + 1 < Din) v.y = xp[1];
                  if (xi + 2 < Din) v.z = xp[2];
                  if (xi + 3 < Din) v.w = xp[3];
                }
              }
            }
            raw[i] = v;
          }
        };
        auto load_image = [&](int cb) {   // consume `raw`
          const int sl = cb & 1;
          wait_free(&bars[CB_IMG_FREE + sl], n_img[sl]);          // spatial MMAs of the block two back have read the slot
          uint8_t* ih = img + (size_t)sl * a.img_bytes;
          uint8_t* il = ih + a.img_bytes / 2;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = ln0 + 32 * i;
            const float4 v = raw[i];
            split_tf32(v.x, bh4[i].x, bl4[i].x); split_tf32(v.y, bh4[i].y, bl4[i].y);
            split_tf32(v.z, bh4[i].z, bl4[i].z); split_tf32(v.w, bh4[i].w, bl4[i].w);
            if (n < N) {
              const uint32_t io = mn32_chunk_offset(n, lq);
              *reinterpret_cast<float4*>(ih + io) = bh4[i];
              *reinterpret_cast<float4*>(il + io) = bl4[i];
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[CB_IMG_FULL + sl]);
          ++n_img[sl];
        };
        auto store_k0 = [&]() {
          wait_free(&bars[CB_K0_FREE], n_k0);                     // gate MMAs of the previous block have read the atom
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int n = ln0 + 32 * i;
            if (n < N) {
              const uint32_t ko = atom_chunk_offset(n, lq);
              *reinterpret_cast<float4*>(k0_hi + ko) = bh4[i];
              *reinterpret_cast<float4*>(k0_hi + atomA + ko) = bl4[i];
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[CB_K0_FULL]);
          ++n_k0;
        };
        auto convert_block = [&](int cb) {   // Y1 block: TMEM -> saved tensor + k = 1 atom; thread = (node m, 16 columns)
          const int sl = cb & 1;
          wait_nth(&bars[CB_Y_FULL + sl], n_y[sl]);
          wait_free(&bars[CB_K1_FREE], n_k1);                     // k = 1 MMAs of the previous block have read the atom
          fence_after_sync();
          uint32_t y[16];
          tmem_ld16_async(tl + CF_TM_Y + (uint32_t)(sl * CF_BLK + half * 16), y);
          tmem_ld_wait();
          tmem_ld_pin16(y);
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[CB_Y_FREE + sl]);
          ++n_y[sl];
          if (mlive) {
            if (half == 0) {            // columns 0..15 = the h-part
              float4* dst = reinterpret_cast<float4*>(y1h + ((long long)m * C + cb) * h);
#pragma unroll
              for (int g = 0; g < 4; ++g)
                dst[g] = make_float4(__uint_as_float(y[4 * g]), __uint_as_float(y[4 * g + 1]), __uint_as_float(y[4 * g + 2]),
                                     __uint_as_float(y[4 * g + 3]));
            } else if (pass == 0) {     // columns 16..16+Din = the x-part (identical in both passes: stored once)
              float* dst = a.Yx1 + (row_b + (long long)m * C + cb) * Din;
              if ((Din & 3) == 0) {
                for (int g = 0; g * 4 < Din; ++g)
                  reinterpret_cast<float4*>(dst)[g] = make_float4(__uint_as_float(y[4 * g]), __uint_as_float(y[4 * g + 1]),
                                                                  __uint_as_float(y[4 * g + 2]), __uint_as_float(y[4 * g + 3]));
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < Din) dst[i] = __uint_as_float(y[i]);
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float4 hi, lo;
            split_tf32(__uint_as_float(y[4 * g]), hi.x, lo.x); split_tf32(__uint_as_float(y[4 * g + 1]), hi.y, lo.y);
            split_tf32(__uint_as_float(y[4 * g + 2]), hi.z, lo.z); split_tf32(__uint_as_float(y[4 * g + 3]), hi.w, lo.w);
            const uint32_t ko = atom_chunk_offset(m, half * 4 + g);
            *reinterpret_cast<float4*>(k1_hi + ko) = hi;
            *reinterpret_cast<float4*>(k1_hi + atomA + ko) = lo;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[CB_K1_FULL]);
          ++n_k1;
        };

        // image of block cb first (the spatial GEMM can start), then the previous block's Y1 -> k = 1 atom (its gate GEMM
        // can start), and only then this block's k = 0 atom, whose buffer that gate GEMM is still reading
        long long tr[24];
        int ntr = 0;
        const bool tracing = a.trace && blockIdx.x == 0 && s == 1 && tid == 0;
#define CF_TR() do { if (tracing && ntr < 24) tr[ntr++] = clock64(); } while (0)
        CF_TR();
        issue_loads(0);
        load_image(0);
        CF_TR();
        store_k0();
        CF_TR();
        if (C > 1) issue_loads(1);
        for (int cb = 1; cb < C; ++cb) {
          load_image(cb);
          CF_TR();
          if (cb + 1 < C) issue_loads(cb + 1);   // in flight while Y1 is converted and the gate GEMM drains
          convert_block(cb - 1);
          CF_TR();
          store_k0();
          CF_TR();
        }
        convert_block(C - 1);
        CF_TR();

        // ---------------- epilogue: thread = (node m, 8-channel chunks of its half) over every output category ----------------
        wait_nth(&bars[CB_D_FULL], n_d);
        ++n_d;
        fence_after_sync();
        CF_TR();
        const int nchunk = Hout / 16;                    // 8-channel chunks per thread (gates 2, candidate 1)
        for (int ch = 0; ch < nchunk; ++ch) {
          const int o0 = half * (Hout / 2) + ch * 8;     // first output channel of this chunk
          const bool need_h = pass == 1 || o0 >= h;      // the r half and the candidate read H (the candidate also u)
          const int oh = pass == 0 ? o0 - h : o0;        // channel offset inside the h-wide state tensors
          float d1[CF_MAXC][8];
          {
            uint32_t t[CF_MAXC][8];
#pragma unroll
            for (int cp = 0; cp < CF_MAXC; ++cp)
              if (cp < C) tmem_ld8_async(tl + (uint32_t)(cp * dstride + Hout + o0), t[cp]);
            tmem_ld_wait();
#pragma unroll
            for (int cp = 0; cp < CF_MAXC; ++cp)
              if (cp < C) {
                tmem_ld_pin8(t[cp]);
#pragma unroll
                for (int i = 0; i < 8; ++i) d1[cp][i] = __uint_as_float(t[cp][i]);
              }
          }
          if (mlive) {   // pre-mix partial outputs P_1 (backward forms dGc from them)
#pragma unroll
            for (int cp = 0; cp < CF_MAXC; ++cp)
              if (cp < C) {
                float* ps = (pass == 0 ? a.Pg : a.Pc) + (row_b + (long long)m * C + cp) * Hout + o0;
                reinterpret_cast<float4*>(ps)[0] = make_float4(d1[cp][0], d1[cp][1], d1[cp][2], d1[cp][3]);
                reinterpret_cast<float4*>(ps)[1] = make_float4(d1[cp][4], d1[cp][5], d1[cp][6], d1[cp][7]);
              }
          }
          // operands of category d are fetched one category ahead (TMEM block, H and u pieces)
          uint32_t tn[8];
          float4 hn0 = make_float4(0.f, 0.f, 0.f, 0.f), hn1 = hn0, un0 = hn0, un1 = hn0;
          auto fetch = [&](int d) {
            tmem_ld8_async(tl + (uint32_t)(d * dstride + o0), tn);
            if (mlive && need_h) {
              const long long o = (row_b + (long long)m * C + d) * h + oh;
              hn0 = *reinterpret_cast<const float4*>(a.h_prev + o);
              hn1 = *reinterpret_cast<const float4*>(a.h_prev + o + 4);
              if (pass == 1) {
                un0 = __ldcg(reinterpret_cast<const float4*>(a.u + o));
                un1 = __ldcg(reinterpret_cast<const float4*>(a.u + o + 4));
              }
            }
          };
          fetch(0);
#pragma unroll
          for (int d = 0; d < CF_MAXC; ++d) {
            if (d < C) {
              tmem_ld_wait();
              tmem_ld_pin8(tn);
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(tn[i]);
              const float4 h0 = hn0, h1 = hn1, u0 = un0, u1 = un1;
              if (d + 1 < C) fetch(d + 1);
#pragma unroll
              for (int cp = 0; cp < CF_MAXC; ++cp) {
                if (cp < C) {
                  const float w = Qs[cp * C + d];
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = fmaf(w, d1[cp][i], v[i]);
                }
              }
              if (!mlive) continue;
              const long long row = row_b + (long long)m * C + d;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float pre = v[i] + bias_s[(pass == 0 ? 0 : 2 * h) + o0 + i];
                if (a.act == STC_ACT_RELU) pre = fmaxf(pre, 0.f);
                v[i] = pre;
              }
              if (pass == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = sigmoidf_fast(v[i]);
                if (o0 < h) {
                  float4* dst = reinterpret_cast<float4*>(a.u + row * h + o0);
                  dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                  dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
                  const long long o = row * h + oh;
                  float4* dr = reinterpret_cast<float4*>(a.r + o);
                  dr[0] = make_float4(v[0], v[1], v[2], v[3]);
                  dr[1] = make_float4(v[4], v[5], v[6], v[7]);
                  float4* drh = reinterpret_cast<float4*>(a.rH + o);
                  drh[0] = make_float4(v[0] * h0.x, v[1] * h0.y, v[2] * h0.z, v[3] * h0.w);
                  drh[1] = make_float4(v[4] * h1.x, v[5] * h1.y, v[6] * h1.z, v[7] * h1.w);
                }
              } else {
                const long long o = row * h + o0;
                const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                const float hp[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                float cc[8], hn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  cc[i] = tanhf_fast(v[i]);
                  hn[i] = fmaf(uu[i], cc[i] - hp[i], hp[i]);
                }
                float4* dc = reinterpret_cast<float4*>(a.c + o);
                dc[0] = make_float4(cc[0], cc[1], cc[2], cc[3]);
                dc[1] = make_float4(cc[4], cc[5], cc[6], cc[7]);
                float4* dh = reinterpret_cast<float4*>(a.h_out + o);
                dh[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                dh[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
              }
            }
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[CB_D_FREE]);
        // pass 0 wrote u and r*H with plain global stores; pass 1 (other threads of this CTA) reads them back
        __threadfence_block();
        CF_TR();
        worker_sync();
        CF_TR();
        if (tracing) {
          for (int i = 1; i < ntr; ++i) printf("W%d %02d %lld\n", pass, i, tr[i] - tr[i - 1]);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

static bool aligned16f(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Shape test shared by forward dispatch (backward keeps using the multi-kernel path on the same saved tensors).
bool cell_fused_eligible(const StcDims& d, const StcSupport& gs) {
  // Opt-in (STC_ENABLE_FUSED=1): measured on B200 at B = 4096 the single-sample-in-flight pipeline below takes 2265 us
  // per cell against ~1250 us for the multi-kernel forward (profiles/r1e_fused_fwd.txt); the default path stays
  // multi-kernel until the fused pipeline keeps two samples in flight.
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("STC_ENABLE_FUSED");
    const char* t = getenv("STC_DISABLE_TC");
    enabled = (e && e[0] && e[0] != '0' && !(t && t[0] && t[0] != '0')) ? 1 : 0;
  }
  if (!enabled) return false;
  return gs.kind == STC_SUPPORT_DENSE && d.Ks == 2 && d.Kc == 2 && d.h == 16 && d.Din >= 1 && d.Din <= 16 && d.C >= 1 &&
         d.C <= CF_MAXC && d.N >= 8 && d.N <= 104;
}

int launch_cell_fwd_fused(const StcDims& d, const StcSupport& gs, const float* Q, const float* xt, long long xt_bs,
                          const float* h_prev, const float* Wg, const float* bg, const float* Wc, const float* bc,
                          float* h_out, float* ws, const WsLayout& w, cudaStream_t st) {
  CellFusedArgs a;
  a.B = d.B; a.N = d.N; a.C = d.C; a.Din = d.Din; a.h = d.h; a.act = d.act; a.has_bias = d.has_bias;
  a.Kp = (d.N + 7) & ~7;
  a.trace = getenv("STC_FUSED_TRACE") ? 1 : 0;
  a.Gs = gs.vals;
  a.Q1 = Q + (size_t)d.C * d.C;
  a.xt = xt; a.xt_bs = xt_bs; a.h_prev = h_prev;
  a.Wg = Wg; a.bg = bg; a.Wc = Wc; a.bc = bc; a.h_out = h_out;
  const size_t Rh = w.R * d.h;
  a.u = ws + w.u; a.r = ws + w.r; a.c = ws + w.c;
  a.rH = ws + w.Yr; a.Yr1 = ws + w.Yr + Rh; a.Yh1 = ws + w.Yh; a.Yx1 = ws + w.Yx;
  a.Pg = ws + w.Pg; a.Pc = ws + w.Pc;
  if (!aligned16f(h_prev) || !aligned16f(h_out) || !aligned16f(ws) || ((d.Din & 3) == 0 && (!aligned16f(xt) || (xt_bs & 3)))) {
    set_error("fused cell kernel needs 16-byte aligned state tensors");
    return STC_ERR_BAD_ARG;
  }
  a.img_bytes = 2u * (uint32_t)a.Kp * ATOM_ROW_BYTES;     // hi + lo of one slot
  size_t o = 0;
  a.off_glo = (uint32_t)o; o += (size_t)4 * a.Kp * ATOM_ROW_BYTES;
  o = round_up(o, 1024);
  a.off_wg = (uint32_t)o; o += 4 * (size_t)64 * ATOM_ROW_BYTES;
  a.off_wc = (uint32_t)o; o += 4 * (size_t)32 * ATOM_ROW_BYTES;
  o = round_up(o, 1024);
  a.off_img = (uint32_t)o; o += 2 * (size_t)a.img_bytes;
  o = round_up(o, 1024);
  a.off_k0 = (uint32_t)o; o += 2 * (size_t)128 * ATOM_ROW_BYTES;
  a.off_k1 = (uint32_t)o; o += 2 * (size_t)128 * ATOM_ROW_BYTES;
  a.off_misc = (uint32_t)o; o += 4 * (32 + 64);
  o = round_up(o, 16);
  a.off_bar = (uint32_t)o; o += 8 * CB_COUNT + 16;
  a.smem_bytes = (uint32_t)o;
  if (a.smem_bytes > 227 * 1024) {
    set_error("fused cell kernel does not fit shared memory (%u B)", a.smem_bytes);
    return STC_ERR_UNSUPPORTED;
  }
  STC_TRY(set_smem(tc_cell_fwd_fused_kernel, a.smem_bytes));
  int grid = device_sm_count();
  if (grid > d.B) grid = d.B;
  const double R = (double)w.R;
  ScopedKernelTimer _t(KK_TC_CELL_FWD, st, 4.0 * R * (3.0 * d.Din + 12.0 * d.h) + 4.0 * (d.N * d.N + 12.0 * (d.Din + d.h) * d.h));
  tc_cell_fwd_fused_kernel<<<grid, CF_THREADS, a.smem_bytes, st>>>(a);
  STC_LAUNCH_OK("tc_cell_fwd_fused_kernel");
  return STC_OK;
}

}  // namespace stc
