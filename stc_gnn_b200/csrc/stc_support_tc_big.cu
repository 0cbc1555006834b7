// tcgen05 / TMEM spatial-support kernel for a DENSE learned support that does not fit one tile (N > 128): the
// N x N softmax support MGP_Gen produces (/root/reference/framework/STC_GNN.py:231-243) applied to a large graph.
//
//  tc_support_big_kernel   Y[b,m,:] = alpha * sum_n A(m,n) X[b,n,:] + beta * Z[b,m,:],  A = Gs^T (forward mode product
//                          'bncl,nm->bmcl', STC_GNN.py:37) or A = Gs (its adjoint) -- the operator of tc_support_kernel
//                          (stc_support_tc.cu), tiled over output nodes and input nodes.
//      A tile is 128 output nodes x 128 flattened (sample, feature) columns (two 64-column halves, each with its own
//      accumulator pair); the input nodes are walked in K segments of 64.  Same operand forms as the one-tile kernel (both
//      validated there): the support block is the A operand in TENSOR MEMORY (lane = output node, column = input node,
//      hi | lo) and serves both halves, X is the MN-major B operand in shared memory (SWIZZLE_128B_BASE32B, rows contiguous
//      in the feature index), 3xTF32 with the cross terms in their own accumulator.  Five roles, mbarrier-pipelined:
//        warp 0        issues the MMAs (one elected lane): per K-step  [main | cross] (+)= A_hi x [X_hi | X_lo]  and
//                      cross += A_lo x X_hi;
//        warps 1-4     stream X half-segments (16-byte loads two items ahead in registers, hi/lo split, swizzled stores
//                      into a ring of up to 5 stages);
//        warps 5-8     stream the support block of the segment into a 2-deep TMEM ring (tcgen05.st; loads one segment
//                      ahead in registers -- the block is read from L2: the whole support is re-used by every column tile);
//        warps 9-12    drain a half's accumulators every TB_GROUP segments into the fp32 staging tile in shared memory
//                      (bounded tensor-core accumulation chains of <= 512 input nodes, profiles/r1_tc_precision.txt) while
//                      the MMA warp works on the other half, and after the last segment write the tile with 16-byte
//                      row-segment stores.
//      TMEM: 2 halves x [main 64 | cross 64] accumulator columns + 2 x [hi 64 | lo 64] support columns = 512.
//      First version (one 64-column half per support block, running sums in registers): 96-106 TFLOP/s useful; its ncu
//      source page showed the support stagers as the busiest role (~950 instructions per warp and block) and the waiting
//      epilogue warps spinning on try_wait in the same schedulers (profiles/r4k_tc_support_big_ncu.txt).
#include "stc_conv_common.cuh"
#include "stc_tc.cuh"

#include <stdlib.h>

namespace stc {

using namespace tc;

constexpr int TB_NT = 64;                    // (b,j) columns per half tile = GEMM N
constexpr int TB_HALVES = 2;                 // halves of a tile sharing one support block
constexpr int TB_TW = TB_NT * TB_HALVES;     // tile width in columns
constexpr int TB_KS = 64;                    // input nodes per K segment
constexpr int TB_XW = 4, TB_AW = 4, TB_EW = 4;   // X-producer / support-stager / epilogue warps
constexpr int TB_THREADS = 32 * (1 + TB_XW + TB_AW + TB_EW);
constexpr int TB_SLOTS = 8;                  // 16-byte chunks per X-producer thread per item (64 rows x 64 columns)
constexpr int TB_MAX_STAGES = 5;             // X ring depth
constexpr int TB_ACC_COLS = TB_HALVES * 2 * TB_NT;   // per half: main + cross-term
constexpr int TB_A_COLS = 2 * TB_KS;         // one support buffer: hi | lo
constexpr int TB_OLD = TB_TW + 4;            // row stride (floats) of the staging tile: conflict-free by row and by chunk
constexpr int TB_GROUP = 8;                  // K segments per accumulation chain

struct TcSupBigPlan {
  int N, W, transpose, g_vec, stages, nseg, nmt;
  long long total_cols, ntiles;
  uint32_t off_x, off_o, off_bar, smem_bytes, imgX;
};

__global__ void __launch_bounds__(TB_THREADS, 1)
tc_support_big_kernel(const float* __restrict__ G, const float* __restrict__ X, long long x_bs, const float* Z,
                      long long z_bs, float* Y, float alpha, float beta, const TcSupBigPlan p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, W = p.W;
  uint8_t* Xbuf = smem + p.off_x;               // [stages][hi | lo][2 column blocks][TB_KS][128 B]
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smem + p.off_bar);   // [stages] X producers -> MMA
  uint64_t* xempty = xfull + TB_MAX_STAGES;                          // [stages] MMA -> X producers
  uint64_t* afull = xempty + TB_MAX_STAGES;                          // [2] support stagers -> MMA
  uint64_t* aempty = afull + 2;                                      // [2] MMA -> support stagers
  uint64_t* accfull = aempty + 2;                                    // [halves] MMA -> epilogue
  uint64_t* accempty = accfull + TB_HALVES;                          // [halves] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + TB_HALVES);
  const uint32_t colblk = (uint32_t)TB_KS * ATOM_ROW_BYTES;

  if (tid == 0) {
    for (int i = 0; i < TB_MAX_STAGES; ++i) {
      mbar_init(&xfull[i], TB_XW);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], TB_AW);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < TB_HALVES; ++i) {
      mbar_init(&accfull[i], 1);
      mbar_init(&accempty[i], TB_EW);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512u);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const long long stride = gridDim.x;
  const int nseg = p.nseg;
  const int ngroups = (nseg + TB_GROUP - 1) / TB_GROUP;
  // tile = column tile * nmt + node tile (node tile fastest: the CTAs running together share a few X column slabs and
  // sweep the support once).  The MMA warp and the support stagers walk (tile, segment); the X producers walk
  // (tile, segment, half); the epilogue walks (tile, chain group, half).

  if (uniform_warp_index() == 0) {
    // =========================== MMA issuer (one elected lane of a converged warp) ===========================
    const uint32_t tb = uniform_u32(*tmem_slot);
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc_tf32_atmem_bmn(128, TB_NT);
      const uint32_t idesc2 = make_idesc_tf32_atmem_bmn(128, 2 * TB_NT);   // [X_hi | X_lo] -> [main | cross]
      int xi = 0, ai = 0, gi = 0;   // X items, support items, chain groups issued so far
      for (long long tile = blockIdx.x; tile < p.ntiles; tile += stride) {
        for (int g = 0; g < ngroups; ++g, ++gi) {
          const int s0 = g * TB_GROUP, s1 = min(nseg, s0 + TB_GROUP);
          for (int seg = s0; seg < s1; ++seg, ++ai) {
            const int ba = ai & 1;
            mbar_wait(&afull[ba], (uint32_t)(ai >> 1) & 1u);
            const uint32_t tA = tb + (uint32_t)(TB_ACC_COLS + ba * TB_A_COLS);
            const int kleft = N - seg * TB_KS;
            const int nks = kleft >= TB_KS ? TB_KS / 8 : (kleft + 7) / 8;   // the rows past N are zeros on both sides
#pragma unroll 1
            for (int half = 0; half < TB_HALVES; ++half, ++xi) {
              if (seg == s0) mbar_wait(&accempty[half], ((uint32_t)gi & 1u) ^ 1u);   // the epilogue drained this half
              const int st = xi % p.stages;
              mbar_wait(&xfull[st], (uint32_t)(xi / p.stages) & 1u);
              fence_after_sync();
              const uint32_t xhi = smem_u32(Xbuf + (size_t)st * 2 * p.imgX);
              const uint64_t xh0 = make_smem_desc_mn32(xhi, colblk, MN32_GROUP_BYTES);
              const uint32_t d_main = tb + (uint32_t)(half * 2 * TB_NT), d_small = d_main + TB_NT;
#pragma unroll 1
              for (int ks = 0; ks < nks; ++ks) {
                const uint64_t o = (uint64_t)(ks * ((2 * MN32_GROUP_BYTES) >> 4));   // K = 8 rows further down
                const uint32_t gh = tA + (uint32_t)(ks * 8), gl = gh + (uint32_t)TB_KS;
                mma_tf32_atmem(d_main, gh, xh0 + o, idesc2, (seg > s0 || ks > 0) ? 1u : 0u);
                mma_tf32_atmem(d_small, gl, xh0 + o, idesc, 1u);
              }
              mma_commit(&xempty[st]);                        // the X stage may be refilled once these MMAs have read it
              if (seg == s1 - 1) mma_commit(&accfull[half]);  // this half's chain is complete
            }
            mma_commit(&aempty[ba]);                          // both halves have read the support block
          }
        }
      }
    }
  } else if (warp <= TB_XW) {
    // =========================== X producers ===========================
    const int pt = tid - 32;
    const int c0 = (pt & 15) << 2, r0 = pt >> 4;       // chunk column of the 64-wide half, rows r0 + 8 i of the segment
    const uint32_t soff0 = (uint32_t)(c0 >> 5) * colblk + mn32_chunk_offset(r0, (c0 & 31) >> 2);
    float4 ra[2][TB_SLOTS];
    auto advance = [&](long long& tile, int& seg, int& half) {
      if (++half == TB_HALVES) {
        half = 0;
        if (++seg == nseg) {
          seg = 0;
          tile += stride;
        }
      }
    };
    // the column (sample, feature) of this thread's chunk only changes with the tile: its two source pointers (one per
    // half) are recomputed there, not per item (two 64-bit divisions per fetch were a third of this role's instructions)
    long long src_tile = -1;
    const float* src_h[TB_HALVES] = {nullptr, nullptr};
    auto fetch = [&](long long tile, int seg, int half, float4 (&r)[TB_SLOTS]) {
      if (tile != src_tile) {
        src_tile = tile;
        const long long cg0 = (tile / p.nmt) * TB_TW + c0;
#pragma unroll
        for (int hf = 0; hf < TB_HALVES; ++hf) {
          const long long cg = cg0 + hf * TB_NT;
          const bool ok = tile < p.ntiles && cg < p.total_cols;   // W % 4 == 0: a chunk never straddles samples
          const long long b = ok ? cg / W : 0;
          src_h[hf] = ok ? X + b * x_bs + (cg - b * W) : nullptr;
        }
      }
      const float* src = half == 0 ? src_h[0] : src_h[1];
      const int n0 = seg * TB_KS + r0;
      if (src != nullptr && n0 + 8 * (TB_SLOTS - 1) < N) {        // every row of the item in range: no per-row predicates
        const float* s0 = src + (long long)n0 * W;
#pragma unroll
        for (int i = 0; i < TB_SLOTS; ++i) r[i] = __ldg(reinterpret_cast<const float4*>(s0 + (long long)(8 * i) * W));
      } else {
#pragma unroll
        for (int i = 0; i < TB_SLOTS; ++i) {
          const int n = n0 + 8 * i;
          r[i] = (src != nullptr && n < N) ? __ldg(reinterpret_cast<const float4*>(src + (long long)n * W))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto stage = [&](int xi, const float4 (&r)[TB_SLOTS]) {
      const int st = xi % p.stages;
      mbar_wait(&xempty[st], ((uint32_t)(xi / p.stages) & 1u) ^ 1u);
      uint8_t* hi = Xbuf + (size_t)st * 2 * p.imgX;
      uint8_t* lo = hi + p.imgX;
#pragma unroll
      for (int i = 0; i < TB_SLOTS; ++i) store_split4(hi, lo, soff0 + (uint32_t)(8 * i) * ATOM_ROW_BYTES, r[i]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&xfull[st]);
    };
    long long tcur = blockIdx.x, tf = blockIdx.x;   // item being staged / item being fetched
    int sc = 0, hc = 0, sf = 0, hf = 0;
    fetch(tf, sf, hf, ra[0]);
    advance(tf, sf, hf);
    fetch(tf, sf, hf, ra[1]);
    advance(tf, sf, hf);
    for (int xi = 0; tcur < p.ntiles; xi += 2) {   // an even number of items per tile: both slots always end together
      stage(xi, ra[0]);
      fetch(tf, sf, hf, ra[0]);
      advance(tf, sf, hf);
      advance(tcur, sc, hc);
      stage(xi + 1, ra[1]);
      fetch(tf, sf, hf, ra[1]);
      advance(tf, sf, hf);
      advance(tcur, sc, hc);
    }
  } else if (warp <= TB_XW + TB_AW) {
    // =========================== support stagers: global (L2) -> registers -> tensor memory ===========================
    const int q = warp & 3;                          // TMEM lane quarter this warp may write
    const int ml = q * 32 + lane;                    // output node within the node tile
    const uint32_t tl = tmem_base + (uint32_t)TB_ACC_COLS + ((uint32_t)(q * 32) << 16);
    float ga[TB_KS];
    long long m_tile = -1;
    int m_cached = 0;
    auto fetch_a = [&](long long tile, int seg) {
      if (tile != m_tile) {     // the 64-bit modulo only when the tile changes
        m_tile = tile;
        m_cached = (int)(tile % p.nmt) * 128 + ml;
      }
      const int m = m_cached;
      const int k0 = seg * TB_KS;
      const bool ok = m < N;
      if (ok && k0 + TB_KS <= N) {   // whole segment in range (every segment but the last): no per-element predicates
        if (p.transpose) {           // A(m,k) = Gs[k][m]: the lanes of a warp read consecutive floats
          const float* gp = G + (size_t)k0 * N + m;
#pragma unroll
          for (int i = 0; i < TB_KS; ++i) ga[i] = __ldg(gp + (size_t)i * N);
        } else if (p.g_vec) {        // A(m,k) = Gs[m][k]: a lane reads its own row, 16 bytes at a time
          const float4* gp = reinterpret_cast<const float4*>(G + (size_t)m * N + k0);
#pragma unroll
          for (int i = 0; i < TB_KS / 4; ++i) {
            const float4 v = __ldg(gp + i);
            ga[4 * i] = v.x; ga[4 * i + 1] = v.y; ga[4 * i + 2] = v.z; ga[4 * i + 3] = v.w;
          }
        } else {
          const float* gp = G + (size_t)m * N + k0;
#pragma unroll
          for (int i = 0; i < TB_KS; ++i) ga[i] = __ldg(gp + i);
        }
      } else if (p.transpose) {
#pragma unroll
        for (int i = 0; i < TB_KS; ++i) ga[i] = (ok && k0 + i < N) ? __ldg(G + (size_t)(k0 + i) * N + m) : 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < TB_KS; ++i) ga[i] = (ok && k0 + i < N) ? __ldg(G + (size_t)m * N + k0 + i) : 0.f;
      }
    };
    long long tile = blockIdx.x;
    int seg = 0;
    fetch_a(tile, seg);
    for (int it = 0; tile < p.ntiles; ++it) {
      const int ba = it & 1;
      mbar_wait(&aempty[ba], ((uint32_t)(it >> 1) & 1u) ^ 1u);   // the MMAs of the item two back have read this buffer
      fence_after_sync();
      const uint32_t ta = tl + (uint32_t)(ba * TB_A_COLS);
#pragma unroll
      for (int c = 0; c < TB_KS; c += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_tf32(ga[c + i], hi[i], lo[i]);
        tmem_st8(ta + (uint32_t)c, hi);
        tmem_st8(ta + (uint32_t)(TB_KS + c), lo);
      }
      tmem_st_wait();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[ba]);
      if (++seg == nseg) {
        seg = 0;
        tile += stride;
      }
      if (tile < p.ntiles) fetch_a(tile, seg);     // in flight while the MMA warp works through the other buffer
    }
  } else {
    // =========================== epilogue: TMEM -> staging tile (running sums) -> coalesced rows ===========================
    const int sp = warp & 3;                         // TMEM lane quarter this warp may read
    const int ml = sp * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(sp * 32) << 16);
    float* Obuf = reinterpret_cast<float*>(smem + p.off_o);
    float* orow = Obuf + ml * TB_OLD;
    const int et = tid - 32 * (1 + TB_XW + TB_AW);
    const int ch = et & 31, er0 = et >> 5;           // global side: 16-byte chunk `ch` of rows er0 + 4 i
    const bool use_z = beta != 0.f;
    int gi = 0;
    for (long long tile = blockIdx.x; tile < p.ntiles; tile += stride) {
      const long long ct = tile / p.nmt;
      const int m0 = (int)(tile - ct * p.nmt) * 128;
      for (int g = 0; g < ngroups; ++g, ++gi) {
#pragma unroll 1
        for (int half = 0; half < TB_HALVES; ++half) {
          mbar_wait_backoff(&accfull[half], (uint32_t)gi & 1u, 128u);   // whole chains: always with back-off (stc_tc.cuh)
          fence_after_sync();
#pragma unroll
          for (int hh = 0; hh < TB_NT / 16; ++hh) {
            uint32_t vm[16], vs[16];
            const uint32_t a0 = tl + (uint32_t)(half * 2 * TB_NT + hh * 16);
            tmem_ld16_async(a0, vm);
            tmem_ld16_async(a0 + TB_NT, vs);
            tmem_ld_wait();
            tmem_ld_pin16(vm);
            tmem_ld_pin16(vs);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float4* slot = reinterpret_cast<float4*>(orow + half * TB_NT + hh * 16 + i);
              float4 v = make_float4(__uint_as_float(vm[i]) + __uint_as_float(vs[i]),
                                     __uint_as_float(vm[i + 1]) + __uint_as_float(vs[i + 1]),
                                     __uint_as_float(vm[i + 2]) + __uint_as_float(vs[i + 2]),
                                     __uint_as_float(vm[i + 3]) + __uint_as_float(vs[i + 3]));
              if (g > 0) {
                const float4 o = *slot;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              *slot = v;
            }
          }
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&accempty[half]);   // the MMA warp may start this half's next chain
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long cgc = ct * TB_TW + 4 * ch;     // this thread's chunk column (never straddles samples: W % 4 == 0)
      if (cgc < p.total_cols) {
        const long long bc = cgc / W;
        const int jc = (int)(cgc - bc * W);
        const int rows = min(128, N - m0);
        float* ydst = Y + (bc * N + m0) * (long long)W + jc;
        const float* zsrc = use_z ? Z + bc * z_bs + (long long)m0 * W + jc : nullptr;
#pragma unroll 4
        for (int r = er0; r < rows; r += 4) {
          float4 o = *reinterpret_cast<const float4*>(Obuf + r * TB_OLD + 4 * ch);
          o.x *= alpha; o.y *= alpha; o.z *= alpha; o.w *= alpha;
          if (use_z) {
            const float4 z = *reinterpret_cast<const float4*>(zsrc + (long long)r * W);   // plain load: Y may alias Z
            o.x = fmaf(beta, z.x, o.x); o.y = fmaf(beta, z.y, o.y); o.z = fmaf(beta, z.z, o.z); o.w = fmaf(beta, z.w, o.w);
          }
          *reinterpret_cast<float4*>(ydst + (long long)r * W) = o;
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the staging tile is rewritten by the next tile
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

static bool aligned16g(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Dense support with N > 128 on the tensor cores; *handled tells the caller (the FFMA tile kernel takes what this
// declines: widths that are not a multiple of 4, unaligned tensors, the fused axpy form).
int try_launch_support_tc_big(const float* G, int N, int B, int width, bool transpose, const float* x, int64_t x_bs,
                              const float* z, int64_t z_bs, float* y, float alpha, float beta, cudaStream_t st,
                              bool* handled) {
  *handled = false;
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("STC_DISABLE_TC");
    const char* e2 = getenv("STC_DISABLE_TC_SUPPORT_BIG");
    disabled = ((e && e[0] && e[0] != '0') || (e2 && e2[0] && e2[0] != '0')) ? 1 : 0;
  }
  if (disabled == 1) return STC_OK;
  if (N <= 128 || width % 4 != 0 || x_bs % 4 != 0 || !aligned16g(x) || !aligned16g(y)) return STC_OK;
  if (beta != 0.f && (z == nullptr || z_bs % 4 != 0 || !aligned16g(z))) return STC_OK;
  TcSupBigPlan p;
  p.N = N;
  p.W = width;
  p.transpose = transpose ? 1 : 0;
  p.g_vec = (!transpose && N % 4 == 0 && aligned16g(G)) ? 1 : 0;
  p.nseg = (N + TB_KS - 1) / TB_KS;
  p.nmt = (N + 127) / 128;
  p.total_cols = (long long)B * width;
  p.ntiles = ((p.total_cols + TB_TW - 1) / TB_TW) * p.nmt;
  p.imgX = (uint32_t)(TB_NT / 32) * TB_KS * ATOM_ROW_BYTES;
  const size_t obytes = (size_t)128 * TB_OLD * sizeof(float);
  const size_t barbytes = 8 * (2 * TB_MAX_STAGES + 4 + 2 * TB_HALVES) + 16;
  p.stages = TB_MAX_STAGES;
  while (p.stages > 2 && 2 * (size_t)p.stages * p.imgX + obytes + barbytes > 227 * 1024) --p.stages;
  size_t o = 0;
  p.off_x = (uint32_t)o; o += 2 * (size_t)p.stages * p.imgX;
  p.off_o = (uint32_t)o; o += obytes;
  p.off_bar = (uint32_t)o; o += barbytes;
  p.smem_bytes = (uint32_t)o;
  STC_TRY(set_smem(tc_support_big_kernel, p.smem_bytes));
  long long grid = device_sm_count();
  if (grid > p.ntiles) grid = p.ntiles;
  // compulsory traffic: read X, write Y (+ read Z) + the support once (its re-reads per column tile are L2 traffic)
  ScopedKernelTimer _t(KK_TC_SUPPORT_BIG, st, 4.0 * B * N * width * (2 + (beta != 0.f ? 1 : 0)) + 4.0 * N * N);
  tc_support_big_kernel<<<(int)grid, TB_THREADS, p.smem_bytes, st>>>(G, x, x_bs, z, z_bs, y, alpha, beta, p);
  STC_LAUNCH_OK("tc_support_big_kernel");
  *handled = true;
  return STC_OK;
}

}  // namespace stc
