// tcgen05 / TMA tensor-core path of the fake-quant convolution (sm_100a only).
//
// Part 1 of this file: two hardware self-tests exported through the C-ABI.  They pin the two
// things that cannot be checked on a CPU-only build box - the UMMA shared-memory / instruction
// descriptor encodings and the TMA box geometry with out-of-bounds zero fill - on tiny problems
// with known answers, under bounded waits (a wrong descriptor must fail a test, not hang a GPU).
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "mnb_common.cuh"
#include "mnb_tc.cuh"

// ------------------------------------------------------------------ host: tensor-map encoding
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

}  // namespace

// rank-`rank` tiled map over a dense tensor; dims/box innermost-first; elem_bytes 4 (f32) / 2 / 1
int mnb_make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                  const uint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return mnb_fail(MNB_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bdim[5], estr[5];
  uint64_t stride = (uint64_t)elem_bytes;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;  // byte stride of dimension i+1
  }
  CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mnb_fail(MNB_E_ARG, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// ------------------------------------------------------------------ self-test 1: UMMA descriptors
// D[128 x N] = A[128 x K] * B[N x K]^T, operands converted by threads into the K-major no-swizzle
// canonical layout  buf[k_chunk][row][16 bytes]  (SBO = 128 B, LBO = rows * 16 B).
template <bool INT8>
__global__ void __launch_bounds__(160) selftest_umma_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int N, int K, int* err) {
  constexpr int EPC = INT8 ? 16 : 8;  // elements per 16-byte chunk
  __shared__ __align__(128) uint8_t a_s[128 * 64 * 2];
  __shared__ __align__(128) uint8_t b_s[128 * 64 * 2];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunk = K / EPC;

  if (tid < 128) {
    for (int kc = 0; kc < nchunk; ++kc) {
      uint8_t* dst = a_s + ((size_t)kc * 128 + tid) * 16;
      for (int e = 0; e < EPC; ++e) {
        float v = A[(size_t)tid * K + kc * EPC + e];
        if (INT8) reinterpret_cast<int8_t*>(dst)[e] = (int8_t)v;
        else reinterpret_cast<__nv_bfloat16*>(dst)[e] = __float2bfloat16_rn(v);
      }
      if (tid < N) {
        uint8_t* db = b_s + ((size_t)kc * N + tid) * 16;
        for (int e = 0; e < EPC; ++e) {
          float v = B[(size_t)tid * K + kc * EPC + e];
          if (INT8) reinterpret_cast<int8_t*>(db)[e] = (int8_t)v;
          else reinterpret_cast<__nv_bfloat16*>(db)[e] = __float2bfloat16_rn(v);
        }
      }
    }
    tc::fence_proxy_async_smem();
  }
  if (warp == 4) {
    tc::tmem_alloc<128>(&tmem_slot);
    if (lane == 0) { tc::mbar_init(&done_bar, 1); tc::fence_barrier_init(); }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 4 && lane == 0) {
    const uint32_t idesc = INT8 ? tc::make_idesc(2, 1, 1, 128, N) : tc::make_idesc(1, 1, 1, 128, N);
    const int ksteps = nchunk / 2;
    for (int ks = 0; ks < ksteps; ++ks) {
      uint64_t ad = tc::smem_desc_kmajor_noswz(tc::smem_u32(a_s) + ks * 2 * 128 * 16, 128 * 16, 128);
      uint64_t bd = tc::smem_desc_kmajor_noswz(tc::smem_u32(b_s) + ks * 2 * N * 16, N * 16, 128);
      if (INT8) tc::mma_i8(tmem, ad, bd, idesc, ks > 0);
      else tc::mma_f16(tmem, ad, bd, idesc, ks > 0);
    }
    tc::mma_commit(&done_bar);
  }
  if (warp < 4) {
    bool ok = tc::mbar_wait(&done_bar, 0, err, 101);
    tc::tc_fence_after();
    if (ok) {
      for (int n0 = 0; n0 < N; n0 += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + n0, r);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32 && n0 + j < N; ++j)
          D[(size_t)(warp * 32 + lane) * N + n0 + j] = INT8 ? (float)(int)r[j] : __uint_as_float(r[j]);
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc<128>(tmem);
}

extern "C" int mnb_selftest_umma(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t int8,
                                 int32_t* err_flag, mnb_stream_t stream) {
  MNB_REQUIRE(A && B && D && err_flag, "NULL self-test pointers");
  MNB_REQUIRE(N >= 16 && N <= 128 && N % 16 == 0, "self-test N must be a multiple of 16 in [16,128]");
  MNB_REQUIRE(K >= 32 && K <= (int8 ? 128 : 64) && K % 32 == 0, "self-test K out of range");
  if (int8) selftest_umma_kernel<true><<<1, 160, 0, (cudaStream_t)stream>>>(A, B, D, N, K, err_flag);
  else selftest_umma_kernel<false><<<1, 160, 0, (cudaStream_t)stream>>>(A, B, D, N, K, err_flag);
  MNB_LAUNCHED(1);
  return 0;
}

// ------------------------------------------------------------------ self-test 2: TMA box + OOB fill
__global__ void __launch_bounds__(128) selftest_tma_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int c1,
                                                           int c2, int box_elems, float* __restrict__ out, int* err) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bar;
  float* buf = reinterpret_cast<float*>(dyn);
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(&bar, (uint32_t)box_elems * 4u);
    tc::tma_load_3d(buf, &tmap, &bar, c0, c1, c2);
  }
  bool ok = tc::mbar_wait(&bar, 0, err, 201);
  if (ok)
    for (int i = threadIdx.x; i < box_elems; i += blockDim.x) out[i] = buf[i];
}

extern "C" int mnb_selftest_tma3d(const float* src, const int64_t* dims3, const int32_t* box3, const int32_t* coord3,
                                  float* out, int32_t* err_flag, mnb_stream_t stream) {
  MNB_REQUIRE(src && dims3 && box3 && coord3 && out && err_flag, "NULL self-test pointers");
  CUtensorMap tmap;
  uint64_t dims[3] = {(uint64_t)dims3[0], (uint64_t)dims3[1], (uint64_t)dims3[2]};
  uint32_t box[3] = {(uint32_t)box3[0], (uint32_t)box3[1], (uint32_t)box3[2]};
  if (int e = mnb_make_tmap(&tmap, src, 4, 3, dims, box)) return e;
  int elems = box3[0] * box3[1] * box3[2];
  MNB_REQUIRE(elems * 4 <= 48 * 1024, "self-test box too large");
  selftest_tma_kernel<<<1, 128, elems * 4, (cudaStream_t)stream>>>(tmap, coord3[0], coord3[1], coord3[2], elems, out,
                                                                    err_flag);
  MNB_LAUNCHED(1);
  return 0;
}

// =====================================================================================
// Part 2: fused fake-quant convolution forward on tcgen05 tensor cores.
//
//   y[b, gNg+n, h, w] = bias + (a_scale * w_scale[n]) * sum_{c,r,s} e_a[b, gCg+c, h+r-pad, w+s-pad] * e_w[n,c,r,s]
//
// Data flow per CTA (persistent, one CTA per SM, bound to a slab of groups whose integer
// weights stay resident in shared memory as bf16):
//
//   TMA warp      : cp.async.bulk.tensor.4d box (W, TH+2pad, CC, TB) of the fp32 NCHW input
//                   -> staging ring (rows outside the image are zero-filled by the TMA unit)
//   8 converter   : staging fp32 -> quantize (DoReFa / IAO level, or exact 3-way bf16 split of a
//     warps         raw fp32 value) -> "position-major" bf16 operand  op[c/8][position][8 ch]
//                   where position = row * (W + 2 pad) + col of the zero-padded tile.  This is
//                   the UMMA K-major no-swizzle canonical layout with positions as GEMM rows, so
//                   a filter tap (r, s) is the SAME buffer with the descriptor start address moved
//                   by (r * BW + s) * 16 bytes: implicit GEMM with no im2col copy at all.
//                   Fused outputs for the backward pass: u8 level codes + STE pass bits.
//   MMA warp      : one thread issues tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), all
//                   operands are exact small integers (or exact bf16 pieces), M = 128 positions,
//                   N = Ng output channels, K = 16 channels per instruction
//   4 epilogue    : tcgen05.ld accumulator rows -> scale / bias -> coalesced fp32 NCHW stores
//     warps
// =====================================================================================
namespace tcconv {

constexpr int NTHREADS = 512;
constexpr int NCONV = 256;  // converter threads (warps 8..15)
constexpr int NEPI = 128;   // epilogue threads (warps 4..7)
constexpr int NST = 3, NOP = 2, NACC = 2;
constexpr int kMaxDynSmem = 227 * 1024 - 2048;  // 227 KB per CTA minus the static barrier block

struct FwdParams {
  int B, C, H, W, K, R, S, pad, G, Cg, Ng;
  int BW, TH, THH, TB, CC, nchunk;
  int npos_in;       // staged positions per channel group (multiple of 8)
  int row_tiles;     // ceil(H / TH)
  int n_tiles;       // ceil(B / TB) * row_tiles
  int slab_groups, n_slabs;
  int mode;          // 0 raw fp32 (exact 3-term bf16 split), else MNB_ACT_DOREFA / MNB_ACT_IAO
  int a_offset;      // e_a = code + a_offset (+ zero_point)
  int tmem_cols;
  int stage_bytes, op_term_bytes, op_buf_bytes, b_group_bytes;
  int off_stage, off_op, off_b;
  mnb_act_qparams qp;
  float a_scale_const;  // DoReFa: 1/(2^a-1); raw: 1
  const int16_t* w_int; const float* w_scale; const float* a_scale; const float* bias;
  float* y; uint8_t* codes; uint32_t* pass_bits; int* err;
};

struct alignas(8) Barriers {
  uint64_t stage_full[NST], stage_empty[NST], op_full[NOP], op_empty[NOP], acc_full[NACC], acc_empty[NACC];
  uint32_t tmem_slot;
  uint32_t op_flags[NOP][8];
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(NTHREADS, 1)
fq_conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Barriers bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* stage_base = smem + p.off_stage;
  uint8_t* op_base = smem + p.off_op;
  uint8_t* b_base = smem + p.off_b;
  const int RS = p.R * p.S, c8_per_group = p.Cg / 8;

  // ---- work assignment: CTA -> slab of groups; tiles of the slab are dealt round-robin
  const int slab = blockIdx.x % p.n_slabs;
  const int rank_in_slab = blockIdx.x / p.n_slabs;
  const int ctas_in_slab = (gridDim.x - slab + p.n_slabs - 1) / p.n_slabs;
  const int g_first = slab * p.slab_groups;
  const int g_count = min(p.slab_groups, p.G - g_first);

  // ---- one-time setup
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { tc::mbar_init(&bar.stage_full[i], 1); tc::mbar_init(&bar.stage_empty[i], NCONV); }
    for (int i = 0; i < NOP; ++i) { tc::mbar_init(&bar.op_full[i], NCONV); tc::mbar_init(&bar.op_empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { tc::mbar_init(&bar.acc_full[i], 1); tc::mbar_init(&bar.acc_empty[i], NEPI); }
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmap_x);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&bar.tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // resident integer weights of this slab: B[gi][tap][c/8][n][8] bf16 (K-major no-swizzle, rows = n)
  {
    const int per_group = RS * p.Cg * p.Ng;
    for (int idx = tid; idx < g_count * per_group; idx += NTHREADS) {
      const int gi = idx / per_group;
      int r = idx - gi * per_group;
      const int n = r / (p.Cg * RS);
      r -= n * (p.Cg * RS);
      const int c = r / RS, tap = r - c * RS;
      const int16_t v = __ldg(p.w_int + ((int64_t)((g_first + gi) * p.Ng + n) * p.Cg + c) * RS + tap);
      uint8_t* dst = b_base + (size_t)gi * p.b_group_bytes +
                     ((size_t)((tap * c8_per_group + (c >> 3)) * p.Ng + n)) * 16 + (c & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn((float)v);
    }
    tc::fence_proxy_async_smem();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = bar.tmem_slot;
  const int n_img_tiles = p.n_tiles;

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = rank_in_slab; tile < n_img_tiles; tile += ctas_in_slab) {
        const int bt = tile / p.row_tiles, rt = tile - bt * p.row_tiles;
        const int b0 = bt * p.TB, h0 = rt * p.TH;
        for (int gi = 0; gi < g_count; ++gi) {
          for (int ch = 0; ch < p.nchunk; ++ch, ++it) {
            const int st = it % NST;
            const uint32_t ph = (it / NST) & 1;
            if (!tc::mbar_wait(&bar.stage_empty[st], ph ^ 1, p.err, 301)) goto done;
            tc::mbar_arrive_expect_tx(&bar.stage_full[st], (uint32_t)p.stage_bytes);
            tc::tma_load_4d(stage_base + (size_t)st * p.stage_bytes, &tmap_x, &bar.stage_full[st], 0, h0 - p.pad,
                            (g_first + gi) * p.Cg + ch * p.CC, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(1, 1, 1, 128, (uint32_t)p.Ng);
      const uint32_t a_lbo = (uint32_t)p.npos_in * 16u, b_lbo = (uint32_t)p.Ng * 16u;
      uint32_t it = 0, item = 0;
      for (int tile = rank_in_slab; tile < n_img_tiles; tile += ctas_in_slab) {
        for (int gi = 0; gi < g_count; ++gi, ++item) {
          const int acc = item % NACC;
          const uint32_t aph = (item / NACC) & 1;
          if (!tc::mbar_wait(&bar.acc_empty[acc], aph ^ 1, p.err, 302)) goto done;
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem + (uint32_t)(acc * p.Ng);
          uint32_t accumulate = 0;
          for (int ch = 0; ch < p.nchunk; ++ch, ++it) {
            const int ob = it % NOP;
            const uint32_t oph = (it / NOP) & 1;
            if (!tc::mbar_wait(&bar.op_full[ob], oph, p.err, 303)) goto done;
            tc::tc_fence_after();
            int nterms = 1;
            if (p.mode == 0) {
              uint32_t any = 0;
#pragma unroll
              for (int w8 = 0; w8 < 8; ++w8) any |= bar.op_flags[ob][w8];
              nterms = any ? 3 : 1;
            }
            const uint32_t op_addr = tc::smem_u32(op_base + (size_t)ob * p.op_buf_bytes);
            const uint32_t b_addr = tc::smem_u32(b_base + (size_t)gi * p.b_group_bytes);
            for (int tap = 0; tap < RS; ++tap) {
              const int r = tap / p.S, s = tap - r * p.S;
              const uint32_t tap_off = (uint32_t)(r * p.BW + s) * 16u;
              for (int j = 0; j < p.CC / 16; ++j) {
                const uint32_t b_start = b_addr + (uint32_t)((tap * c8_per_group + ch * (p.CC / 8) + 2 * j) * p.Ng) * 16u;
                const uint64_t bd = tc::smem_desc_kmajor_noswz(b_start, b_lbo, 128);
                for (int t = 0; t < nterms; ++t) {
                  const uint32_t a_start = op_addr + (uint32_t)t * p.op_term_bytes + (uint32_t)(2 * j) * a_lbo + tap_off;
                  const uint64_t ad = tc::smem_desc_kmajor_noswz(a_start, a_lbo, 128);
                  tc::mma_f16(d_tmem, ad, bd, idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            tc::mma_commit(&bar.op_empty[ob]);  // operand buffer free once these MMAs retire
          }
          tc::mma_commit(&bar.acc_full[acc]);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================================================= epilogue
    const int q = warp - 4;            // TMEM lane quarter of this warp (warp % 4)
    const int pos = q * 32 + lane;     // GEMM row = padded-tile position
    const int tb = pos / (p.THH * p.BW);
    const int rem = pos - tb * (p.THH * p.BW);
    const int th = rem / p.BW, wc = rem - th * p.BW;
    const float a_sc = p.a_scale ? __ldg(p.a_scale) : p.a_scale_const;
    uint32_t item = 0;
    for (int tile = rank_in_slab; tile < n_img_tiles; tile += ctas_in_slab) {
      const int bt = tile / p.row_tiles, rt = tile - bt * p.row_tiles;
      const int b = bt * p.TB + tb, h = rt * p.TH + th;
      const bool valid = tb < p.TB && th < p.TH && wc < p.W && b < p.B && h < p.H;
      for (int gi = 0; gi < g_count; ++gi, ++item) {
        const int acc = item % NACC;
        const uint32_t aph = (item / NACC) & 1;
        if (!tc::mbar_wait(&bar.acc_full[acc], aph, p.err, 304)) goto done;
        tc::tc_fence_after();
        const int ch0 = (g_first + gi) * p.Ng;
        float* yrow = p.y + (((int64_t)b * p.K + ch0) * p.H + h) * p.W + wc;
        for (int n0 = 0; n0 < p.Ng; n0 += 32) {
          uint32_t r[32];
          tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Ng + n0), r);
          tc::tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = n0 + j;
              if (n < p.Ng) {
                const float sc = __fmul_rn(a_sc, __ldg(p.w_scale + ch0 + n));
                const float bs = p.bias ? __ldg(p.bias + ch0 + n) : 0.f;
                yrow[(int64_t)n * p.H * p.W] = __fadd_rn(__fmul_rn(__uint_as_float(r[j]), sc), bs);
              }
            }
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&bar.acc_empty[acc]);
      }
    }
  } else if (warp >= 8) {
    // ================================================================= converters
    const int ct = tid - 256;
    const int cw = ct >> 5;
    MnbActQ q;
    if (p.mode != 0) q = mnb_load_actq(p.qp);
    const int a_off = p.a_offset + ((p.mode == MNB_ACT_IAO && p.qp.zero_point) ? (int)__ldg(p.qp.zero_point) : 0);
    const int per_img = p.THH * p.BW;
    uint32_t it = 0;
    for (int tile = rank_in_slab; tile < n_img_tiles; tile += ctas_in_slab) {
      const int bt = tile / p.row_tiles, rt = tile - bt * p.row_tiles;
      const int b0 = bt * p.TB, h0 = rt * p.TH;
      for (int gi = 0; gi < g_count; ++gi) {
        for (int ch = 0; ch < p.nchunk; ++ch, ++it) {
          const int st = it % NST, ob = it % NOP;
          const uint32_t ph = (it / NST) & 1, oph = (it / NOP) & 1;
          if (!tc::mbar_wait(&bar.stage_full[st], ph, p.err, 305)) goto done;
          if (!tc::mbar_wait(&bar.op_empty[ob], oph ^ 1, p.err, 306)) goto done;
          const float* stg = reinterpret_cast<const float*>(stage_base + (size_t)st * p.stage_bytes);
          uint8_t* opb = op_base + (size_t)ob * p.op_buf_bytes;
          const int cbase = (g_first + gi) * p.Cg + ch * p.CC;
          uint32_t any_low = 0;
          const int total = p.npos_in * (p.CC / 8);
          for (int idx0 = ct & ~31; idx0 < total; idx0 += NCONV) {  // warp-uniform trip count
            const int idx = idx0 + lane;
            const bool live = idx < total;
            const int c8 = live ? idx / p.npos_in : 0, ip = live ? idx - c8 * p.npos_in : 0;
            const int tb = ip / per_img;
            const int rem = ip - tb * per_img;
            const int hr = rem / p.BW, wc = rem - hr * p.BW;
            const int w = wc - p.pad, h = h0 - p.pad + hr, b = b0 + tb;
            const bool inside = live && tb < p.TB && w >= 0 && w < p.W && h >= 0 && h < p.H && b < p.B;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j] = inside ? stg[((size_t)(tb * p.CC + c8 * 8 + j) * p.THH + hr) * p.W + w] : 0.f;
            uint4 hi;
            if (p.mode == 0) {
              // exact 3-way split x = hi + mid + lo (8 + 8 + 8 significand bits)
              float m[8], l[8];
              uint32_t nz = 0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float h1 = __bfloat162float(__float2bfloat16_rn(v[j]));
                const float r1 = v[j] - h1;
                const float m1 = __bfloat162float(__float2bfloat16_rn(r1));
                m[j] = m1; l[j] = r1 - m1;
                nz |= (r1 != 0.f);
              }
              any_low |= nz;
              hi = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              const uint4 mid = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
              const uint4 lo = make_uint4(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]), pack_bf16x2(l[4], l[5]), pack_bf16x2(l[6], l[7]));
              if (live) {
                *reinterpret_cast<uint4*>(opb + (size_t)p.op_term_bytes + (size_t)idx * 16) = mid;
                *reinterpret_cast<uint4*>(opb + (size_t)2 * p.op_term_bytes + (size_t)idx * 16) = lo;
              }
            } else {
              // fused fake-quant: integer level (exact in bf16), plus the saved codes / STE bits
              const bool owned = inside && hr >= p.pad && hr < p.pad + p.TH;
              float e[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                bool pass; float xq;
                const int code = mnb_act_quantize_one(q, v[j], pass, xq);
                e[j] = inside ? (float)(code + a_off) : 0.f;
                const int64_t fi = (((int64_t)b * p.C + cbase + c8 * 8 + j) * p.H + h) * p.W + w;
                if (p.codes && owned) p.codes[fi] = (uint8_t)code;
                if (p.pass_bits) {
                  // lanes that fall into the same 32-bit word combine their bits, one atomic per word
                  const uint32_t word = owned ? (uint32_t)(fi >> 5) : 0xffffffffu;
                  const uint32_t peers = __match_any_sync(0xffffffffu, word);
                  const uint32_t mine = (owned && pass) ? (1u << (fi & 31)) : 0u;
                  const uint32_t val = __reduce_or_sync(peers, mine);
                  if (owned && val && (__ffs(peers) - 1) == lane) atomicOr(p.pass_bits + word, val);
                }
              }
              hi = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
            }
            if (live) *reinterpret_cast<uint4*>(opb + (size_t)idx * 16) = hi;
          }
          if (p.mode == 0) {
            any_low = __reduce_or_sync(0xffffffffu, any_low);
            if (lane == 0) bar.op_flags[ob][cw] = any_low;
          }
          tc::fence_proxy_async_smem();
          tc::mbar_arrive(&bar.op_full[ob]);
          tc::mbar_arrive(&bar.stage_empty[st]);
        }
      }
    }
  }
done:
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols));
  }
}

}  // namespace tcconv

// Fused forward entry point.  x: fp32 NCHW.  qp == NULL: x is used as-is (raw fp32, exact 3-term
// bf16 split, e.g. wbwtab's +-1 activations); else x is fake-quantized on the fly and, when given,
// `codes` / `pass_bits` (zero-initialised by the caller) receive what the backward pass needs.
// Returns MNB_E_UNSUPPORTED when the geometry is outside what this kernel covers (the caller then
// uses the generic kernels): stride 1, dilation 1, "same" odd square filters, C/g % 16 == 0,
// K/g % 16 == 0 and <= 256, W in {4..64} with 128 % W == 0 or padded rows <= 128, resident weights.
extern "C" int mnb_fq_conv2d_fwd_tc(const mnb_conv_shape* s, const float* x, const mnb_act_qparams* qp,
                                    const int16_t* w_int, const float* w_scale, const float* bias, float* y,
                                    uint8_t* codes, uint32_t* pass_bits, int32_t* err_flag, mnb_stream_t stream) {
  using namespace tcconv;
  MNB_REQUIRE(s && x && w_int && w_scale && y && err_flag, "NULL pointer");
  FwdParams p{};
  p.B = s->batch; p.C = s->in_c; p.H = s->in_h; p.W = s->in_w; p.K = s->out_c; p.R = s->ker_h; p.S = s->ker_w;
  p.G = s->groups;
  MNB_REQUIRE(p.B > 0 && p.C > 0 && p.H > 0 && p.W > 0 && p.K > 0 && p.G > 0 && p.C % p.G == 0 && p.K % p.G == 0,
              "bad conv shape");
  auto unsupported = [](const char* why) { return mnb_fail(MNB_E_UNSUPPORTED, "tc conv: %s", why); };
  if (s->stride_h != 1 || s->stride_w != 1 || s->dil_h != 1 || s->dil_w != 1) return unsupported("stride/dilation != 1");
  if (p.R != p.S || (p.R & 1) == 0 || s->pad_h != p.R / 2 || s->pad_w != p.R / 2) return unsupported("not a 'same' odd square filter");
  p.pad = p.R / 2; p.Cg = p.C / p.G; p.Ng = p.K / p.G;
  if (p.Cg % 16 || p.Ng % 16 || p.Ng > 256) return unsupported("channels per group");
  if ((p.W * 4) % 16 || p.W > 64) return unsupported("row width");
  p.BW = p.W + 2 * p.pad;
  p.TH = std::min(p.H, 128 / p.BW);
  if (p.TH < 1) return unsupported("padded row wider than 128 positions");
  p.THH = p.TH + 2 * p.pad;
  p.TB = 1;
  if (p.pad == 0 && p.TH == p.H) p.TB = std::min(p.B, 128 / (p.H * p.W));  // whole small images per tile
  if (p.TB < 1) p.TB = 1;
  p.CC = (p.Cg % 32 == 0) ? 32 : 16;
  p.nchunk = p.Cg / p.CC;
  const int halo = (p.R - 1) * p.BW + (p.S - 1);
  const int npos = std::max(p.TB * p.THH * p.BW, 128) + halo;  // MMA rows read [tap_off, tap_off + 128)
  p.npos_in = (npos + 7) / 8 * 8;
  p.row_tiles = (p.H + p.TH - 1) / p.TH;
  p.n_tiles = ((p.B + p.TB - 1) / p.TB) * p.row_tiles;
  p.mode = qp ? qp->mode : 0;
  if (qp) {
    MNB_REQUIRE(qp->mode == MNB_ACT_DOREFA || qp->mode == MNB_ACT_IAO, "fused quantizer must be DoReFa or IAO");
    p.qp = *qp;
    p.a_offset = qp->mode == MNB_ACT_IAO ? qp->qmin : 0;
  }
  p.stage_bytes = p.W * p.THH * p.CC * p.TB * 4;
  p.op_term_bytes = p.npos_in * (p.CC / 8) * 16;
  p.op_buf_bytes = p.op_term_bytes * (p.mode == 0 ? 3 : 1);
  p.b_group_bytes = p.R * p.S * p.Cg * p.Ng * 2;
  const int fixed = NST * p.stage_bytes + NOP * p.op_buf_bytes + 2048;
  const int budget = kMaxDynSmem - 2048 - fixed;
  if (budget < p.b_group_bytes) return unsupported("weights of one group do not fit in shared memory");
  int max_groups = std::max(1, std::min(p.G, std::min(budget / p.b_group_bytes, std::max(1, 64 * 1024 / p.b_group_bytes))));
  while (p.G % max_groups) --max_groups;  // equal slabs: every CTA does the same work per tile
  p.slab_groups = max_groups;
  p.n_slabs = p.G / p.slab_groups;
  p.off_stage = 0;
  p.off_op = (NST * p.stage_bytes + 1023) / 1024 * 1024;
  p.off_b = p.off_op + (NOP * p.op_buf_bytes + 1023) / 1024 * 1024;
  const int smem_bytes = p.off_b + p.slab_groups * p.b_group_bytes;
  if (smem_bytes > kMaxDynSmem) return unsupported("shared memory budget");
  int cols = 32;
  while (cols < NACC * p.Ng) cols <<= 1;
  p.tmem_cols = cols;
  p.w_int = w_int; p.w_scale = w_scale; p.bias = bias; p.y = y; p.codes = codes; p.pass_bits = pass_bits; p.err = err_flag;
  p.a_scale = (qp && qp->mode == MNB_ACT_IAO) ? qp->scale : nullptr;
  p.a_scale_const = (qp && qp->mode == MNB_ACT_DOREFA) ? (float)(1.0 / (double)((1 << qp->bits) - 1)) : 1.f;
  CUtensorMap tmap;
  uint64_t dims[4] = {(uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.C, (uint64_t)p.B};
  uint32_t box[4] = {(uint32_t)p.W, (uint32_t)p.THH, (uint32_t)p.CC, (uint32_t)p.TB};
  if (int e = mnb_make_tmap(&tmap, x, 4, 4, dims, box)) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(fq_conv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (ce != cudaSuccess) return mnb_fail((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    attr_set = true;
  }
  const int64_t items = (int64_t)p.n_tiles * p.n_slabs;
  int grid = (int)std::min<int64_t>(items, MNB_NUM_SMS);
  grid = std::max(grid, p.n_slabs);
  fq_conv_fwd_tc_kernel<<<grid, NTHREADS, smem_bytes, (cudaStream_t)stream>>>(tmap, p);
  MNB_LAUNCHED(1);
  return 0;
}
