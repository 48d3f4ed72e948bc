// Weight gradient of the fake-quant convolution on tcgen05 tensor cores.
//
//   dWq[gNg+k, c, r, s] = s_a * sum_{b,h,w} dy[b, gNg+k, h, w] * e_a[b, gCg+c, h+r-p, w+s-p]
//
// The reduction runs over pixels, so both MMA operands are the position-major bf16 buffers the
// forward kernel builds (op[ch/8][position][8 ch]) - read here as MN-major UMMA operands:
//   A = dy^T  (128 gradient channels  x 16 positions per MMA), exact 3-term bf16 split of fp32 dy
//   B = e_a^T (N activation channels  x 16 positions per MMA), exact integer levels re-quantized
//                                                              from the fp32 input (or +-1 / bf16-exact raw x)
// and a filter tap (r, s) is again just a shifted start address of B.  Accumulators
// D[tap][128 x N] (fp32) live in TMEM for the whole kernel; every CTA owns a block of gradient
// channels, streams its share of the pixel tiles through TMA -> converter warps -> MMA, and writes
// one partial dW at the end; a second kernel reduces the partials in a fixed order (deterministic).
#include <cuda.h>
#include <cuda_bf16.h>

#include "mnb_common.cuh"
#include "mnb_tc.cuh"

namespace tcwgrad {

constexpr int NTHREADS = 512, NCONV = 448, NEPI = 128, MAXST = 8, BASEST = 4;  // converters: warps 2..15 (4..7 also run the final epilogue)
constexpr int KX = 4, KD = 3;  // operand entries per converter thread: activation / gradient chunks
constexpr int kMaxDynSmem = 227 * 1024 - 2048;

struct Params {
  int B, C, K, H, W, R, S, pad, G, Cg, Ng;
  int BW, TH, THH, TB, CC, nst;
  int npos_x, npos_d, ksteps, row_tiles, n_tiles;   // npos_d: gradient positions per tile (multiple of 16)
  int op_buf_bytes, nbuf;                            // one {activation, 3 x gradient} operand buffer; nbuf (1|2) are cycled
  int Gb, nsplit, n_block, n_slabs, ranks;
  int nseg_max;          // accumulation segments of the CTA with the most tiles (= partial slots per CTA)
  int seg_tiles;         // tiles per accumulation segment (tcgen05.mma truncates the running fp32 sum: bounded chains)
  int tap_groups, tpc;   // filter taps are split over `tap_groups` CTA sets of `tpc` taps (TMEM holds tpc * n_block columns)
  int quant_mode, a_offset, tmem_cols;
  int slot_bytes, stage_x_bytes, stage_d_bytes, xop_bytes, dop_term_bytes, off_stage, off_xop, off_dop;
  mnb_act_qparams qp;
  float* partial;     // [ranks][K * Cg * RS]
  int* err;
  int* inexact;       // set when a raw fp32 activation is not bf16-exact (result then invalid)
};

struct alignas(16) Shared {
  uint64_t stage_full[MAXST], stage_empty[MAXST], op_full[2], op_empty[2], done;
  uint32_t tmem_slot;
  uint32_t abort;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// (a, b) -> packed bf16 pairs of the exact pieces hi, mid, lo with a = hi_a + mid_a + lo_a (same for b)
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t& hp, uint32_t& mp, uint32_t& lp) {
  hp = pack_bf16x2(a, b);
  const float ra = a - __uint_as_float(hp << 16), rb = b - __uint_as_float(hp & 0xffff0000u);
  mp = pack_bf16x2(ra, rb);
  const float la = ra - __uint_as_float(mp << 16), lb = rb - __uint_as_float(mp & 0xffff0000u);
  lp = pack_bf16x2(la, lb);
}

__global__ void __launch_bounds__(NTHREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Shared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* stage_base = smem + p.off_stage;
  uint8_t* xop = smem + p.off_xop;
  uint8_t* dop = smem + p.off_dop;
  const int RS = p.R * p.S;

  // ---- block of gradient channels owned by this CTA
  const int slab = blockIdx.x % p.n_slabs, rank = blockIdx.x / p.n_slabs;
  const int tg = slab % p.tap_groups, cb = slab / p.tap_groups;   // tap group, channel block
  const int gbi = cb / p.nsplit, ms = cb - gbi * p.nsplit;
  const int tap0 = tg * p.tpc, tap1 = min(RS, tap0 + p.tpc);
  const int dy_ch0 = gbi * p.Gb * p.Ng + ms * 128;
  const int m_real = min(128, p.Gb * p.Ng - ms * 128);
  const int x_ch0 = gbi * p.Gb * p.Cg;
  const int x_chunks = p.n_block / p.CC, d_chunks = m_real / p.CC;
  const int my_tiles = (p.n_tiles > rank) ? (p.n_tiles - rank + p.ranks - 1) / p.ranks : 0;

  if (tid == 0) {
    for (int i = 0; i < p.nst; ++i) { tc::mbar_init(&sh.stage_full[i], 1); tc::mbar_init(&sh.stage_empty[i], NCONV / 32); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&sh.op_full[i], NCONV / 32); tc::mbar_init(&sh.op_empty[i], 1); }
    tc::mbar_init(&sh.done, 1);
    sh.abort = 0;
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmap_x);
    tc::prefetch_tmap(&tmap_dy);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&sh.tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // operand buffers start as zeros: padded columns / rows and unused channel rows are never written
  for (int i = tid; i < p.nbuf * p.op_buf_bytes / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(xop)[i] = make_uint4(0, 0, 0, 0);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sh.tmem_slot;

  // TMEM -> the partial dW of (segment, CTA) (warps 4..7; plain stores: every accumulation segment has its own slot and the
  // reduce kernel adds them in a fixed order with round-to-nearest adds - the reason the K loop is cut into segments,
  // see Params::seg_tiles)
  auto drain = [&](bool have, int seg) {
    const int q = warp - 4, m = q * 32 + lane;
    const int64_t wsize = (int64_t)p.K * p.Cg * RS;
    float* mine = p.partial + ((int64_t)seg * p.ranks + rank) * wsize;
    const int k_abs = dy_ch0 + m;
    const int k_group = k_abs / p.Ng;
    for (int tap = tap0; tap < tap1; ++tap) {
      for (int n0 = 0; n0 < p.n_block; n0 += 32) {
        uint32_t r[32];
        if (have) {
          tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((tap - tap0) * p.n_block + n0), r);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (m < m_real) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + j;
            if (n < p.n_block) {
              const int xg = (x_ch0 + n) / p.Cg;           // group of this activation channel
              if (xg == k_group) {
                const int c = (x_ch0 + n) - xg * p.Cg;
                mine[((int64_t)k_abs * p.Cg + c) * RS + tap] = __uint_as_float(r[j]);
              }
            }
          }
        }
      }
    }
    tc::tc_fence_before();
  };

  if (warp == 0) {
    // ================================================================= TMA producer
    if (lane == 0) {
      int st = -1;          // ring position advanced incrementally (no per-chunk integer division)
      uint32_t ph = 1;
      for (int tile = rank; tile < p.n_tiles; tile += p.ranks) {
        const int bt = tile / p.row_tiles, rt = tile - bt * p.row_tiles;
        const int b0 = bt * p.TB, h0 = rt * p.TH;
        for (int ch = 0; ch < x_chunks + d_chunks; ++ch) {
          if (++st == p.nst) st = 0;
          ph ^= (st == 0);
          if (!tc::mbar_wait(&sh.stage_empty[st], ph ^ 1, p.err, 401)) goto done;
          uint8_t* dst = stage_base + (size_t)st * p.slot_bytes;
          if (ch < x_chunks) {
            tc::mbar_arrive_expect_tx(&sh.stage_full[st], (uint32_t)p.stage_x_bytes);
            if (p.pad == 0) tc::tma_load_3d(dst, &tmap_x, &sh.stage_full[st], h0 * p.W, x_ch0 + ch * p.CC, b0);
            else tc::tma_load_4d(dst, &tmap_x, &sh.stage_full[st], 0, h0 - p.pad, x_ch0 + ch * p.CC, b0);
          } else {
            tc::mbar_arrive_expect_tx(&sh.stage_full[st], (uint32_t)p.stage_d_bytes);
            tc::tma_load_3d(dst, &tmap_dy, &sh.stage_full[st], h0 * p.W, dy_ch0 + (ch - x_chunks) * p.CC, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (warp-converged, lane 0 issues)
    {
      const uint32_t lead = lane == 0;
      const uint32_t idesc = tc::make_idesc_major(1, 1, 1, 128, (uint32_t)p.n_block, 1, 1);
      // descriptors differ only in the 14-bit start-address field: build once, then add (bytes >> 4)
      const uint64_t b_desc0 = tc::smem_desc_mnmajor_noswz(tc::smem_u32(xop), 128, (uint32_t)p.npos_x * 16u);
      const uint64_t a_desc0 = tc::smem_desc_mnmajor_noswz(tc::smem_u32(dop), 128, (uint32_t)p.npos_d * 16u);
      const uint32_t buf16 = (uint32_t)p.op_buf_bytes >> 4;
      const uint32_t a_term = (uint32_t)p.dop_term_bytes >> 4;
      uint32_t t = 0, tseg = 0;
      for (int tile = rank; tile < p.n_tiles; tile += p.ranks, ++t) {
        const uint32_t ob = p.nbuf == 2 ? (t & 1u) : 0u, oph = p.nbuf == 2 ? ((t >> 1) & 1u) : (t & 1u);
        tc::mbar_wait_soft(&sh.op_full[ob], oph, p.err, 402, &sh.abort);
        tc::tc_fence_after();
        uint32_t d_col = 0;
        for (int tap = tap0; tap < tap1; ++tap, d_col += (uint32_t)p.n_block) {
          const int r = tap / p.S, s2 = tap - r * p.S;
          const uint64_t b_tap = b_desc0 + (uint64_t)(ob * buf16 + (uint32_t)(r * p.BW + s2));
          const uint32_t d_tmem = tmem + d_col;
          for (int ps = 0; ps < p.ksteps; ++ps) {
            const uint64_t bd = b_tap + (uint64_t)(ps * 16);   // 16 positions x 16 bytes = 256 B
            const uint64_t ad = a_desc0 + (uint64_t)(ob * buf16 + (uint32_t)(ps * 16));
            tc::mma_f16_guarded(d_tmem, ad, bd, idesc, (tseg | (uint32_t)ps) != 0, lead);
            tc::mma_f16_guarded(d_tmem, ad + a_term, bd, idesc, 1, lead);
            tc::mma_f16_guarded(d_tmem, ad + 2 * a_term, bd, idesc, 1, lead);
          }
        }
        if (lead) tc::mma_commit(&sh.op_empty[ob]);
        __syncwarp();
        if (++tseg == (uint32_t)p.seg_tiles && (int)(t + 1) < my_tiles) {
          // segment complete: warps 4..7 drain the accumulators into the partial before they convert the next tile
          // (their op_full arrival for that tile orders the drain before the overwriting MMA below)
          if (lead) tc::mma_commit(&sh.done);
          __syncwarp();
          tseg = 0;
        }
      }
      if (lead) tc::mma_commit(&sh.done);
      __syncwarp();
    }
  } else {
    // ================================================================= converters (warps 2..15)
    const int ct = tid - 64;
    MnbActQ q;
    if (p.quant_mode != 0) q = mnb_load_actq(p.qp);
    const int a_off = p.a_offset + ((p.quant_mode == MNB_ACT_IAO && p.qp.zero_point) ? (int)__ldg(p.qp.zero_point) : 0);
    const int per_img = p.THH * p.BW, c8s = p.CC / 8;
    const int total_x = p.npos_x * c8s;
    const int nvalid = p.TB * p.TH * p.W;
    const int total_d = nvalid * c8s;
    int xso[KX], xmeta[KX];
#pragma unroll
    for (int k = 0; k < KX; ++k) {
      const int idx = ct + k * NCONV;
      xso[k] = -1; xmeta[k] = 0;
      if (idx < total_x) {
        const int c8 = idx / p.npos_x, ip = idx - c8 * p.npos_x;
        const int tb = ip / per_img;
        const int rem = ip - tb * per_img;
        const int hr = rem / p.BW, wc = rem - hr * p.BW;
        const int w = wc - p.pad;
        if (tb < p.TB && w >= 0 && w < p.W) {
          xso[k] = ((tb * p.CC + c8 * 8) * p.THH + hr) * p.W + w;
          xmeta[k] = hr | (tb << 8);
        }
      }
    }
    int dso[KD], ddst[KD];
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int idx = ct + k * NCONV;
      dso[k] = -1; ddst[k] = 0;
      if (idx < total_d) {
        const int c8 = idx / nvalid, vp = idx - c8 * nvalid;
        const int tb = vp / (p.TH * p.W);
        const int rem = vp - tb * (p.TH * p.W);
        const int th = rem / p.W, w = rem - th * p.W;
        dso[k] = ((tb * p.CC + c8 * 8) * p.TH + th) * p.W + w;
        ddst[k] = (c8 * p.npos_d + (tb * p.THH + th) * p.BW + w) * 16;  // byte offset inside a chunk's 8-ch groups
      }
    }
    const int x_chstride = p.THH * p.W, d_chstride = p.TH * p.W;
    uint32_t t = 0;
    int st = -1;
    uint32_t ph = 1;
    for (int tile = rank; tile < p.n_tiles; tile += p.ranks, ++t) {
      const int bt = tile / p.row_tiles, rt = tile - bt * p.row_tiles;
      const int b0 = bt * p.TB, h0 = rt * p.TH;
      const uint32_t ob = p.nbuf == 2 ? (t & 1u) : 0u, oph = p.nbuf == 2 ? ((t >> 1) & 1u) : (t & 1u);
      if (warp < 8 && warp >= 4 && t > 0 && t % (uint32_t)p.seg_tiles == 0) {
        const uint32_t seg = t / (uint32_t)p.seg_tiles - 1;
        if (!tc::mbar_wait(&sh.done, seg & 1u, p.err, 406)) goto done;
        tc::tc_fence_after();
        drain(true, (int)seg);
      }
      if (!tc::mbar_wait(&sh.op_empty[ob], oph ^ 1, p.err, 404)) goto done;  // MMAs of the tile two back retired
      uint8_t* xop_b = xop + (size_t)ob * p.op_buf_bytes;
      uint8_t* dop_b = dop + (size_t)ob * p.op_buf_bytes;
      for (int ch = 0; ch < x_chunks + d_chunks; ++ch) {
        if (++st == p.nst) st = 0;
        ph ^= (st == 0);
        if (!tc::mbar_wait(&sh.stage_full[st], ph, p.err, 405)) goto done;
        const float* stg = reinterpret_cast<const float*>(stage_base + (size_t)st * p.slot_bytes);
        if (ch < x_chunks) {
          uint8_t* dstb = xop_b + (size_t)ch * c8s * p.npos_x * 16;
#pragma unroll
          for (int k = 0; k < KX; ++k) {
            const int idx = ct + k * NCONV;
            if (idx >= total_x) break;
            const int so = xso[k];
            if (so < 0) continue;  // halo column / dead position: stays zero
            uint32_t u[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) u[j] = __float_as_uint(stg[so + j * x_chstride]);
            uint4 v;
            if (p.quant_mode == 0) {
              uint32_t low = 0;
#pragma unroll
              for (int j = 0; j < 8; ++j) low |= u[j] & 0xffffu;
              if (low) atomicOr(p.inexact, 1);
              v = make_uint4(__byte_perm(u[0], u[1], 0x7632), __byte_perm(u[2], u[3], 0x7632),
                             __byte_perm(u[4], u[5], 0x7632), __byte_perm(u[6], u[7], 0x7632));
            } else {
              const int hr = xmeta[k] & 255, tb = xmeta[k] >> 8;
              const int h = h0 - p.pad + hr;
              const bool inside = h >= 0 && h < p.H && (b0 + tb) < p.B;
              float e[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                bool pass;
                const int code = mnb_act_code_certified(q, __uint_as_float(u[j]), pass);
                e[j] = inside ? (float)(code + a_off) : 0.f;
              }
              v = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
            }
            *reinterpret_cast<uint4*>(dstb + (size_t)idx * 16) = v;
          }
        } else {
          const int dch = ch - x_chunks;
          uint8_t* dstb = dop_b + (size_t)dch * c8s * p.npos_d * 16;
#pragma unroll
          for (int k = 0; k < KD; ++k) {
            if (ct + k * NCONV >= total_d) break;
            const int so = dso[k];
            uint32_t hp[4], mp[4], lp[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              split3_pair(stg[so + (2 * j) * d_chstride], stg[so + (2 * j + 1) * d_chstride], hp[j], mp[j], lp[j]);
            uint8_t* d0 = dstb + ddst[k];
            *reinterpret_cast<uint4*>(d0) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            *reinterpret_cast<uint4*>(d0 + p.dop_term_bytes) = make_uint4(mp[0], mp[1], mp[2], mp[3]);
            *reinterpret_cast<uint4*>(d0 + 2 * p.dop_term_bytes) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&sh.stage_empty[st]);
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&sh.op_full[ob]);
    }
  }
  if (warp >= 4 && warp < 8) {
    // ================================================================= final epilogue: last segment -> partial dW
    const bool have = my_tiles > 0;
    const uint32_t last_seg = have ? (uint32_t)((my_tiles - 1) / p.seg_tiles) : 0u;
    if (have) {
      if (!tc::mbar_wait(&sh.done, last_seg & 1u, p.err, 403)) goto done;
      tc::tc_fence_after();
    }
    drain(have, (int)last_seg);
    for (int seg = have ? (int)last_seg + 1 : 1; seg < p.nseg_max; ++seg) drain(false, seg);   // ranks with fewer tiles: zeros
  }
done:
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols));
  }
}

__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float* __restrict__ partial, int64_t n, int ranks,
                                                              const float* a_scale, float a_scale_const,
                                                              float* __restrict__ dwq) {
  const float sc = a_scale ? __ldg(a_scale) : a_scale_const;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four chains in flight, fixed order: deterministic
    int j = 0;
    for (; j + 4 <= ranks; j += 4) {
      s0 += partial[(int64_t)j * n + i];
      s1 += partial[(int64_t)(j + 1) * n + i];
      s2 += partial[(int64_t)(j + 2) * n + i];
      s3 += partial[(int64_t)(j + 3) * n + i];
    }
    for (; j < ranks; ++j) s0 += partial[(int64_t)j * n + i];
    dwq[i] = __fmul_rn((s0 + s1) + (s2 + s3), sc);
  }
}

static int plan(const mnb_conv_shape* s, int quant_mode, Params& p, int& smem_bytes) {
  MNB_REQUIRE(s != nullptr, "conv shape is NULL");
  auto unsupported = [](const char* why) { return mnb_fail(MNB_E_UNSUPPORTED, "tc wgrad: %s", why); };
  p.B = s->batch; p.C = s->in_c; p.K = s->out_c; p.H = s->in_h; p.W = s->in_w; p.R = s->ker_h; p.S = s->ker_w; p.G = s->groups;
  MNB_REQUIRE(p.B > 0 && p.C > 0 && p.K > 0 && p.H > 0 && p.W > 0 && p.G > 0 && p.C % p.G == 0 && p.K % p.G == 0, "bad conv shape");
  if (s->stride_h != 1 || s->stride_w != 1 || s->dil_h != 1 || s->dil_w != 1) return unsupported("stride/dilation != 1");
  if (p.R != p.S || (p.R & 1) == 0 || s->pad_h != p.R / 2 || s->pad_w != p.R / 2) return unsupported("not a 'same' odd square filter");
  p.pad = p.R / 2; p.Cg = p.C / p.G; p.Ng = p.K / p.G;
  const int RS = p.R * p.S;
  if (p.Cg % 16 || p.Ng % 16) return unsupported("channels per group");
  if ((p.W * 4) % 16 || p.W > 64 || p.H > 255) return unsupported("image size");
  // block of gradient channels: Gb whole groups (Gb*Ng <= 128) or a 128-slice of one big group.  An MMA
  // costs ~61 cycles whatever its N (<= 128), so wide blocks are cheaper per channel even if their taps no
  // longer fit TMEM at once and must be split over `tap_groups` CTA sets (each re-reading the inputs).
  p.Gb = 1; p.nsplit = 1; p.tap_groups = 1; p.tpc = RS;
  if (p.Ng >= 128) {
    p.nsplit = (p.Ng + 127) / 128;
    if (p.Cg > 256) return unsupported("more than 256 activation channels per group");
    p.tpc = std::min(RS, 512 / p.Cg);
    if (p.tpc < 1) return unsupported("accumulators exceed tensor memory");
    p.tap_groups = (RS + p.tpc - 1) / p.tpc;
    p.tpc = (RS + p.tap_groups - 1) / p.tap_groups;
  } else {
    double best = 1e30;
    for (int gb = 1; gb <= p.G; ++gb) {
      if (p.G % gb || gb * p.Ng > 128 || gb * p.Cg > 256) continue;
      int tpc = std::min(RS, 512 / (gb * p.Cg));
      if (tpc < 1) continue;
      const int tgs = (RS + tpc - 1) / tpc;
      tpc = (RS + tgs - 1) / tgs;
      const double mma = tpc * 24.0 * 61.0;
      const double bytes = 4.0 * 128.0 * gb * (p.Cg * 1.3 + p.Ng);   // rough bytes per tile (activation halo ~1.3x)
      const double cost = tgs * std::max(mma, bytes / 23.0) / gb;
      if (cost < best) { best = cost; p.Gb = gb; p.tpc = tpc; p.tap_groups = tgs; }
    }
  }
  p.n_block = p.Gb * p.Cg;
  if (p.n_block > 256 || p.tpc * p.n_block > 512) return unsupported("accumulators exceed tensor memory");
  int cc0 = 32;
  if (p.n_block % 32 || (p.Gb * p.Ng) % 32 || (p.nsplit > 1 && p.Ng % 32)) cc0 = 16;
  if (p.nsplit > 1 && (p.Ng % 128) % cc0) return unsupported("ragged channel split");
  p.n_slabs = (p.G / p.Gb) * p.nsplit * p.tap_groups;
  p.BW = p.W + 2 * p.pad;
  const int th_max = std::min(p.H, 128 / p.BW);
  if (th_max < 1) return unsupported("padded row wider than 128 positions");
  const int halo = (p.R - 1) * p.BW + (p.S - 1);
  // rows per tile: the largest that lets TWO operand buffers (converter and MMA overlap) plus a staging
  // ring of >= 3 slots fit in shared memory
  bool ok = false;
  auto try_tile = [&](int TH, int TB, int nbuf) -> bool {
    p.TH = TH; p.TB = TB; p.nbuf = nbuf;
    p.THH = p.TH + 2 * p.pad;
    p.npos_d = ((p.pad > 0 ? p.TH * p.BW : p.TB * p.TH * p.W) + 15) / 16 * 16;
    if (p.npos_d > 128) return false;
    p.ksteps = p.npos_d / 16;
    p.npos_x = (p.npos_d + halo + 7) / 8 * 8;
    p.CC = cc0;
    if (p.npos_x * (p.CC / 8) > KX * NCONV || p.TB * p.TH * p.W * (p.CC / 8) > KD * NCONV) {
      if (p.CC == 32 && p.npos_x * 2 <= KX * NCONV) p.CC = 16; else return false;
    }
    p.stage_x_bytes = p.W * p.THH * p.CC * p.TB * 4;
    p.stage_d_bytes = p.W * p.TH * p.CC * p.TB * 4;
    p.slot_bytes = (std::max(p.stage_x_bytes, p.stage_d_bytes) + 127) / 128 * 128;
    p.xop_bytes = (p.n_block / 8) * p.npos_x * 16;
    p.dop_term_bytes = 16 * p.npos_d * 16;  // 128 channel rows (16 groups of 8) x npos_d positions
    p.op_buf_bytes = (p.xop_bytes + 3 * p.dop_term_bytes + 1023) / 1024 * 1024;
    p.off_xop = 0;
    p.off_dop = p.xop_bytes;
    p.off_stage = p.nbuf * p.op_buf_bytes;
    p.nst = BASEST;
    while (p.nst > 3 && p.off_stage + p.nst * p.slot_bytes > kMaxDynSmem) --p.nst;
    if (p.off_stage + p.nst * p.slot_bytes > kMaxDynSmem) return false;
    while (p.nst < MAXST && p.off_stage + (p.nst + 1) * p.slot_bytes <= kMaxDynSmem) ++p.nst;  // more bytes in flight
    smem_bytes = p.off_stage + p.nst * p.slot_bytes;
    return true;
  };
  // 1x1 filters (no halo re-reads, few MMAs per tile): one 128-position operand set, converter and MMA
  // alternate.  Filters with taps: two smaller operand sets so that the (tap-heavy) MMAs overlap the converter.
  if (p.pad == 0) {
    const int TH = th_max;
    const int tb = (TH == p.H) ? std::max(1, std::min(p.B, 128 / (p.H * p.W))) : 1;
    ok = try_tile(TH, tb, 1);
  }
  if (!ok) {
    int best_th = 0, best_tb = 0, best_tiles = 1 << 30;
    for (int TH = th_max; TH >= 1; --TH) {
      const int tb_max = (p.pad == 0 && TH == p.H) ? std::max(1, std::min(p.B, 128 / (p.H * p.W))) : 1;
      for (int TB = tb_max; TB >= 1; TB >>= 1) {
        if (!try_tile(TH, TB, 2)) continue;
        const int tiles = ((p.B + TB - 1) / TB) * ((p.H + TH - 1) / TH);
        if (tiles <= best_tiles) { best_tiles = tiles; best_th = TH; best_tb = TB; }  // ties: smaller tile
        break;
      }
    }
    ok = best_th > 0 && try_tile(best_th, best_tb, 2);
  }
  if (!ok) return unsupported("shared memory budget");
  p.row_tiles = (p.H + p.TH - 1) / p.TH;
  p.n_tiles = ((p.B + p.TB - 1) / p.TB) * p.row_tiles;
  p.ranks = std::max(1, std::min(p.n_tiles, MNB_NUM_SMS / p.n_slabs));  // one wave of CTAs
  // chains of at most ~256 MMAs per accumulator (measured truncation bias of tcgen05.mma: 1.8e-8 |D| per instruction)
  int chain = 256;
  if (const char* e = getenv("MNB_WG_CHAIN")) chain = std::max(1, atoi(e));
  p.seg_tiles = std::max(1, chain / (3 * p.ksteps));
  p.nseg_max = mnb_ceil_div(mnb_ceil_div(p.n_tiles, p.ranks), p.seg_tiles);
  p.quant_mode = quant_mode;
  int cols = 32;
  while (cols < p.tpc * p.n_block) cols <<= 1;
  p.tmem_cols = cols;
  return 0;
}

}  // namespace tcwgrad

extern "C" int64_t mnb_wgrad_tc_scratch_bytes(const mnb_conv_shape* s) {
  tcwgrad::Params p{};
  int smem = 0;
  if (tcwgrad::plan(s, 0, p, smem)) return -1;
  return (int64_t)p.nseg_max * p.ranks * p.K * p.Cg * p.R * p.S * 4;
}

extern "C" int mnb_conv2d_wgrad_tc(const mnb_conv_shape* s, const float* dy, const float* x, const mnb_act_qparams* qp,
                                   float* dwq, void* scratch, int32_t* inexact_flag, int32_t* err_flag,
                                   mnb_stream_t stream) {
  using namespace tcwgrad;
  MNB_REQUIRE(s && dy && x && dwq && scratch && inexact_flag && err_flag, "NULL pointer");
  if (qp) MNB_REQUIRE(qp->mode == MNB_ACT_DOREFA || qp->mode == MNB_ACT_IAO, "fused quantizer must be DoReFa or IAO");
  Params p{};
  int smem_bytes = 0;
  if (int e = plan(s, qp ? qp->mode : 0, p, smem_bytes)) return e;
  if (qp) {
    if (qp->mode == MNB_ACT_DOREFA) MNB_REQUIRE(qp->bits >= 2 && qp->bits <= 8, "DoReFa a_bits must be in [2,8]");
    p.qp = *qp;
    p.a_offset = qp->mode == MNB_ACT_IAO ? qp->qmin : 0;
  }
  p.partial = reinterpret_cast<float*>(scratch); p.err = err_flag; p.inexact = inexact_flag;
  CUtensorMap tx, td;
  // TMA cost is per box row: tiles without a halo (dy always, x of a 1x1 filter) are read as ONE contiguous row of
  // TH*W floats per channel over a collapsed (H*W, C, B) view; rows past the image end are zero-filled as before.
  uint64_t dd[3] = {(uint64_t)p.H * p.W, (uint64_t)p.K, (uint64_t)p.B};
  uint32_t bd[3] = {(uint32_t)(p.TH * p.W), (uint32_t)p.CC, (uint32_t)p.TB};
  if (int e = mnb_make_tmap(&td, dy, 4, 3, dd, bd)) return e;
  if (p.pad == 0) {
    uint64_t dx[3] = {(uint64_t)p.H * p.W, (uint64_t)p.C, (uint64_t)p.B};
    uint32_t bx[3] = {(uint32_t)(p.TH * p.W), (uint32_t)p.CC, (uint32_t)p.TB};
    if (int e = mnb_make_tmap(&tx, x, 4, 3, dx, bx)) return e;
  } else {
    uint64_t dx[4] = {(uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.C, (uint64_t)p.B};
    uint32_t bx[4] = {(uint32_t)p.W, (uint32_t)p.THH, (uint32_t)p.CC, (uint32_t)p.TB};
    if (int e = mnb_make_tmap(&tx, x, 4, 4, dx, bx)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (ce != cudaSuccess) return mnb_fail((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    attr_set = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  wgrad_tc_kernel<<<p.n_slabs * p.ranks, NTHREADS, smem_bytes, st>>>(tx, td, p);
  const int64_t n = (int64_t)p.K * p.Cg * p.R * p.S;
  const float a_const = (qp && qp->mode == MNB_ACT_DOREFA) ? (float)(1.0 / (double)((1 << qp->bits) - 1)) : 1.f;
  const float* a_ptr = (qp && qp->mode == MNB_ACT_IAO) ? qp->scale : nullptr;
  int blocks = (int)std::min<int64_t>(mnb_ceil_div(n, 256), MNB_NUM_SMS * 8);
  wgrad_tc_reduce_kernel<<<blocks, 256, 0, st>>>(p.partial, n, p.ranks * p.nseg_max, a_ptr, a_const, dwq);
  MNB_LAUNCHED(2);
  return 0;
}
