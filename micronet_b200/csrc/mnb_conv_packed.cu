// EXPERIMENTAL (round-2 groundwork, off unless MNB_PACKED_OPERANDS=1; not yet validated on hardware): packed bf16
// activations between the BatchNorm + binarizer producer and the tensor-core forward convolution.
//
//   producer : mnb_bn_sign_fwd_packed  = mnb_bn_sign_fwd  +  xp[b][c/8][h][w][8] (bf16 +-1, OUTPUT channel order)
//   consumer : mnb_fq_conv2d_fwd_packed_tc: y = conv2d(xp, w_int * w_scale) + bias, same tiles / descriptors / epilogue as
//              mnb_conv_tc_fwd.cu, but the A operand is not converted at all: one 5-D TMA box per tile,
//                dims (8 ch, W, H, B, C/8)  [the channel-octet dimension declared last],  box (8, W+2p, TH+2p, TB, slabC/8),
//              lands in shared memory as [C/8][TB][rows][W+2p][8] = the position-major UMMA operand
//              op[c/8][position][8] of DESIGN.md 2, halo rows AND columns zero-filled by the TMA unit (W is not
//              the innermost dimension here; the innermost start coordinate stays 0).  One hand-off per tile and slab
//              instead of one per 32-channel chunk, no staging ring, no converter warps, 2 B instead of 4 B per element.
//
// Why: DESIGN.md 6 - the 1x1 kernels' floor is the per-chunk converter hand-off, not HBM, MMA or TMA throughput.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "mnb_common.cuh"
#include "mnb_tc.cuh"

namespace tcpacked {

constexpr int NTHREADS = 256;   // warp 0: TMA, warp 1: MMA issue, warp 2: TMEM alloc, warps 4..7: epilogue
constexpr int NACC = 2, MAXBUF = 4;
constexpr int kMaxDynSmem = 227 * 1024 - 4096;

struct Params {
  int B, Cin, Cout, H, W, R, pad, G, cin_g, cout_g;
  int BW, TH, THH, TB, npos;          // npos = TB * THH * BW positions per box and channel octet
  int row_tiles, n_tiles, slab_groups, n_slabs, nbuf, nbuf_log2;
  int op_bytes, op_box_bytes, b_group_bytes, off_b, tmem_cols;
  // Private copy of everything the MMA-issue warp needs, precomputed on the host.  Goal: descriptor arithmetic in
  // UNIFORM registers - no ELECT + 4-5 R2UR.BROADCAST in front of every tcgen05.mma (this kernel: 0 of them, ~18
  // uniform-datapath instructions per MMA).  Rules established by SASS inspection of this kernel's variants
  // (cuobjdump -sass, count R2UR.BROADCAST per UTCHMMA; no GPU needed):
  //   (1) no integer division / modulo in the address or loop-bound chain (2-D grid instead of blockIdx.x / n_slabs,
  //       nested r / s loops instead of tap / R);
  //   (2) ring index = counter & mask, never a loop-carried conditional reset (if (++i == n) i = 0);
  //   (3) no barrier wait in front of the tile loop (wait for the weights inside it);
  //   (4) no input shared with per-thread code of other roles - common sub-expressions are computed once, in vector
  //       registers (hence this block) - and that includes blockIdx.x / gridDim.x: a copy made at kernel scope
  //       (const int rank = blockIdx.x;) and used by other roles' tile loops was the last thing that kept this kernel on
  //       the broadcast form.  Every role reads the special registers itself.
  struct Mma {
    uint32_t idesc, a_lbo, b_lbo, a_buf16, a_group16, a_kstep16, b_kstep16, b_tap16, b_group16;
    uint32_t ksteps, R, BW, slab_groups, slab_cols, cout_g, buf_mask, buf_log2, off_b, n_tiles;
  } m;
  const uint8_t* w_pack;    // bf16 [g][tap][k/8][n][8]
  const float* w_scale; const float* bias;
  float* out;
  int* err;
};

struct alignas(16) Shared {
  uint64_t op_full[MAXBUF], op_empty[MAXBUF], acc_full[NACC], acc_empty[NACC], b_full;
  uint32_t tmem_slot;
  uint32_t abort;
  alignas(16) float epi_scale[256];
  alignas(16) float epi_bias[256];
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const int16_t* __restrict__ w_int, __nv_bfloat16* __restrict__ out,
                                                           int G, int cin_g, int cout_g, int RS) {
  const int per_group = RS * cin_g * cout_g, total = G * per_group;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int g = idx / per_group;       // destination [g][tap][c8][n][e]
    int r = idx - g * per_group;
    const int tap = r / (cin_g * cout_g);
    r -= tap * (cin_g * cout_g);
    const int c8 = r / (cout_g * 8);
    r -= c8 * (cout_g * 8);
    const int n = r >> 3, c = c8 * 8 + (r & 7);
    out[idx] = __float2bfloat16_rn((float)__ldg(w_int + ((int64_t)(g * cout_g + n) * cin_g + c) * RS + tap));
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) fwd_packed_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ Shared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* b_base = smem + p.off_b;
  const int RS = p.R * p.R, c8_per_group = p.cin_g / 8;
  // grid = (CTAs per slab, slabs): the tile loops then run on blockIdx / gridDim only.  An integer division in the
  // work assignment (blockIdx.x / n_slabs) makes ptxas treat the tile loop's trip count as possibly divergent, and
  // everything inside it - the MMA issue included - leaves the uniform datapath.
  // NOTE: every role reads blockIdx.x / gridDim.x ITSELF inside its own branch.  Copies made here, at kernel scope,
  // are materialised in vector registers and shared by all roles, and the MMA warp's tile loop then leaves the uniform
  // datapath (41 instead of 0 R2UR.BROADCAST in this kernel's SASS).
  const int slab = blockIdx.y;
  const int g_first = slab * p.slab_groups;
  const int slab_cols = p.slab_groups * p.cout_g;

  if (tid == 0) {
    for (int i = 0; i < MAXBUF; ++i) { tc::mbar_init(&sh.op_full[i], 1); tc::mbar_init(&sh.op_empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { tc::mbar_init(&sh.acc_full[i], 1); tc::mbar_init(&sh.acc_empty[i], 128); }
    tc::mbar_init(&sh.b_full, 1);
    sh.abort = 0;
    tc::fence_barrier_init();
    tc::prefetch_tmap(&tmap);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&sh.tmem_slot)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int n = tid; n < slab_cols; n += NTHREADS) {
    sh.epi_scale[n] = __ldg(p.w_scale + g_first * p.cout_g + n);
    sh.epi_bias[n] = p.bias ? __ldg(p.bias + g_first * p.cout_g + n) : 0.f;
  }
  // slack rows behind every operand buffer (read by the MMAs of invalid halo positions only) must at least be finite
  for (int i = tid; i < p.nbuf * p.op_bytes / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = sh.tmem_slot;

  if (warp == 0) {
    // ================================================================= TMA producer: weights once, one box per tile
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(p.slab_groups * p.b_group_bytes);
      tc::mbar_arrive_expect_tx(&sh.b_full, bytes);
      tc::bulk_load_1d(b_base, p.w_pack + (size_t)g_first * p.b_group_bytes, bytes, &sh.b_full);
      uint32_t item = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++item) {
        const int bt = tile / p.row_tiles, rt = tile - bt * p.row_tiles;
        const uint32_t buf = item & (uint32_t)(p.nbuf - 1), ph = (item >> p.nbuf_log2) & 1u;
        if (!tc::mbar_wait(&sh.op_empty[buf], ph ^ 1, p.err, 601)) break;
        tc::mbar_arrive_expect_tx(&sh.op_full[buf], (uint32_t)p.op_box_bytes);
        tc::tma_load_5d(smem + (size_t)buf * p.op_bytes, &tmap, &sh.op_full[buf], 0, -p.pad, rt * p.TH - p.pad, bt * p.TB,
                        g_first * c8_per_group);
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer (warp-converged, lane 0 issues)
    // uses p.m.* (private parameters), blockIdx / gridDim and its own counters only - see Params::Mma
    const uint32_t lead = lane == 0;
    const uint32_t tmem_m = tmem;
    const uint64_t a_desc0 = tc::smem_desc_kmajor_noswz(tc::smem_u32(smem), p.m.a_lbo, 128);
    const uint64_t b_desc0 = tc::smem_desc_kmajor_noswz(tc::smem_u32(smem) + p.m.off_b, p.m.b_lbo, 128);
    uint32_t item = 0;
    for (uint32_t tile = blockIdx.x; tile < p.m.n_tiles; tile += gridDim.x, ++item) {
      const uint32_t acc = item & 1u, aph = (item >> 1) & 1u;
      const uint32_t buf = item & p.m.buf_mask, ph = (item >> p.m.buf_log2) & 1u;   // ring of 2 or 4 operand buffers
      tc::mbar_wait_soft(&sh.b_full, 0, p.err, 602, &sh.abort);   // weights resident (a wait in FRONT of the loop costs the uniform datapath)
      tc::mbar_wait_soft(&sh.acc_empty[acc], aph ^ 1u, p.err, 603, &sh.abort);
      tc::mbar_wait_soft(&sh.op_full[buf], ph, p.err, 604, &sh.abort);
      tc::tc_fence_after();
      for (uint32_t gi = 0; gi < p.m.slab_groups; ++gi) {
        const uint32_t d_tmem = tmem_m + acc * p.m.slab_cols + gi * p.m.cout_g;
        const uint64_t a_g = a_desc0 + (uint64_t)(buf * p.m.a_buf16 + gi * p.m.a_group16);
        const uint64_t b_g = b_desc0 + (uint64_t)(gi * p.m.b_group16);
        for (uint32_t r = 0; r < p.m.R; ++r)          // no tap / R: integer division has no uniform-datapath form
          for (uint32_t s2 = 0; s2 < p.m.R; ++s2)
            for (uint32_t j = 0; j < p.m.ksteps; ++j)
              tc::mma_f16_guarded(d_tmem, a_g + (uint64_t)(r * p.m.BW + s2 + j * p.m.a_kstep16),
                                  b_g + (uint64_t)((r * p.m.R + s2) * p.m.b_tap16 + j * p.m.b_kstep16), p.m.idesc,
                                  (r | s2 | j) != 0u, lead);
      }
      if (lead) { tc::mma_commit(&sh.op_empty[buf]); tc::mma_commit(&sh.acc_full[acc]); }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue: TMEM -> scale + bias -> fp32 NCHW
    const int q = warp - 4;
    const int pos = q * 32 + lane;                       // accumulator row = position of the zero-padded tile
    const int tb = pos / (p.THH * p.BW);
    const int rem = pos - tb * (p.THH * p.BW);
    const int th = rem / p.BW, wc = rem - th * p.BW;
    const int64_t plane = (int64_t)p.H * p.W;
    const int ch_first = g_first * p.cout_g;
    uint32_t item = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++item) {
      const int bt = tile / p.row_tiles, rt = tile - bt * p.row_tiles;
      const int b = bt * p.TB + tb, h = rt * p.TH + th;
      const bool valid = tb < p.TB && th < p.TH && wc < p.W && b < p.B && h < p.H;
      const uint32_t acc = item & 1u, aph = (item >> 1) & 1u;
      if (!tc::mbar_wait(&sh.acc_full[acc], aph, p.err, 605)) break;
      tc::tc_fence_after();
      float* orow = p.out + (((int64_t)b * p.Cout + ch_first) * p.H + h) * p.W + wc;
      for (int n0 = 0; n0 < slab_cols; n0 += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)slab_cols + (uint32_t)n0, r);
        tc::tmem_ld_wait();
        if (valid) {
          float sc[32], bs[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const float4 a = *reinterpret_cast<const float4*>(&sh.epi_scale[n0 + 4 * v]);
            const float4 c = *reinterpret_cast<const float4*>(&sh.epi_bias[n0 + 4 * v]);
            sc[4 * v] = a.x; sc[4 * v + 1] = a.y; sc[4 * v + 2] = a.z; sc[4 * v + 3] = a.w;
            bs[4 * v] = c.x; bs[4 * v + 1] = c.y; bs[4 * v + 2] = c.z; bs[4 * v + 3] = c.w;
          }
          float* op = orow + (int64_t)n0 * plane;
#pragma unroll
          for (int j = 0; j < 32; ++j, op += plane)
            if (n0 + j < slab_cols) *op = fmaf(__uint_as_float(r[j]), sc[j], bs[j]);
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&sh.acc_empty[acc]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols));
  }
}

static int plan(const mnb_conv_shape* s, Params& p, int& smem_bytes) {
  MNB_REQUIRE(s != nullptr, "conv shape is NULL");
  auto unsupported = [](const char* why) { return mnb_fail(MNB_E_UNSUPPORTED, "packed tc conv: %s", why); };
  p.B = s->batch; p.Cin = s->in_c; p.Cout = s->out_c; p.H = s->in_h; p.W = s->in_w; p.R = s->ker_h; p.G = s->groups;
  MNB_REQUIRE(p.B > 0 && p.Cin > 0 && p.Cout > 0 && p.H > 0 && p.W > 0 && p.G > 0 && p.Cin % p.G == 0 && p.Cout % p.G == 0,
              "bad conv shape");
  if (s->stride_h != 1 || s->stride_w != 1 || s->dil_h != 1 || s->dil_w != 1) return unsupported("stride/dilation != 1");
  if (s->ker_h != s->ker_w || (p.R & 1) == 0 || s->pad_h != p.R / 2 || s->pad_w != p.R / 2) return unsupported("not a 'same' odd square filter");
  p.pad = p.R / 2; p.cin_g = p.Cin / p.G; p.cout_g = p.Cout / p.G;
  if (p.cin_g % 16 || p.cout_g % 16 || p.cout_g > 256) return unsupported("channels per group");
  if (p.W > 64 || p.H > 255) return unsupported("image size");
  p.BW = p.W + 2 * p.pad;
  p.TH = std::min(p.H, 128 / p.BW);
  if (p.TH < 1) return unsupported("padded row wider than 128 positions");
  p.THH = p.TH + 2 * p.pad;
  p.TB = 1;
  if (p.pad == 0 && p.TH == p.H) p.TB = std::max(1, std::min(p.B, 128 / (p.H * p.W)));
  p.npos = p.TB * p.THH * p.BW;
  if (p.BW > 256 || p.THH > 256 || p.TB > 256) return unsupported("box dimension");
  p.b_group_bytes = p.R * p.R * p.cin_g * p.cout_g * 2;
  const int halo = (p.R - 1) * p.BW + (p.R - 1);
  // slab: as many groups as keep (a) the accumulators within 256 columns x 2, (b) the box's octet dimension <= 256,
  // (c) weights + two operand buffers inside shared memory
  int best = 0;
  for (int sg = p.G; sg >= 1; --sg) {
    if (p.G % sg || sg * p.cout_g > 256 || sg * p.cin_g / 8 > 256) continue;
    const int box = (sg * p.cin_g / 8) * p.npos * 16;
    const int op = (box + (halo + 128) * 16 + 1023) / 1024 * 1024;
    if (sg * p.b_group_bytes + 2 * op <= kMaxDynSmem) { best = sg; break; }
  }
  if (!best) return unsupported("weights of one group and two operand boxes do not fit in shared memory");
  p.slab_groups = best;
  p.n_slabs = p.G / best;
  p.op_box_bytes = (best * p.cin_g / 8) * p.npos * 16;
  p.op_bytes = (p.op_box_bytes + (halo + 128) * 16 + 1023) / 1024 * 1024;
  p.nbuf = (best * p.b_group_bytes + 4 * p.op_bytes <= kMaxDynSmem) ? 4 : 2;   // power of two: ring index = counter & mask
  p.nbuf_log2 = p.nbuf == 4 ? 2 : 1;
  p.off_b = p.nbuf * p.op_bytes;
  smem_bytes = p.off_b + best * p.b_group_bytes;
  p.row_tiles = (p.H + p.TH - 1) / p.TH;
  p.n_tiles = ((p.B + p.TB - 1) / p.TB) * p.row_tiles;
  int cols = 32;
  const int slack = (best * p.cout_g) % 32 ? 32 : 0;   // the epilogue reads 32-column chunks
  while (cols < NACC * best * p.cout_g + slack) cols <<= 1;
  if (cols > 512) return unsupported("accumulators exceed tensor memory");
  p.tmem_cols = cols;
  return 0;
}

// ---------------------------------------------------------------------------------------------- producer
// one warp = 32 consecutive positions of one OUTPUT channel octet of one image
__global__ void __launch_bounds__(256) bn_sign_packed_fwd_kernel(const float* __restrict__ x, int batch, int channels, int hw,
                                                                 int sg, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float* __restrict__ y,
                                                                 uint32_t* __restrict__ bits, uint4* __restrict__ xp) {
  const int lane = threadIdx.x & 31;
  const int c8n = channels / 8, p32n = hw / 32, cpg = channels / sg;
  const int64_t items = (int64_t)batch * c8n * p32n;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < items; w += nwarps) {
    const int p32 = (int)(w % p32n);
    const int64_t t = w / p32n;
    const int oc8 = (int)(t % c8n), b = (int)(t / c8n);
    const int pos = p32 * 32 + lane;
    uint32_t h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int oc = oc8 * 8 + j;
      const int c = sg > 1 ? (oc % sg) * cpg + oc / sg : oc;   // inverse of out[:, a*sg + b] = in[:, b*cpg + a]
      const int64_t fi = ((int64_t)b * channels + c) * hw + pos;
      const float bn = fmaf(__ldg(x + fi) - __ldg(mean + c), __ldg(gamma + c) * __ldg(invstd + c), __ldg(beta + c));
      const bool neg = bn < 0.f;
      if (y) y[((int64_t)b * channels + oc) * hw + pos] = neg ? -1.f : 1.f;
      const uint32_t word = __ballot_sync(0xffffffffu, fabsf(bn) < 1.f);
      if (lane == 0) bits[fi >> 5] = word;
      h[j] = neg ? 0xBF80u : 0x3F80u;   // bf16 -1 / +1
    }
    xp[((int64_t)b * c8n + oc8) * hw + pos] = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16),
                                                         h[6] | (h[7] << 16));
  }
}

}  // namespace tcpacked

extern "C" int mnb_bn_sign_fwd_packed(const float* x, int32_t batch, int32_t channels, int32_t hw, const float* mean,
                                      const float* invstd, const float* gamma, const float* beta, int32_t out_shuffle_groups,
                                      float* y, uint32_t* pass_bits, void* x_packed, mnb_stream_t stream) {
  MNB_REQUIRE(x && mean && invstd && gamma && beta && pass_bits && x_packed, "NULL bn_sign_fwd_packed pointer");   // y may be NULL
  MNB_REQUIRE(batch > 0 && channels > 0 && hw > 0, "bad bn_sign_fwd_packed shape");
  MNB_REQUIRE(out_shuffle_groups >= 1 && channels % out_shuffle_groups == 0, "shuffle groups %d do not divide %d channels",
              out_shuffle_groups, channels);
  if (channels % 8 || hw % 32 || (reinterpret_cast<uintptr_t>(x_packed) & 15))
    return mnb_fail(MNB_E_UNSUPPORTED, "packed producer needs channels %% 8 == 0, H*W %% 32 == 0, 16-byte aligned output");
  const int64_t warps = (int64_t)batch * (channels / 8) * (hw / 32);
  int blocks = (int)std::min<int64_t>(mnb_ceil_div(warps, 8), MNB_NUM_SMS * 16);
  tcpacked::bn_sign_packed_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      x, batch, channels, hw, out_shuffle_groups, mean, invstd, gamma, beta, y, pass_bits, reinterpret_cast<uint4*>(x_packed));
  MNB_LAUNCHED(1);
  return 0;
}

// host-only: the tile / slab / pipeline plan of a shape (for tests and sizing), out[8] = {slab_groups, n_slabs, nbuf,
// smem_bytes, tmem_cols, TH, TB, n_tiles}; returns 0 or MNB_E_UNSUPPORTED
extern "C" int mnb_fq_conv2d_fwd_packed_plan(const mnb_conv_shape* s, int32_t* out) {
  MNB_REQUIRE(out != nullptr, "NULL plan output");
  tcpacked::Params p{};
  int smem_bytes = 0;
  if (int e = tcpacked::plan(s, p, smem_bytes)) return e;
  const int v[8] = {p.slab_groups, p.n_slabs, p.nbuf, smem_bytes, p.tmem_cols, p.TH, p.TB, p.n_tiles};
  for (int i = 0; i < 8; ++i) out[i] = v[i];
  return 0;
}

extern "C" int mnb_fq_conv2d_fwd_packed_tc(const mnb_conv_shape* s, const void* x_packed, const int16_t* w_int,
                                           const float* w_scale, const float* bias, float* y, void* wpack_scratch,
                                           int32_t* err_flag, mnb_stream_t stream) {
  using namespace tcpacked;
  MNB_REQUIRE(s && x_packed && w_int && w_scale && y && wpack_scratch && err_flag, "NULL pointer");
  Params p{};
  int smem_bytes = 0;
  if (int e = plan(s, p, smem_bytes)) return e;
  p.w_pack = reinterpret_cast<const uint8_t*>(wpack_scratch);
  p.w_scale = w_scale; p.bias = bias; p.out = y; p.err = err_flag;
  {
    const uint32_t c8g = (uint32_t)p.cin_g / 8;
    p.m = Params::Mma{tc::make_idesc(1, 1, 1, 128, (uint32_t)p.cout_g), (uint32_t)p.npos * 16u, (uint32_t)p.cout_g * 16u,
                      (uint32_t)p.op_bytes >> 4, c8g * (uint32_t)p.npos, 2u * (uint32_t)p.npos, 2u * (uint32_t)p.cout_g,
                      c8g * (uint32_t)p.cout_g, (uint32_t)p.b_group_bytes >> 4, (uint32_t)p.cin_g / 16, (uint32_t)p.R,
                      (uint32_t)p.BW, (uint32_t)p.slab_groups, (uint32_t)(p.slab_groups * p.cout_g), (uint32_t)p.cout_g,
                      (uint32_t)p.nbuf - 1, (uint32_t)p.nbuf_log2, (uint32_t)p.off_b, (uint32_t)p.n_tiles};
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int total = p.G * p.b_group_bytes / 2;
  pack_weights_kernel<<<std::min(mnb_ceil_div(total, 256), MNB_NUM_SMS * 4), 256, 0, st>>>(
      w_int, reinterpret_cast<__nv_bfloat16*>(wpack_scratch), p.G, p.cin_g, p.cout_g, p.R * p.R);
  // xp[b][c8][h][w][8]: box traversal order (8, w, h, b, c8) so that a box lands as [c8][tb][row][col][8]
  const uint64_t HW = (uint64_t)p.H * p.W, C8 = (uint64_t)p.Cin / 8;
  uint64_t dims[5] = {8, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B, C8};
  uint64_t strides[4] = {16, (uint64_t)p.W * 16, C8 * HW * 16, HW * 16};
  uint32_t box[5] = {8, (uint32_t)p.BW, (uint32_t)p.THH, (uint32_t)p.TB, (uint32_t)(p.slab_groups * p.cin_g / 8)};
  CUtensorMap tmap;
  if (int e = mnb_make_tmap_strided(&tmap, x_packed, 2, 5, dims, strides, box)) return e;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(fwd_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (ce != cudaSuccess) return mnb_fail((int)ce, "cudaFuncSetAttribute: %s", cudaGetErrorString(ce));
    attr_set = true;
  }
  const int per_slab = std::max(1, std::min(p.n_tiles, MNB_NUM_SMS / p.n_slabs));
  fwd_packed_kernel<<<dim3(per_slab, p.n_slabs), NTHREADS, smem_bytes, st>>>(tmap, p);
  MNB_LAUNCHED(2);
  return 0;
}
