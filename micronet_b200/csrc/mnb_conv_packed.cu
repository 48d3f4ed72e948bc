// Packed bf16 output of the BatchNorm + binarizer producer: mnb_bn_sign_fwd_packed = mnb_bn_sign_fwd plus
// xp[b][c/8][h][w][8] (bf16 +-1, OUTPUT channel order) - the operand plane the packed-operand tensor-core family
// (mnb_pk.cu) reads with TMA, so that a wbwtab layer behind a fused BatchNorm + binarizer needs no pack pass and no
// converter warps (fused.BNSignFn -> functional._pk_forward(prepacked=...)).
// (The round-1 consumer kernel that lived in this file was superseded by mnb_pk.cu before it ever ran and is gone.)
#include <cuda_bf16.h>

#include <algorithm>

#include "mnb_common.cuh"

namespace tcpacked {

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
