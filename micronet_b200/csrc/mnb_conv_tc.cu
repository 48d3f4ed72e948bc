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
  // driver-API call: make sure this host thread (e.g. an autograd worker) has the primary context bound - once per
  // thread (a cudaFree inside a stream capture would invalidate the capture)
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
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

int mnb_make_tmap_strided(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return mnb_fail(MNB_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  { static thread_local bool ctx_bound = false; if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; } }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i < rank - 1) gstr[i] = strides_bytes[i];
  }
  CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64
                           : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return mnb_fail(MNB_E_ARG, "cuTensorMapEncodeTiled (strided) failed with CUresult %d", (int)r);
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

// MN-major variant (the wgrad kernel's operand form): operands stored as buf[m/8][k][8 elements],
// i.e. D[128 x N] = sum_k At[k][m] * Bt[k][n] with both operands "transposed" in shared memory.
__global__ void __launch_bounds__(160) selftest_umma_mn_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ D, int N, int K, int* err) {
  __shared__ __align__(128) uint8_t a_s[128 * 64 * 2];
  __shared__ __align__(128) uint8_t b_s[128 * 64 * 2];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 128) {
    for (int k = 0; k < K; ++k) {
      reinterpret_cast<__nv_bfloat16*>(a_s)[((size_t)(tid >> 3) * K + k) * 8 + (tid & 7)] = __float2bfloat16_rn(A[(size_t)tid * K + k]);
      if (tid < N)
        reinterpret_cast<__nv_bfloat16*>(b_s)[((size_t)(tid >> 3) * K + k) * 8 + (tid & 7)] = __float2bfloat16_rn(B[(size_t)tid * K + k]);
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
    const uint32_t idesc = tc::make_idesc_major(1, 1, 1, 128, N, 1, 1);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad = tc::smem_desc_mnmajor_noswz(tc::smem_u32(a_s) + ks * 256, 128, K * 16);
      uint64_t bd = tc::smem_desc_mnmajor_noswz(tc::smem_u32(b_s) + ks * 256, 128, K * 16);
      tc::mma_f16(tmem, ad, bd, idesc, ks > 0);
    }
    tc::mma_commit(&done_bar);
  }
  if (warp < 4) {
    bool ok = tc::mbar_wait(&done_bar, 0, err, 102);
    tc::tc_fence_after();
    if (ok) {
      for (int n0 = 0; n0 < N; n0 += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + n0, r);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32 && n0 + j < N; ++j) D[(size_t)(warp * 32 + lane) * N + n0 + j] = __uint_as_float(r[j]);
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
  MNB_REQUIRE(K >= 32 && K <= (int8 == 1 ? 128 : 64) && K % 32 == 0, "self-test K out of range");
  if (int8 == 2) {
    MNB_REQUIRE(K <= 64, "MN-major self-test K <= 64");
    selftest_umma_mn_kernel<<<1, 160, 0, (cudaStream_t)stream>>>(A, B, D, N, K, err_flag);
  } else if (int8) selftest_umma_kernel<true><<<1, 160, 0, (cudaStream_t)stream>>>(A, B, D, N, K, err_flag);
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


// ------------------------------------------------------------------ micro-benchmark: MMA issue / execution rate
// `iters` back-to-back tcgen05.mma (M=128, N, K=16, bf16, K-major no-swizzle operands of zeros) from one
// thread, cycling over `n_acc` accumulators; A start address shifted by `a_shift16` 16-byte units.
// out[0] = SM cycles from first issue until the final commit arrives.
__global__ void __launch_bounds__(64) mma_rate_kernel(int N, int n_acc, int a_shift16, int iters, int mn_major,
                                                      long long* out, int* err) {
  __shared__ __align__(1024) uint8_t ab[48 * 1024 - 1024];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (int)sizeof(ab) / 16; i += 64) reinterpret_cast<uint4*>(ab)[i] = make_uint4(0, 0, 0, 0);
  tc::fence_proxy_async_smem();
  if (warp == 1) {
    tc::tmem_alloc<512>(&tmem_slot);
    if (lane == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = mn_major ? tc::make_idesc_major(1, 1, 1, 128, N, 1, 1) : tc::make_idesc(1, 1, 1, 128, N);
    const uint32_t a_addr = tc::smem_u32(ab) + (uint32_t)a_shift16 * 16u, b_addr = tc::smem_u32(ab) + 24 * 1024;
    const uint64_t ad = mn_major ? tc::smem_desc_mnmajor_noswz(a_addr, 128, 136 * 16) : tc::smem_desc_kmajor_noswz(a_addr, 136 * 16, 128);
    const uint64_t bd = mn_major ? tc::smem_desc_mnmajor_noswz(b_addr, 128, 136 * 16) : tc::smem_desc_kmajor_noswz(b_addr, (uint32_t)N * 16, 128);
    const uint32_t guard = lane == 0;
    const uint32_t mask = (uint32_t)n_acc - 1;  // n_acc is a power of two
    const long long t0 = clock64();
    if (a_shift16 >= 100) {
      if (lane == 0) for (int i = 0; i < iters; ++i) tc::mma_f16(tmem + (uint32_t)((i & mask) * N), ad + (i & 3), bd, idesc, 1);
    } else {
      for (int i = 0; i < iters; ++i) tc::mma_f16_guarded(tmem + (uint32_t)((i & mask) * N), ad + (i & 3), bd, idesc, 1, guard);
    }
    const long long t1 = clock64();
    if (lane == 0) tc::mma_commit(&bar);
    bool ok = tc::mbar_wait(&bar, 0, err, 501);
    const long long t2 = clock64();
    if (lane == 0) { out[0] = ok ? (t2 - t0) : -1; out[1] = t1 - t0; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem);
}

extern "C" int mnb_selftest_mma_rate(int32_t N, int32_t n_acc, int32_t a_shift16, int32_t iters, int32_t mn_major,
                                     int64_t* out2, int32_t* err_flag, mnb_stream_t stream) {
  MNB_REQUIRE(out2 && err_flag && N >= 16 && N <= 256 && N % 16 == 0 && n_acc >= 1 && n_acc * N <= 512, "bad mma_rate arguments");
  mma_rate_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(N, n_acc, a_shift16, iters, mn_major, reinterpret_cast<long long*>(out2), err_flag);
  MNB_LAUNCHED(1);
  return 0;
}
