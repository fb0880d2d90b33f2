// Library-wide state of libcvvae_b200 (error string, launch counter, driver entry point) and the
// UMMA descriptor probe used by the GPU test-suite.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cvvae {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// One CTA: TMA a [rows_total][64] 16-bit matrix and a [n][64] matrix into SWIZZLE_128B shared memory, run a
// single 128 x n x 64 UMMA whose A descriptor starts `row_shift` rows (128 B each) into the slab and whose 8-row
// groups are `sbo_rows` rows apart (8 = dense; 16 = every other group, i.e. a tile of 8-position image rows cut out
// of a wider slab).
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB, float* out, int n,
                                                    int rows_total, int row_shift, int base_offset_mode, int sbo_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // rows_total * 128 B (<= 48 KB)
  uint8_t* sB = smem + 49152;          // n * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152 + 32768);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::mbar_init(done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(bar, static_cast<uint32_t>(rows_total + n) * 128u);
    ptx::tma_load_3d(sA, &tmA, bar, 0, 0, 0);                                   // two boxes of rows_total / 2 rows
    ptx::tma_load_3d(sA + (rows_total / 2) * 128, &tmA, bar, 0, rows_total / 2, 0);
    ptx::tma_load_3d(sB, &tmB, bar, 0, 0, 0);
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();
    const uint32_t a0 = ptx::smem_u32(sA) + static_cast<uint32_t>(row_shift) * 128u;
    const uint32_t b0 = ptx::smem_u32(sB);
    const uint32_t bo = base_offset_mode ? ((a0 >> 7) & 7u) : 0u;
    const uint32_t idesc = ptx::umma_idesc_f16(0, 128, n);
    for (int k = 0; k < 4; ++k)
      ptx::umma_f16(tmem, ptx::umma_desc_k_sw128(a0 + k * 32, static_cast<uint32_t>(sbo_rows) * 128u, bo),
                    ptx::umma_desc_k_sw128(b0 + k * 32, 1024), idesc, k > 0);
    ptx::umma_commit(done);
  }
  __syncwarp();
  ptx::mbar_wait(done, 0);
  ptx::tc_fence_after();
  const int r = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < n; ++j) out[r * n + c0 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 256);
  }
}

}  // namespace cvvae

using namespace cvvae;

extern "C" const char* cvvae_last_error(void) { return g_err; }
extern "C" int cvvae_abi_version(void) { return CVVAE_ABI_VERSION; }
extern "C" int64_t cvvae_launch_count(void) { return g_launches.load(); }

extern "C" int cvvae_probe_umma_shift(const void* a_rows, const void* b_rows, float* out, int32_t n, int32_t row_shift,
                                      int32_t base_offset_mode, int32_t sbo_rows, void* stream_) {
  CVVAE_CHECK_ARG(a_rows && b_rows && out && n >= 16 && n <= 256 && n % 16 == 0 && row_shift >= 0 && row_shift <= 64 &&
                      sbo_rows >= 8 && sbo_rows <= 16,
                  "cvvae_probe_umma_shift: bad argument");
  PFN_encodeTiled enc = get_encode_tiled();
  CVVAE_CHECK_ARG(enc, "cuTensorMapEncodeTiled entry point not available");
  const int rows_total = 320;   // 15 groups x 16 rows + 8 + shift 64 <= 320
  CUtensorMap tmA, tmB;
  cuuint32_t estr[3] = {1, 1, 1};
  {
    cuuint64_t dims[3] = {64, (cuuint64_t)rows_total, 1};
    cuuint64_t strides[2] = {128, 128ull * rows_total};
    cuuint32_t box[3] = {64, (cuuint32_t)(rows_total / 2), 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(a_rows), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVVAE_CHECK_ARG(r == CUDA_SUCCESS, "probe: tensor map A failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[3] = {64, (cuuint64_t)n, 1};
    cuuint64_t strides[2] = {128, 128ull * n};
    cuuint32_t box[3] = {64, (cuuint32_t)n, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(b_rows), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVVAE_CHECK_ARG(r == CUDA_SUCCESS, "probe: tensor map B failed (%d)", (int)r);
  }
  const size_t smem = 1024 + 49152 + 32768 + 64;
  CVVAE_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream_)>>>(tmA, tmB, out, n, rows_total, row_shift, base_offset_mode,
                                                                      sbo_rows);
  CVVAE_LAUNCH_CHECK();
  return CVVAE_OK;
}
