// Pixel pre/post-processing either side of encode()/decode(), fused into one pass each:
//   uint8 [T,H,W,3] frames  ->  16-bit [1,3,T,H,W] in [-1,1]   (cvvae_inference_video.py:30-38: `.half() / 127.5 - 1.0`)
//   16-bit [1,3,T,H,W]      ->  uint8 [T,H,W,3]                (cvvae_inference_video.py:47-50: clamp, +1, *127.5, uint8)
// Bit-exact with the reference expressions: every intermediate is rounded to the 16-bit type exactly where PyTorch
// rounds it (each elementwise op reads/writes the tensor dtype, arithmetic in fp32).
#include "common.cuh"

namespace cvvae {

template <int DT>
__global__ void __launch_bounds__(256) u8_to_f16_kernel(const uint8_t* __restrict__ in, typename Elem<DT>::T* __restrict__ out,
                                                        long long thw) {
  using E = Elem<DT>;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < thw;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const typename E::T h = E::from_f(static_cast<float>(in[i * 3 + c]));   // .half() (exact for 0..255; bf16 rounds)
      const typename E::T q = E::from_f(E::to_f(h) / 127.5f);                  // / 127.5
      out[c * thw + i] = E::from_f(E::to_f(q) - 1.0f);                         // - 1.0
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(256) f16_to_u8_kernel(const typename Elem<DT>::T* __restrict__ in, uint8_t* __restrict__ out,
                                                        long long thw) {
  using E = Elem<DT>;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < thw;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = E::to_f(in[c * thw + i]);
      v = fminf(fmaxf(v, -1.0f), 1.0f);                 // torch.clamp(x, -1, 1)   (NaN propagates like torch: stays NaN)
      const typename E::T a = E::from_f(v + 1.0f);      // + 1.0
      const typename E::T m = E::from_f(E::to_f(a) * 127.5f);  // * 127.5
      out[i * 3 + c] = static_cast<uint8_t>(static_cast<int>(E::to_f(m)));   // .to(uint8): truncation
    }
  }
}

}  // namespace cvvae

using namespace cvvae;

extern "C" int cvvae_video_u8_to_f16(const uint8_t* thwc, void* out_cthw, int32_t T, int32_t H, int32_t W, int32_t dtype,
                                     void* stream) {
  CVVAE_CHECK_ARG(thwc && out_cthw && T > 0 && H > 0 && W > 0, "cvvae_video_u8_to_f16: bad argument");
  const long long thw = 1ll * T * H * W;
  const unsigned blocks = static_cast<unsigned>(thw / 256 + 1 < 16ll * num_sms() ? thw / 256 + 1 : 16ll * num_sms());
  CVVAE_DISPATCH_DTYPE(dtype, {
    u8_to_f16_kernel<DT><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        thwc, reinterpret_cast<typename Elem<DT>::T*>(out_cthw), thw);
  });
  CVVAE_LAUNCH_CHECK();
  return CVVAE_OK;
}

extern "C" int cvvae_video_f16_to_u8(const void* in_cthw, uint8_t* thwc, int32_t T, int32_t H, int32_t W, int32_t dtype,
                                     void* stream) {
  CVVAE_CHECK_ARG(thwc && in_cthw && T > 0 && H > 0 && W > 0, "cvvae_video_f16_to_u8: bad argument");
  const long long thw = 1ll * T * H * W;
  const unsigned blocks = static_cast<unsigned>(thw / 256 + 1 < 16ll * num_sms() ? thw / 256 + 1 : 16ll * num_sms());
  CVVAE_DISPATCH_DTYPE(dtype, {
    f16_to_u8_kernel<DT><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const typename Elem<DT>::T*>(in_cthw), thwc, thw);
  });
  CVVAE_LAUNCH_CHECK();
  return CVVAE_OK;
}
