// rx_hop.cu -- candidate search of the hop reversal on the GPU: every clock value with the known
// CLK1-6 whose hop lands on the first observed channel (the reference walks its 2^27-entry table with stride 64,
// lib/piconet_impl.cc:96-129).  One thread per candidate clock evaluates the hop selection kernel (hop_select.hpp);
// a warp's verdicts leave as one ballot word, so the host rebuilds the ascending candidate list from 256 KB.
#include "../../include/btb200.h"
#include "hop_select.hpp"
#include <cuda_runtime.h>
#include <vector>

namespace btb200 {
__global__ void k_hop_candidates(uint32_t addr, int afh, int aliased, uint32_t clock6, int first_channel, uint32_t *__restrict__ masks)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;           // candidate index: clock = clock6 + 64 i, i < 2^21
  const uint32_t clock = clock6 + 64u * i;
  int ch = hop_select(addr, afh != 0, clock);
  if (aliased) ch = hop_aliased(ch);
  const uint32_t m = __ballot_sync(0xffffffffu, ch == first_channel);
  if ((threadIdx.x & 31) == 0) masks[i >> 5] = m;
}
}  // namespace btb200

extern "C" int btb200_hop_select(uint32_t address28, int afh, uint32_t clock)
{
  return btb200::hop_select(address28 & 0xfffffffu, afh != 0, clock & 0x7ffffffu);
}

extern "C" int btb200_hop_candidates(int device, uint32_t address28, int afh, int aliased, uint32_t clock6, int first_channel,
                                     uint32_t *out, uint32_t cap, uint32_t *count)
{
  if (!out || !count || clock6 >= 64 || device < 0) return BTB200_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return BTB200_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return BTB200_ERR_NO_DEVICE;
  const uint32_t n = 1u << 21, words = n / 32;
  uint32_t *d = nullptr;
  if (cudaMalloc(&d, words * sizeof(uint32_t)) != cudaSuccess) return BTB200_ERR_NOMEM;
  btb200::k_hop_candidates<<<n / 256, 256>>>(address28 & 0xfffffffu, afh, aliased, clock6, first_channel, d);
  std::vector<uint32_t> h(words);
  const cudaError_t e = cudaMemcpy(h.data(), d, words * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return BTB200_ERR_CUDA;
  uint32_t k = 0, total = 0;
  for (uint32_t w = 0; w < words; w++) {
    uint32_t m = h[w];
    while (m) {
      const uint32_t b = (uint32_t)__builtin_ctz(m);
      m &= m - 1;
      if (k < cap) out[k++] = clock6 + 64u * (w * 32u + b);
      total++;
    }
  }
  *count = total;
  return BTB200_OK;
}
