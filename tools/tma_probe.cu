// Stand-alone probe of the tiled TMA load the VIO tap staging uses (one B200): which descriptor / addressing form works.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o /tmp/tma_probe tools/tma_probe.cu && /tmp/tma_probe
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../fast_livo2_b200/csrc/esikf_dev.cuh"
using namespace esikf;

struct Maps {
  alignas(64) unsigned char map[4][128];
  int enabled;
};

__global__ void probe_static(const __grid_constant__ CUtensorMap m, int x, int y, int bytes, unsigned char *out) {
  __shared__ alignas(128) unsigned char tile[1408];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_2d(tile, &m, x, y, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
__global__ void probe_dynamic(const __grid_constant__ Maps m, int level, int x, int y, int bytes, unsigned char *out) {
  __shared__ alignas(128) unsigned char tile[1408];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const Maps *mp = m.enabled ? &m : nullptr;
  if (mp && threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    tma_load_2d(tile, mp->map[level], x, y, &bar);
  }
  if (mp) mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int W = 640, H = 512;
  std::vector<unsigned char> img(W * H);
  for (int i = 0; i < W * H; i++) img[i] = (unsigned char)((i * 2654435761u) >> 24);
  unsigned char *d_img, *d_out;
  cudaMalloc(&d_img, W * H);
  cudaMalloc(&d_out, 4096);
  cudaMemcpy(d_img, img.data(), W * H, cudaMemcpyHostToDevice);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  maps.enabled = 1;
  for (int l = 0; l < 4; l++) {
    const cuuint64_t dims[2] = {W, H};
    const cuuint64_t strides[1] = {W};
    const cuuint32_t box[2] = {16u << l, 11u << l};
    const cuuint32_t estr[2] = {1u, 1u << l};
    CUresult r = ((encode_fn)fn)(reinterpret_cast<CUtensorMap *>(maps.map[l]), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode level %d: %d\n", l, (int)r);
  }
  std::vector<unsigned char> out(4096);
  const int x0 = 101, y0 = 57;
  for (int variant = 0; variant < 2; variant++)
    for (int l = 0; l < 4; l++) {
      const int inner = 16 << l, bytes = 11 * inner, s = 1 << l;
      cudaMemset(d_out, 0xee, 4096);
      if (variant == 0) {
        CUtensorMap m;
        memcpy(&m, maps.map[l], 128);
        probe_static<<<1, 64>>>(m, x0, y0, bytes, d_out);
      } else {
        probe_dynamic<<<1, 64>>>(maps, l, x0, y0, bytes, d_out);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("variant %d level %d: %s\n", variant, l, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(out.data(), d_out, bytes, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 11; r++)
        for (int c = 0; c < inner; c++)
          if (out[r * inner + c] != img[(y0 + r * s) * W + x0 + c]) bad++;
      printf("variant %s level %d (rows every %d, %d bytes a row): %d mismatching bytes of %d\n", variant ? "struct+dynamic index" : "single descriptor", l, s, inner, bad, bytes);
    }
  return 0;
}
