// Stand-alone probe of the tiled TMA load the VIO tap staging uses (one B200): which descriptor / addressing form works.
// One variant per process (an illegal instruction kills the context):  for v in 0 1 2 3 4 5 6 7; do tools/bin/tma_probe $v; done
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o tools/bin/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../fast_livo2_b200/csrc/esikf_dev.cuh"
using namespace esikf;

__device__ __forceinline__ void tma_load_2d_plain(void *dst_smem, const void *tmap, int x, int y, unsigned long long *bar) {  // CUTLASS SM90_TMA_LOAD_2D spelling
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst_smem)), "l"(tmap),
               "r"(smem_u32(bar)), "r"(x), "r"(y)
               : "memory");
}

template <int FORM>
__device__ __forceinline__ void body(const void *desc, int x, int y, int bytes, unsigned char *out) {
  __shared__ alignas(128) unsigned char tile[4096];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, bytes);
    if (FORM == 0) tma_load_2d(tile, desc, x, y, &bar); else tma_load_2d_plain(tile, desc, x, y, &bar);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
__global__ void probe_param(const __grid_constant__ CUtensorMap m, int x, int y, int bytes, unsigned char *out) { body<0>(&m, x, y, bytes, out); }
__global__ void probe_param_plain(const __grid_constant__ CUtensorMap m, int x, int y, int bytes, unsigned char *out) { body<1>(&m, x, y, bytes, out); }
__global__ void probe_global(const CUtensorMap *m, int x, int y, int bytes, unsigned char *out) { body<0>(m, x, y, bytes, out); }

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                              CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 640, H = 512;
  std::vector<unsigned char> img(W * H);
  for (int i = 0; i < W * H; i++) img[i] = (unsigned char)((i * 2654435761u) >> 24);
  unsigned char *d_img, *d_out;
  cudaMalloc(&d_img, W * H);
  cudaMalloc(&d_out, 4096);
  cudaMemcpy(d_img, img.data(), W * H, cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0xee, 4096);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  // variants: 0 u8 box 16x11 param | 1 same, CUTLASS spelling | 2 same, descriptor in global memory | 3 u8 box 64x11 | 4 u8 box 32x22 rows every 2
  //           5 uint32 view (W/4 x H), box 4x11 | 6 u8 box 16x11 at an aligned start (x0 = 96) | 7 u8 box 128x8 (a plain wide tile)
  CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
  cuuint64_t dims[2] = {W, H}, strides[1] = {W};
  cuuint32_t box[2] = {16, 11}, estr[2] = {1, 1};
  int x0 = 101, y0 = 57, esz = 1;
  if (variant == 3) box[0] = 64;
  if (variant == 4) box[0] = 32, box[1] = 22, estr[1] = 2;
  if (variant == 5) dt = CU_TENSOR_MAP_DATA_TYPE_UINT32, dims[0] = W / 4, box[0] = 4, x0 = 25, esz = 4;
  if (variant == 6) x0 = 96;
  if (variant == 7) box[0] = 128, box[1] = 8, x0 = 128;
  alignas(64) CUtensorMap m;
  CUresult r = ((encode_fn)fn)(&m, dt, 2, d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int rows = (box[1] + estr[1] - 1) / estr[1], inner = box[0] * esz, bytes = rows * inner;
  printf("variant %d: encode %d, box %ux%u stride %u, %d bytes\n", variant, (int)r, box[0], box[1], estr[1], bytes);
  fflush(stdout);
  if (variant == 1) probe_param_plain<<<1, 64>>>(m, x0, y0, bytes, d_out);
  else if (variant == 2) {
    CUtensorMap *dm;
    cudaMalloc(&dm, sizeof(m));
    cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    probe_global<<<1, 64>>>(dm, x0, y0, bytes, d_out);
  } else probe_param<<<1, 64>>>(m, x0, y0, bytes, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    return 1;
  }
  std::vector<unsigned char> out(4096);
  cudaMemcpy(out.data(), d_out, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int rr = 0; rr < rows; rr++)
    for (int c = 0; c < inner; c++)
      if (out[rr * inner + c] != img[(y0 + rr * (int)estr[1]) * W + x0 * esz + c]) bad++;
  printf("variant %d: OK run, %d mismatching bytes of %d\n", variant, bad, bytes);
  return 0;
}
