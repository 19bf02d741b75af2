// Standalone probe of the 2-D TMA tile load used by k_refine (debugging aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int bytes, uint32_t* out, int nwords) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"((uint32_t)__cvta_generic_to_shared(sm)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(x), "r"(y), "r"(b)
        : "memory");
  }
  uint32_t ok = 0;
  for (int spin = 0; !ok && spin < (1 << 22); ++spin)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(b) : "memory");
  __syncthreads();
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(sm)[i];
  if (threadIdx.x == 0) out[nwords] = ok;
}

int run(EncodeTiledFn fn, int W, int H, int bx, int by, int x, int y, int elem) {
  std::vector<uint32_t> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i + 1;
  void* d; cudaMalloc(&d, (size_t)W * H * elem);
  cudaMemcpy(d, h.data(), (size_t)W * H * (elem == 4 ? 4 : 1), cudaMemcpyHostToDevice);
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  cuuint64_t strides[1] = {(cuuint64_t)W * elem};
  cuuint32_t box[2] = {(cuuint32_t)bx, (cuuint32_t)by};
  cuuint32_t es[2] = {1, 1};
  CUresult rc = fn(&tm, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("W=%d H=%d box=%dx%d at (%d,%d) elem=%d encode rc=%d ", W, H, bx, by, x, y, elem, (int)rc);
  if (rc != CUDA_SUCCESS) { printf("\n"); cudaFree(d); return 1; }
  int bytes = bx * by * elem, nwords = bytes / 4;
  uint32_t* out; cudaMalloc(&out, (nwords + 1) * 4);
  k<<<1, 128, bytes>>>(tm, x, y, bytes, out, nwords);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<uint32_t> o(nwords + 1);
    cudaMemcpy(o.data(), out, (nwords + 1) * 4, cudaMemcpyDeviceToHost);
    printf("ok=%u first=%u ", o[nwords], o[0]);
    if (elem == 4) {
      int bad = 0;
      for (int r = 0; r < by; ++r) for (int c = 0; c < bx; ++c) {
        int gx = x + c, gy = y + r;
        uint32_t exp = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? (uint32_t)(gy * W + gx) + 1 : 0;
        bad += o[r * bx + c] != exp;
      }
      printf("mismatches=%d", bad);
    }
  }
  printf("\n");
  cudaFree(d); cudaFree(out);
  return e != cudaSuccess;
}

int main(int argc, char** argv) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
  if (!p) return 1;
  EncodeTiledFn fn = (EncodeTiledFn)p;
  int v = argc > 1 ? atoi(argv[1]) : 0;
  if (v == 0) return run(fn, 256, 256, 32, 26, 16, 16, 4);
  if (v == 1) return run(fn, 256, 256, 28, 26, 16, 16, 4);
  if (v == 2) return run(fn, 256, 256, 28, 26, -5, -5, 4);
  if (v == 3) return run(fn, 4096, 4096, 68, 66, 471, 471, 4);
  if (v == 4) return run(fn, 320, 280, 32, 26, 16, 16, 1);
  if (v == 5) return run(fn, 256, 256, 28, 26, 250, 250, 4);
  if (v == 6) return run(fn, 256, 256, 28, 26, 17, 16, 4);
  if (v == 7) return run(fn, 256, 256, 28, 26, 16, -5, 4);
  if (v == 8) return run(fn, 4096, 4096, 68, 66, 480, 471, 4);
  if (v == 9) return run(fn, 256, 256, 28, 26, -4, 16, 4);
  if (v == 10) return run(fn, 256, 256, 28, 26, 240, 16, 4);
  if (v == 11) return run(fn, 4096, 4096, 68, 66, 472, 471, 4);
  return 0;
}
