// Developer probe: one TMA box load of doubles (3-D tensor map, box 130 x 5 x 1 starting at (-1, y, z)) with the tensor map
// (a) in a small kernel parameter, (b) behind 9 KB of other parameters, (c) in global memory. Prints which forms work.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstdlib>
__constant__ int c_boxx, c_x0;
struct Maps { alignas(64) unsigned char m[128]; };
struct Big { double pad[1100]; };
__device__ __forceinline__ void load_box(const void* map, double* out, int x, int y, int z) {
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned s0 = ((unsigned)__cvta_generic_to_shared(smem) + 127u) & ~127u;
  const unsigned bar = s0 + 6272u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar), "r"(unsigned(c_boxx*5*8)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 :: "r"(s0), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
  }
  __syncthreads();
  asm volatile("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra W_%=;\n}\n" :: "r"(bar), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < 650; i += blockDim.x) {
    double v; asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(s0 + 8u*i));
    out[i] = v;
  }
}
__global__ void k_small(const __grid_constant__ Maps m, double* out) { load_box(m.m, out, c_x0, 3, 2); }
__global__ void k_big(const __grid_constant__ Big b, int dummy, const __grid_constant__ Maps m, double* out) { if (b.pad[3] == 12345.0) out[700] = 1; load_box(m.m, out, c_x0, 3, 2); }
__global__ void k_glob(const Maps* m, double* out) { load_box(m->m, out, c_x0, 3, 2); }
typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int BOXX = argc > 1 ? atoi(argv[1]) : 130, X0 = argc > 2 ? atoi(argv[2]) : -1, DT = argc > 3 ? atoi(argv[3]) : 0, NXARG = argc > 4 ? atoi(argv[4]) : 512;
  const int nx = NXARG, ny = 64, nz = 8;
  std::vector<double> h(size_t(nx)*ny*nz);
  for (size_t i = 0; i < h.size(); ++i) h[i] = double(i);
  double *d, *out; cudaMalloc(&d, h.size()*8); cudaMalloc(&out, 1024*8);
  cudaMemcpy(d, h.data(), h.size()*8, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  Maps m;
  cuuint64_t dims[3] = { nx, ny, nz }, str[2] = { nx*8ull, nx*ny*8ull };
  cuuint32_t box[3] = { cuuint32_t(BOXX), 5, 1 }, es[3] = { 1, 1, 1 };
  CUresult r = ((enc_t)p)((CUtensorMap*)m.m, DT == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : DT == 1 ? CU_TENSOR_MAP_DATA_TYPE_INT64 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("box %d x0 %d dtype %d nx %d: encode rc=%d\n", BOXX, X0, DT, nx, int(r));
  cudaMemcpyToSymbol(c_boxx, &BOXX, 4); cudaMemcpyToSymbol(c_x0, &X0, 4);
  Maps* dm; cudaMalloc(&dm, sizeof(Maps)); cudaMemcpy(dm, &m, sizeof(Maps), cudaMemcpyHostToDevice);
  Big b{}; 
  for (int form = 0; form < 3; ++form) {
    cudaMemset(out, 0, 1024*8);
    if (form == 0) k_small<<<1, 128, 6272+8+128>>>(m, out);
    if (form == 1) k_big<<<1, 128, 6272+8+128>>>(b, 1, m, out);
    if (form == 2) k_glob<<<1, 128, 6272+8+128>>>(dm, out);
    cudaError_t e = cudaDeviceSynchronize();
    double ho[650]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    // expected: row j, col i -> element (x = -1+i, y = 3+j, z = 2); x = -1 is out of bounds -> 0
    int bad = 0;
    for (int j = 0; j < 5; ++j) for (int i = 0; i < BOXX; ++i) {
      const int x = X0+i;
      const double want = (x < 0 || x >= nx) ? 0.0 : double((size_t(2)*ny + (3+j))*nx + x);
      if (ho[j*BOXX+i] != want) ++bad;
    }
    printf("form %d: %s, mismatches %d\n", form, cudaGetErrorString(e), bad);
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
